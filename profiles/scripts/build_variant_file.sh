#!/bin/bash
# usage: profiles/scripts/build_variant_file.sh NAME FILE "-DFLAGS"  -> variants/libising_NAME.so with FILE.cu rebuilt with the flags
set -e
cd "$(dirname "$0")/../.."
NAME=$1; F=$2; EXTRA=$3
CS=pyisingmontecarlo_b200/csrc
OBJ=variants/obj_$NAME; mkdir -p $OBJ
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden"
nvcc $FLAGS $EXTRA -Xptxas=-v -c $CS/$F.cu -o $OBJ/$F.o > $OBJ/$F.log 2>&1
OTHERS=$(ls $CS/_obj/*.o | grep -v "/$F.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libising_$NAME.so $OBJ/$F.o $OTHERS
echo built variants/libising_$NAME.so
