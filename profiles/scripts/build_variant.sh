#!/bin/bash
# usage: profiles/scripts/build_variant.sh NAME "-DISING_ROWS_THREADS=128 ..."   -> variants/libising_NAME.so
# rebuilds only the row-walk translation units with the extra flags, links with the stock objects
set -e
cd "$(dirname "$0")/../.."
NAME=$1; EXTRA=$2
CS=pyisingmontecarlo_b200/csrc
OBJ=variants/obj_$NAME; mkdir -p $OBJ
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden"
for f in sweep_rows3d sweep_rows2d; do
  nvcc $FLAGS $EXTRA -Xptxas=-v -c $CS/$f.cu -o $OBJ/$f.o > $OBJ/$f.log 2>&1 &
done
wait
OTHERS=$(ls $CS/_obj/*.o | grep -v sweep_rows)
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libising_$NAME.so $OBJ/sweep_rows3d.o $OBJ/sweep_rows2d.o $OTHERS
echo built variants/libising_$NAME.so
