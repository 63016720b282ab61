"""Profiling driver: two sweeps of BASELINE config 5 on one GPU (65536^2 lattice bit-packed along x,
512 MiB, HBM-streamed)."""
import os
import sys

sys.path.insert(0, os.getcwd())
import pyisingmontecarlo_b200 as pkg

lat = pkg.SingleLattice2D(65536, seed=3)
lat.sweeps([0.44, 0.44])
print("ok", lat.strip.stats())
