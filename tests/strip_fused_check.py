"""Run by tests/test_gpu_single_lattice.py in a process of its own with ISING_STRIP_FUSE=1 and
ISING_STRIP_FUSE_MIN_ROWS=1, which make ising_strip_sweeps take the fused two-colour pass even
for short bands: the fused sweep must produce the bits of the CPU mirror and of the two-phase path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import oracle_lib
from pyisingmontecarlo_b200 import _native as nat

assert os.environ.get("ISING_STRIP_FUSE_MIN_ROWS") == "1" and os.environ.get("ISING_STRIP_FUSE") == "1"
ctx = nat.Context.get(0)
betas = [0.3, 0.44, 0.8, 0.44, 0.5, 0.2, 0.6]
for Lx, Ly, k, j in ((256, 32, 2, -1.0), (1024, 40, 1, 1.0), (8192, 64, 3, -1.0), (65536, 48, 4, -1.0),
                     (65536, 16, 8, -1.0), (128, 16, 2, -1.0)):
    fused = nat.Strip(ctx, Lx, Ly, 0, Ly, j, 21, ghost=2 * k)
    fused.stats(reset=True)
    fused.sweeps(betas, None, k)
    launches = fused.stats()["launches"]
    expect_fused = Lx % 256 == 0
    # one launch per sweep when fused, two otherwise
    assert launches == (len(betas) if expect_fused else 2 * len(betas)), (Lx, Ly, k, launches)
    phases = nat.Strip(ctx, Lx, Ly, 0, Ly, j, 21)
    for beta in betas:
        for colour in (0, 1):
            phases.wrap_local(1 - colour)
            phases.phase(colour, beta)
    a, b = fused.rows(), phases.rows()
    assert (a == b).all(), (Lx, Ly, k, int((a != b).sum()))
    if Lx * Ly <= 8192 * 64:
        _, st_ref = oracle_lib.msc_mirror_single(Lx, Ly, j, 21, betas, 6, 7)
        assert (a == st_ref).all(), (Lx, Ly, k)
    assert fused.global_sums() == phases.global_sums()
    fused.close(); phases.close()
print("STRIP_FUSED_OK")
