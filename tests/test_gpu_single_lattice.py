"""Config 5 path: one large bit-packed 2D lattice in row strips with halo exchange."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "exact_2d_ising.json")))


def test_single_strip_matches_mirror(native, oracle):
    import pyisingmontecarlo_b200 as pkg

    for (Lx, Ly, j, planes, rounds) in [(64, 6, -1.0, 6, 10), (128, 8, 1.0, 5, 7), (192, 4, -0.5, 7, 10)]:
        lat = pkg.SingleLattice2D(Lx, Ly, j=j, seed=77, planes=planes, rounds=rounds)
        betas = [0.3, 0.44, 0.8, 0.44]
        en = []
        for b in betas:
            lat.sweeps([b])
            en.append(lat.energy())
        en_ref, st_ref = oracle.msc_mirror_single(Lx, Ly, j, 77, betas, planes, rounds)
        assert (lat.local_rows() == st_ref).all()
        assert en == list(en_ref)
        s = st_ref.astype(int) * 2 - 1
        assert lat.magnetization() == s.sum()


def test_strips_reproduce_single_strip(native):
    """Three strips exchanging boundary rows by hand (what the NCCL send/recv does) == one strip."""
    ctx = native.Context.get(0)
    Lx, Ly = 128, 24
    whole = native.Strip(ctx, Lx, Ly, 0, Ly, -1.0, 5)
    parts = [native.Strip(ctx, Lx, Ly, lo, hi, -1.0, 5) for lo, hi in ((0, 6), (6, 16), (16, 24))]
    for beta in (0.4, 0.5, 0.3):
        for colour in (0, 1):
            whole.wrap_local(1 - colour)
            for k, p in enumerate(parts):
                up, down = parts[(k - 1) % 3], parts[(k + 1) % 3]
                p.set_ghost(1 - colour, 0, up.get_boundary(1 - colour, 1))
                p.set_ghost(1 - colour, 1, down.get_boundary(1 - colour, 0))
            whole.phase(colour, beta)
            for p in parts:
                p.phase(colour, beta)
    assert (np.concatenate([p.rows() for p in parts]) == whole.rows()).all()


def test_large_lattice_vs_onsager(native):
    """8192 x 8192 single lattice (64 Mi spins): equilibrium energy and magnetisation on both
    sides of T_c against Onsager."""
    import pyisingmontecarlo_b200 as pkg

    L = 8192
    lat = pkg.SingleLattice2D(L, seed=3)
    lat.sweeps([0.30] * 100)
    e = lat.energy() / L**2
    assert abs(e - GOLD["onsager"]["b0.3"]["e_per_site"]) < 1.5e-3, e
    assert abs(lat.magnetization()) / L**2 < 2e-3
    lat.set_all(True)
    lat.sweeps([0.60] * 100)
    ex = GOLD["onsager"]["b0.6"]
    assert abs(lat.energy() / L**2 - ex["e_per_site"]) < 1.5e-3
    assert abs(lat.magnetization() / L**2 - ex["m"]) < 1.5e-3


def test_multi_gpu_parity_under_torchrun(native):
    """Runs tests/multi_gpu_check.py on every visible GPU (skipped on a 1-GPU box)."""
    import subprocess
    import sys

    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
