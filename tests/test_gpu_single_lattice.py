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


@pytest.mark.parametrize("k", [0, 1, 3, 8])
def test_exchange_batches_match_mirror(native, oracle, k):
    """exchange_every = k (deep ghosts, redundant ghost-row updates, partial last batch) and the
    per-phase exchange (k = 0) produce the mirror's bits."""
    import pyisingmontecarlo_b200 as pkg

    Lx, Ly = 128, 16
    betas = [0.3, 0.44, 0.8, 0.44, 0.5, 0.2, 0.6]
    lat = pkg.SingleLattice2D(Lx, Ly, j=-1.0, seed=9, exchange_every=k)
    assert lat.strip.ghost == max(1, 2 * k)
    lat.sweeps(betas)
    en_ref, st_ref = oracle.msc_mirror_single(Lx, Ly, -1.0, 9, betas, 6, 7)
    assert (lat.local_rows() == st_ref).all()
    assert lat.energy() == en_ref[-1]


def test_deep_strips_reproduce_single_strip(native):
    """Three strips with 4 ghost rows exchanging 4 boundary rows once per 2 sweeps by hand (what
    exchange_deep does over NCCL) == one strip with the per-phase wrap."""
    ctx = native.Context.get(0)
    Lx, Ly, G = 128, 24, 4
    whole = native.Strip(ctx, Lx, Ly, 0, Ly, -1.0, 5)
    bounds = ((0, 6), (6, 16), (16, 24))
    parts = [native.Strip(ctx, Lx, Ly, lo, hi, -1.0, 5, ghost=G) for lo, hi in bounds]
    betas = (0.4, 0.5, 0.3, 0.6)
    for beta in betas:
        for colour in (0, 1):
            whole.wrap_local(1 - colour)
            whole.phase(colour, beta)
    for i in range(0, len(betas), 2):
        bufs = []
        for p in parts:
            b = np.empty((2, 2, G, p.words), dtype=np.uint32)
            p.halo_deep(0, G, b)
            bufs.append(b)
        for k, p in enumerate(parts):
            r = np.empty((2, 2, G, p.words), dtype=np.uint32)
            r[0] = bufs[(k - 1) % 3][1]        # upper neighbour's last rows
            r[1] = bufs[(k + 1) % 3][0]        # lower neighbour's first rows
            p.halo_deep(1, G, r)
        for q in range(4):
            for p in parts:
                p.phase_ext(q & 1, betas[i + q // 2], 3 - q, advance=bool(q & 1), sync=True)
    assert (np.concatenate([p.rows() for p in parts]) == whole.rows()).all()


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


def test_fused_two_colour_sweep_matches_mirror_and_phases(native):
    """The single-pass sweep (both colours, out of place, bands with redundant halo rows) against the
    CPU mirror and the two-phase path.  Opt-in (ISING_STRIP_FUSE=1: it is slower than two launches,
    strip.cu); switched on, and forced on for short bands, in a process of its own."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ({}, {"ISING_STRIP_NO_TMA": "1"}):      # rows staged by TMA / loaded by the warps
        env = dict(os.environ, ISING_STRIP_FUSE="1", ISING_STRIP_FUSE_MIN_ROWS="1", **extra)
        res = subprocess.run([sys.executable, os.path.join(root, "tests", "strip_fused_check.py")],
                             capture_output=True, text=True, timeout=600, env=env)
        assert "STRIP_FUSED_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


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
                         capture_output=True, text=True, timeout=240)
    assert "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
