"""Halo-exchange plumbing of the strip decomposition over torch.distributed/gloo (world_size 2,
CPU): a numpy stand-in for the device strip must give the same lattice as one whole strip."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeStrip:
    """Same interface as _native.Strip; the 'update' is a deterministic function of a site's
    four neighbours, the global row and the sweep, so it exercises exactly the data the halo
    exchange must deliver."""

    def __init__(self, Lx, Ly, lo, hi):
        self.Lx, self.Ly, self.lo, self.hi = Lx, Ly, lo, hi
        self.words = Lx // 64
        rng = np.random.default_rng(1)
        full = rng.integers(0, 2**32, size=(2, Ly, self.words), dtype=np.uint64).astype(np.uint32)
        self.s = np.zeros((2, hi - lo + 2, self.words), dtype=np.uint32)
        self.s[:, 1:-1] = full[:, lo:hi]
        self.t = 0

    def get_boundary(self, colour, which, dst=None):
        return self.s[colour, -2 if which else 1].copy()

    def set_ghost(self, colour, which, src):
        self.s[colour, -1 if which else 0] = np.asarray(src, dtype=np.uint32)

    def wrap_local(self, colour):
        self.s[colour, 0] = self.s[colour, -2]
        self.s[colour, -1] = self.s[colour, 1]

    def phase(self, colour, beta):
        o = 1 - colour
        rows = np.arange(self.lo, self.hi, dtype=np.uint32)[:, None]
        cur = self.s[colour, 1:-1]
        mix = (self.s[o, :-2] ^ np.roll(self.s[o, 1:-1], 1, axis=1)) + self.s[o, 2:] * np.uint32(3) \
            + self.s[o, 1:-1] * np.uint32(5) + rows * np.uint32(2654435761) + np.uint32(self.t)
        self.s[colour, 1:-1] = cur ^ mix.astype(np.uint32)
        if colour == 1:
            self.t += 1

    def interior(self):
        return self.s[:, 1:-1].copy()


class DeepFakeStrip:
    """FakeStrip with `ghost` ghost rows per side and the deep-halo interface of _native.Strip
    (halo_deep / wrap_deep / phase_ext); same update rule, keyed by the global row."""

    def __init__(self, Lx, Ly, lo, hi, ghost):
        self.Lx, self.Ly, self.lo, self.hi, self.g = Lx, Ly, lo, hi, ghost
        self.words = Lx // 64
        rng = np.random.default_rng(1)
        full = rng.integers(0, 2**32, size=(2, Ly, self.words), dtype=np.uint64).astype(np.uint32)
        self.s = np.zeros((2, hi - lo + 2 * ghost, self.words), dtype=np.uint32)
        self.s[:, ghost:ghost + hi - lo] = full[:, lo:hi]
        self.t = 0

    def halo_deep(self, direction, depth, buf, sync=True):
        g, n = self.g, self.hi - self.lo
        b = buf.reshape(2, 2, depth, self.words)
        if direction == 0:
            b[0] = self.s[:, g:g + depth]
            b[1] = self.s[:, g + n - depth:g + n]
        else:
            self.s[:, g - depth:g] = b[0]
            self.s[:, g + n:g + n + depth] = b[1]

    def wrap_deep(self, depth):
        g, n = self.g, self.hi - self.lo
        self.s[:, g - depth:g] = self.s[:, g + n - depth:g + n]
        self.s[:, g + n:g + n + depth] = self.s[:, g:g + depth]

    def phase_ext(self, colour, beta, ext, advance=False, sync=False):
        o, g, n = 1 - colour, self.g, self.hi - self.lo
        a, b = g - ext, g + n + ext                      # storage rows updated
        rows = ((np.arange(a, b) - g + self.lo) % self.Ly).astype(np.uint32)[:, None]
        cur = self.s[colour, a:b]
        mix = (self.s[o, a - 1:b - 1] ^ np.roll(self.s[o, a:b], 1, axis=1)) + self.s[o, a + 1:b + 1] * np.uint32(3) \
            + self.s[o, a:b] * np.uint32(5) + rows * np.uint32(2654435761) + np.uint32(self.t)
        self.s[colour, a:b] = cur ^ mix.astype(np.uint32)
        if advance:
            self.t += 1

    def interior(self):
        return self.s[:, self.g:self.g + self.hi - self.lo].copy()


def _run_deep(strip, rank, world, k, dist=None):
    from pyisingmontecarlo_b200.single_lattice import sweep_batches
    import torch

    # 5 sweeps in batches of k (the last one shorter): the loop SingleLattice2D.sweeps runs
    sweep_batches(strip, [0.4] * 5, k, rank, world, dist, None, torch.device("cpu"))
    return strip.interior()


def _worker_deep(rank, world, port, out, k):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from pyisingmontecarlo_b200.tempering import shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(20, rank, world)
    res = _run_deep(DeepFakeStrip(128, 20, lo, hi, 2 * k), rank, world, k, dist)
    np.save(out + f".{rank}.npy", res)
    dist.destroy_process_group()


def _run(strip, rank, world, dist=None):
    from pyisingmontecarlo_b200.single_lattice import exchange_halos
    import torch

    for _ in range(5):
        for colour in (0, 1):
            exchange_halos(strip, 1 - colour, rank, world, dist, None, torch.device("cpu"))
            strip.phase(colour, 0.4)
    return strip.interior()


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from pyisingmontecarlo_b200.tempering import shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(20, rank, world)
    res = _run(FakeStrip(128, 20, lo, hi), rank, world, dist)
    np.save(out + f".{rank}.npy", res)
    dist.destroy_process_group()


def test_halo_exchange_over_gloo_equals_whole_lattice(native, tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "strip")
    for world in (2, 3):
        mp.spawn(_worker, args=(world, port + world, out), nprocs=world, join=True)
        whole = _run(FakeStrip(128, 20, 0, 20), 0, 1)
        got = np.concatenate([np.load(out + f".{r}.npy") for r in range(world)], axis=1)
        assert (got == whole).all()


def test_deep_halo_batches_over_gloo_equal_per_phase_exchange(native, tmp_path):
    """Communication-avoiding batches (one exchange of 2k rows per k sweeps, ghost rows updated
    redundantly) give the same lattice as one exchange per colour phase."""
    import torch.multiprocessing as mp

    whole = _run(FakeStrip(128, 20, 0, 20), 0, 1)
    for k in (1, 2, 3):
        assert (_run_deep(DeepFakeStrip(128, 20, 0, 20, 2 * k), 0, 1, k) == whole).all()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "deep")
    for world, k in ((2, 2), (3, 3)):
        mp.spawn(_worker_deep, args=(world, port + world, out, k), nprocs=world, join=True)
        got = np.concatenate([np.load(out + f".{r}.npy") for r in range(world)], axis=1)
        assert (got == whole).all()
