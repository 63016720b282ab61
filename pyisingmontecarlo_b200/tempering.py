"""Classical parallel tempering with the replica loop of the reference's `LatticeTempering`
(src/tempering.rs:43-113, 156-222, 297-299).

The reference class drives SSE-QMC replicas (transverse field > 0); that stepper stays on the
reference.  What is rebuilt here is the loop the north star names -- run / swap / sample with
the same cadence and return layout -- around a classical stepper: every temperature is one
replica bit of the packed device layout, so the whole ladder advances in one sweep, and a swap
exchanges two betas instead of moving configurations.

Multi-GPU (one process per GPU): rank r owns a contiguous block of configurations; per swap step
there is one all-gather of R energies, then every rank takes the same Philox-keyed decisions.
On GPUs the whole cycle - energies, NCCL all-gather, swap kernel, threshold tables, samples - is
enqueued by the C library on its stream (ising_pt_set_comm + ising_pt_timesteps_sample);
torch.distributed only carries the 128-byte NCCL id at start-up.  `run_tempering_loop` is the
same loop in host code over any torch.distributed backend (gloo in the CPU tests).
"""
import secrets

import numpy as np

from . import _native as nat
from .lattice import _edges_to_arrays

_U64 = 2**64 - 1


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units owned by `rank` (the reference's rayon axis)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Collective:
    """all-gather of small host arrays over torch.distributed (NCCL on GPUs, gloo on CPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1

    def allgather_concat(self, arr, counts):
        """Concatenation over ranks of 1-D/2-D arrays whose leading sizes are `counts`."""
        if not self.active:
            return arr
        import torch

        backend = self.dist.get_backend(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        width = int(np.prod(arr.shape[1:])) if arr.ndim > 1 else 1
        cmax = max(max(counts), 1)
        buf = np.zeros((cmax, width), dtype=arr.dtype)
        buf[: arr.shape[0]] = arr.reshape(arr.shape[0], width)
        t = torch.from_numpy(buf.view(np.uint8) if arr.dtype == np.bool_ else buf).to(dev)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.group)
        parts = []
        for r, o in enumerate(outs):
            a = o.cpu().numpy()
            if arr.dtype == np.bool_:
                a = a.view(np.bool_)
            parts.append(a[: counts[r]])
        out = np.concatenate(parts, axis=0)
        return out.reshape((out.shape[0],) + arr.shape[1:])


def run_tempering_loop(stepper, nbetas, nvars, timesteps, replica_swap_freq, sampling_freq,
                       collective=None, counts=None):
    """The loop of qmc_timesteps_sample (tempering.rs:156-222) over a (possibly sharded) stepper.

    stepper: .sweeps(t) -> local energies by configuration; .swap_step(all_energies);
             .slots() -> slot_of_config[nbetas]; .local_states() -> bool[local, nvars]
    Returns (states bool[R, timesteps // sampling_freq, nvars], energies float64[R]); states[r]
    is the configuration currently at beta_r, the swap happens before the sample when both fall
    on the same step.
    """
    if replica_swap_freq <= 0 or sampling_freq <= 0:
        raise ValueError("replica_swap_freq and sampling_freq must be > 0 "
                         "(the reference's loop never terminates on 0)")
    coll = collective or _Collective()
    counts = counts or [nbetas]
    ns = timesteps // sampling_freq
    states = np.zeros((nbetas, ns, nvars), dtype=np.bool_)
    acc = np.zeros(nbetas)
    remaining, to_swap, to_sample, k = timesteps, replica_swap_freq, sampling_freq, 0
    while remaining > 0:
        t = min(to_sample, to_swap, remaining)
        local_e = stepper.sweeps(t)
        all_e = coll.allgather_concat(np.ascontiguousarray(local_e, dtype=np.float64), counts)
        slots = stepper.slots().astype(np.int64)          # slot of every configuration
        acc[slots] += all_e * t
        to_sample -= t
        to_swap -= t
        remaining -= t
        if to_swap == 0:
            stepper.swap_step(all_e)
            to_swap = replica_swap_freq
        if to_sample == 0:
            if k < ns:
                all_s = coll.allgather_concat(stepper.local_states(), counts)
                states[stepper.slots().astype(np.int64), k] = all_s
            k += 1
            to_sample = sampling_freq
    return states, acc / timesteps


class LatticeTempering:
    """Replica container with the reference's surface (tempering.rs:43-113): add_graph(..., beta),
    qmc_timesteps, qmc_timesteps_sample, get_total_swaps, get_graph_itime, clone, save_to_file /
    read_from_file.  Classical replicas only."""

    def __init__(self, edges, seed=None, use_allocator=None, *, device=None, process_group=None):
        if len(edges) == 0:
            raise ValueError("Must supply some edges for graph")
        self._a, self._b, self._j = _edges_to_arrays(edges)
        self.nvars = int(max(self._a.max(), self._b.max())) + 1
        self._seed = None if seed is None else int(seed) & _U64
        self._device = device
        self._group = process_group
        self._betas = []
        self._longitudinal = None
        self._pt = None
        self._graph = None

    # tempering.rs:70-113
    def add_graph(self, transverse, longitudinal, beta, edges=None, enable_rvb_update=None,
                  enable_heatbath_update=None, seed=None, use_allocator=None):
        if transverse != 0.0:
            raise NotImplementedError("SSE quantum replicas (transverse field > 0) remain on the "
                                      "reference; the B200 engine tempers classical replicas")
        if edges is not None:
            raise NotImplementedError("per-replica edge lists are not supported: replicas share the lattice")
        if self._pt is not None:
            raise RuntimeError("add_graph after the first run is not supported")
        # replicas of a ladder differ in beta only: one longitudinal field for all of them (a field or
        # couplings of unequal size put the ladder on the float-field kernels, like Lattice)
        if self._longitudinal is not None and float(longitudinal) != self._longitudinal:
            raise NotImplementedError("replicas of one ladder share the longitudinal field: got "
                                      f"{longitudinal} after {self._longitudinal}")
        self._longitudinal = float(longitudinal)
        self._betas.append(float(beta))

    def get_num_graphs(self):
        return len(self._betas)

    def _ensure(self):
        if self._pt is None:
            if not self._betas:
                raise ValueError("no replicas: call add_graph first")
            ctx = nat.Context.get(self._device)
            bias = np.full(self.nvars, self._longitudinal) if self._longitudinal else None
            self._graph = nat.Graph.from_edges(ctx, self.nvars, self._a, self._b, self._j, bias)
            self._coll = _Collective(self._group)
            R = len(self._betas)
            self._counts = [shard_range(R, r, self._coll.world)[1] - shard_range(R, r, self._coll.world)[0]
                            for r in range(self._coll.world)]
            lo, hi = shard_range(R, self._coll.rank, self._coll.world)
            if hi <= lo:
                raise ValueError("fewer replicas than ranks")
            seed = self._seed if self._seed is not None else secrets.randbits(64)
            if self._coll.active and self._seed is None:
                raise ValueError("a seed is required when tempering across ranks")
            self._pt = nat.Tempering(self._graph, self._betas, seed, lo, hi)
            self._comm = None
            if self._coll.active and self._coll.dist.get_backend(self._group) == "nccl":
                # the collectives of the swap cycle are the library's own NCCL calls
                self._comm = nat.Comm.from_torch(ctx, self._group)
                self._pt.set_comm(self._comm)
        return self._pt

    def qmc_timesteps(self, t):
        """tempering.rs:150-152: t timesteps on every replica, no swaps."""
        self._ensure().sweeps(int(t), want_energies=False)

    def qmc_timesteps_sample(self, timesteps, replica_swap_freq=None, sampling_freq=None):
        """tempering.rs:156-222 -> (states bool[R, timesteps // sampling_freq, nvars],
        energies float64[R]) -- states first, as in the reference."""
        sampling_freq = 1 if sampling_freq is None else int(sampling_freq)
        replica_swap_freq = 1 if replica_swap_freq is None else int(replica_swap_freq)
        pt = self._ensure()
        if not self._coll.active or self._comm is not None:
            if replica_swap_freq <= 0 or sampling_freq <= 0:
                raise ValueError("replica_swap_freq and sampling_freq must be > 0 "
                                 "(the reference's loop never terminates on 0)")
            return pt.timesteps_sample(int(timesteps), replica_swap_freq, sampling_freq)
        return run_tempering_loop(pt, len(self._betas), self.nvars, int(timesteps), replica_swap_freq,
                                  sampling_freq, self._coll, self._counts)

    timesteps = qmc_timesteps
    timesteps_sample = qmc_timesteps_sample

    # tempering.rs:307-347 (save_to_file / read_from_file) for the classical ladder.  Unlike the
    # reference ("Does not save state of RNG") a restored ladder continues bit for bit: the RNG
    # is counter-based, so seed + counters are its whole state.  A ladder that is sharded over ranks
    # writes one file per rank (path + ".rank<r>of<n>": the rank's configurations, the shared slot
    # permutation and counters) and is read back by the same number of ranks.
    @staticmethod
    def _rank_path(path, coll):
        return f"{path}.rank{coll.rank}of{coll.world}" if coll.active else path

    def save_to_file(self, path):
        pt = self._ensure()
        ck = pt.checkpoint()
        with open(self._rank_path(path, self._coll), "wb") as f:
            np.savez_compressed(f, kind="LatticeTempering", a=self._a, b=self._b, j=self._j,
                                betas=np.asarray(self._betas), seed=np.uint64(pt.seed),
                                longitudinal=np.float64(self._longitudinal or 0.0),
                                world=self._coll.world, rank=self._coll.rank, **ck)

    @staticmethod
    def read_from_file(path, reseed=None, *, device=None, process_group=None):
        coll = _Collective(process_group)
        with np.load(LatticeTempering._rank_path(path, coll)) as d:
            if str(d["kind"]) != "LatticeTempering":
                raise IOError(f"{path} is not a LatticeTempering checkpoint")
            if int(d["world"]) != coll.world or int(d["rank"]) != coll.rank:
                raise IOError(f"{path} was written by rank {int(d['rank'])} of {int(d['world'])}, "
                              f"this is rank {coll.rank} of {coll.world}")
            edges = [((int(x), int(y)), float(w)) for x, y, w in zip(d["a"], d["b"], d["j"])]
            obj = LatticeTempering(edges, int(d["seed"]) if reseed is None else int(reseed), device=device,
                                   process_group=process_group)
            longitudinal = float(d["longitudinal"]) if "longitudinal" in d.files else 0.0
            for b in d["betas"]:
                obj.add_graph(0.0, longitudinal, float(b))
            pt = obj._ensure()
            ck = {k: d[k] for k in ("packed", "sweeps", "slots", "swap_step", "total_swaps")}
            if reseed is not None:
                ck["sweeps"], ck["swap_step"] = 0, 0
            pt.restore(ck)
        return obj

    def get_total_swaps(self):
        return 0 if self._pt is None else self._pt.total_swaps()

    def clone(self):
        """tempering.rs:302-304: a copy of the container and of all its graphs.  The copy owns its
        own device state (configurations, slot permutation, sweep and swap counters); the RNG is
        counter-based, so original and copy continue identically until they are driven differently.
        A ladder that is sharded over ranks is cloned rank by rank (a collective-free call)."""
        edges = [((int(x), int(y)), float(w)) for x, y, w in zip(self._a, self._b, self._j)]
        obj = LatticeTempering(edges, self._seed, device=self._device, process_group=self._group)
        obj._betas = list(self._betas)
        obj._longitudinal = self._longitudinal
        if self._pt is not None:
            obj._seed = int(self._pt.seed)
            obj._ensure().restore(self._pt.checkpoint())
        return obj

    def get_graph_itime(self, g):
        """tempering.rs:119-148 returns the imaginary-time slices bool[cutoff, nvars] of QMC graph g;
        a classical replica has one slice: bool[1, nvars], the configuration currently at beta_g."""
        n = len(self._betas)
        if not 0 <= int(g) < n:
            raise ValueError(f"Attempted to get graph {g} of {n}")
        pt = self._ensure()
        if self._coll.active:
            states = self._coll.allgather_concat(pt.local_states(), self._counts)
        else:
            states = pt.local_states()
        cfg = int(np.argsort(pt.slots().astype(np.int64))[int(g)])   # configuration at slot g
        return np.ascontiguousarray(states[cfg][None, :])

    def __getattr__(self, name):
        if name.startswith("run_quantum_monte_carlo"):
            raise NotImplementedError(f"{name}: only the classical replica loop is provided here")
        raise AttributeError(name)
