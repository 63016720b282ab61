"""`ClassicIsing` -- the stateful classical API of the reference (src/classicising.rs:1-180):
the same kernels as `Lattice`, but the experiments live on the device between calls."""
import secrets

import numpy as np

from . import _native as nat
from .lattice import _edges_to_arrays, warn_non_basic_moves

_U64 = 2**64 - 1


class ClassicIsing:
    """Unlike the Lattice class this maintains a set of graphs with internal state."""

    def __init__(self, edges, longitudinal=None, num_experiments=None, seed=None, use_basic_moves=None,
                 *, device=None):
        # classicising.rs:26-60 (the reference unwraps the max index: an empty list panics)
        if len(edges) == 0:
            raise ValueError("Must supply some edges for graph")
        self._a, self._b, self._j = _edges_to_arrays(edges)
        self.nvars = int(max(self._a.max(), self._b.max())) + 1
        self._longitudinal = 0.0 if longitudinal is None else float(longitudinal)
        self._seed = (int(seed) & _U64) if seed is not None else secrets.randbits(64)
        self._use_basic_moves = bool(use_basic_moves) if use_basic_moves is not None else False
        self._edge_importance = False
        # worm moves of a timestep: sites per worm (1..8); see csrc/moves.cu
        self.worm_len = 4
        self._device = device
        ctx = nat.Context.get(device)
        bias = None if self._longitudinal == 0.0 else np.full(self.nvars, self._longitudinal)
        self._graph = nat.Graph.from_edges(ctx, self.nvars, self._a, self._b, self._j, bias)
        # all initial experiments at once: random starts are a pure function of (seed, site,
        # replica), so this equals num_experiments calls of add_graph() without their
        # O(E^2 nvars) state copies (classicising.rs:45-58 builds them in one loop as well)
        self._n = 1 if num_experiments is None else int(num_experiments)
        self._sim = nat.Sim(self._graph, self._n, self._seed) if self._n > 0 else None

    def add_graph(self, initial_state=None, edge_move_importance_sampling=None):
        """classicising.rs:62-79: one more experiment, random start or the given state.
        edge_move_importance_sampling (classicising.rs:75-77) is a property of the object here:
        all experiments share one packed state, so the last value given applies to all of them."""
        if edge_move_importance_sampling is not None:
            self._edge_importance = bool(edge_move_importance_sampling)
        old = None if self._sim is None else self._sim.states()
        self._n += 1
        sim = nat.Sim(self._graph, self._n, self._seed)      # experiment e always owns stream e
        if self._sim is not None:
            # the random streams are keyed by (site, replica, sweep): keep the sweep counter running
            # so that the old experiments do not replay the draws of sweeps 0..T-1
            sim.counter = self._sim.counter
        if old is not None or initial_state is not None:
            st = sim.states()                                   # random start of the new experiment
            if old is not None:
                st[: self._n - 1] = old
            if initial_state is not None:
                init = np.asarray(initial_state, dtype=np.bool_).ravel()
                if len(init) != self.nvars:
                    raise ValueError("initial_state must have nvars entries")
                st[self._n - 1] = init
            sim.set_states(st)
        if self._sim is not None:
            self._sim.close()
        self._sim = sim

    def _set_moves(self, nspinupdates, nedgeupdates, nwormupdates, only_basic_moves):
        """The move counts of do_time_step (classicising.rs:100-106) in units of whole passes:
        nspinupdates in {None, nvars, 0} = one / no colour-class sweep, nedgeupdates = k * nedges
        = k passes of two-spin edge moves, nwormupdates = worm moves per experiment.  With all
        three None and only_basic_moves None / False only the sweep runs, with a warning (D1)."""
        basic = only_basic_moves if only_basic_moves is not None else self._use_basic_moves
        if nspinupdates not in (None, 0, self.nvars):
            raise NotImplementedError("nspinupdates must be None, 0 or nvars: single-spin attempts come as "
                                      "whole colour-class sweeps on the GPU path")
        nedges = len(self._a)
        if nedgeupdates is not None and nedgeupdates % nedges:
            raise NotImplementedError("nedgeupdates must be a multiple of the number of edges: edge moves "
                                      "come as whole passes over the bonds on the GPU path")
        spin = 0 if nspinupdates == 0 else 1
        edge = 0 if nedgeupdates is None else int(nedgeupdates) // nedges
        worms = 0 if nwormupdates is None else int(nwormupdates)
        if basic:
            edge = worms = 0
        elif nedgeupdates is None and nwormupdates is None:
            warn_non_basic_moves(False)
        if self._edge_importance and not edge:
            raise NotImplementedError("edge_move_importance_sampling only affects the edge moves: pass "
                                      "nedgeupdates (a multiple of the number of edges)")
        self._sim.set_moves(spin, edge, worms, self.worm_len, self._edge_importance)

    def run_monte_carlo(self, beta, timesteps, nspinupdates=None, nedgeupdates=None, nwormupdates=None,
                        only_basic_moves=None):
        """classicising.rs:88-110: advances every experiment, returns nothing."""
        self._set_moves(nspinupdates, nedgeupdates, nwormupdates, only_basic_moves)
        self._sim.sweeps(np.full(int(timesteps), float(beta)))

    def run_monte_carlo_sampling(self, beta, timesteps, nspinupdates=None, nedgeupdates=None,
                                 nwormupdates=None, only_basic_moves=None, thermalization_time=None,
                                 sampling_freq=None, *, packed=False):
        """classicising.rs:119-179 -> (energies float64[E, n_s], states bool[E, n_s, nvars]);
        packed=True (additive) returns uint32[n_s, nvars, ceil(E/32)] bit-packed samples instead."""
        self._set_moves(nspinupdates, nedgeupdates, nwormupdates, only_basic_moves)
        thermalization_time = 0 if thermalization_time is None else int(thermalization_time)
        sampling_freq = 1 if sampling_freq is None else int(sampling_freq)
        if sampling_freq == 0:
            raise ZeroDivisionError("sampling_freq must be non-zero (the reference panics)")
        return self._sim.run_sampling(beta, thermalization_time, sampling_freq,
                                      int(timesteps) // sampling_freq, packed=packed)

    def run_monte_carlo_observables(self, beta, timesteps, thermalization_time=None, sampling_freq=None,
                                    overlaps=True):
        """Additive: the sampling loop of classicising.rs:119-179 with the per-sample state copy
        replaced by on-device reductions -> (energies[E, n_s], M[E, n_s], Q[E // 2, n_s])."""
        thermalization_time = 0 if thermalization_time is None else int(thermalization_time)
        sampling_freq = 1 if sampling_freq is None else int(sampling_freq)
        if sampling_freq == 0:
            raise ZeroDivisionError("sampling_freq must be non-zero (the reference panics)")
        return self._sim.run_observables(beta, thermalization_time, sampling_freq,
                                         int(timesteps) // sampling_freq, overlaps)

    # checkpointing (additive: the reference persists only its QMC classes).  seed + sweep counter
    # + packed spins are the whole state, so a restored object continues bit for bit.
    def save_to_file(self, path):
        with open(path, "wb") as f:
            np.savez_compressed(f, kind="ClassicIsing", a=self._a, b=self._b, j=self._j,
                                longitudinal=self._longitudinal, n=self._n,
                                seed=np.uint64(self._seed), sweeps=np.uint64(self._sim.counter),
                                packed=self._sim.packed())

    @staticmethod
    def read_from_file(path, reseed=None, *, device=None):
        """reseed=None resumes the saved random streams exactly; a value starts new ones (the
        reference's read_from_file(path, reseed), tempering.rs:332-338)."""
        with np.load(path) as d:
            if str(d["kind"]) != "ClassicIsing":
                raise IOError(f"{path} is not a ClassicIsing checkpoint")
            edges = [((int(x), int(y)), float(w)) for x, y, w in zip(d["a"], d["b"], d["j"])]
            obj = ClassicIsing(edges, float(d["longitudinal"]), int(d["n"]),
                               int(d["seed"]) if reseed is None else int(reseed), device=device)
            obj._sim.set_packed(d["packed"])
            obj._sim.counter = int(d["sweeps"]) if reseed is None else 0
        return obj

    # additive: the resident state, for tests and users
    def get_states(self):
        return self._sim.states()

    def get_energies(self):
        return self._sim.energies()
