"""B200-native classical Ising Monte Carlo engine behind the `py_monte_carlo.Lattice` API.

Only the data-parallel classical path is here: `Lattice.run_monte_carlo*`,
the annealing runs, replay mode and the parallel-tempering loop.  Compute lives in
`libising_b200.so` (hand-written sm_100a CUDA behind the C ABI of include/ising_b200.h).
"""
from ._native import (AmbiguousReplay, Context, Graph, NativeLibraryMissing, Sim,  # noqa: F401
                      Strip, Tempering)
from .classic import ClassicIsing  # noqa: F401
from .lattice import Lattice  # noqa: F401
from .single_lattice import SingleLattice2D, exchange_halos  # noqa: F401
from .tempering import LatticeTempering, run_tempering_loop, shard_range  # noqa: F401

__all__ = ["Lattice", "ClassicIsing", "LatticeTempering", "SingleLattice2D", "Sim", "Strip", "Tempering", "Graph", "Context",
           "AmbiguousReplay", "NativeLibraryMissing", "run_tempering_loop", "shard_range"]
