"""Drop-in module name of the reference (src/lib.rs:14-22): `import py_monte_carlo`.

Only the classes on the classical hot path are provided (Lattice, and the classical replica
loop of LatticeTempering); QmcRunner / QmcIsing stay on the reference build."""
from pyisingmontecarlo_b200 import ClassicIsing, Lattice, LatticeTempering  # noqa: F401

__all__ = ["Lattice", "ClassicIsing", "LatticeTempering"]
