"""Drop-in module name of the reference (src/lib.rs:14-22): `import py_monte_carlo`.

The classes on the classical hot path are provided (Lattice, ClassicIsing, and the classical
replica loop of LatticeTempering); QmcRunner / QmcIsing (SSE quantum Monte Carlo) stay on the
reference build and raise NotImplementedError here."""
from pyisingmontecarlo_b200 import ClassicIsing, Lattice, LatticeTempering  # noqa: F401

__all__ = ["Lattice", "ClassicIsing", "LatticeTempering"]


def __getattr__(name):
    if name in ("QmcRunner", "QmcIsing"):
        raise NotImplementedError(
            f"py_monte_carlo.{name}: the SSE quantum Monte Carlo classes (src/qmcrunner.rs, "
            "src/qmcising.rs) are out of scope of the B200 engine and remain on the reference build")
    raise AttributeError(name)
