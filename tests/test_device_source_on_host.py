"""The row-walk sweep kernel's own source, run on the host, against the mirror (CPU suite).

`csrc/sweep_rows.cuh` is the kernel behind BASELINE configs 2 and 3.  Its GPU parity tests need
a B200; this one needs none: tests/host_emulation/ runs the kernel source itself one OS thread
per CUDA thread (thread indices, __syncthreads, shared memory, atomics stand-ins) and the result
is compared bit for bit with oracle/msc_mirror.c, the scalar restatement every GPU parity test
uses.  What is covered here and nowhere else on the CPU: the row geometry of the walk (units,
tiles, wrap-around rows, parities), the Philox rounds shared between the words of a site, the
multiply-add forms of the class select and of the tie compare, the vertical counters of the
accumulating phase with their block reduction, in every instantiation the launcher can choose
(2D / 3D, uniform / +-J, 1 / 2 / 4 words per thread, one row or several per unit, the
256-thread and the 128-thread shape).  The library itself is not involved and stays CUDA-only.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pyisingmontecarlo_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "host_emulation")


def _cut(text, start_marker, end_marker=None, what=""):
    """text without [start_marker, end_marker) (to the end when end_marker is None); both must exist"""
    i = text.find(start_marker)
    assert i >= 0, f"marker not found ({what}): {start_marker!r}"
    if end_marker is None:
        return text[:i]
    k = text.find(end_marker, i)
    assert k >= 0, f"marker not found ({what}): {end_marker!r}"
    return text[:i] + text[k:]


def prepare_sources(dst):
    """Copies of the device headers with what only nvcc can digest taken out.  Every edit is an
    exact, asserted text operation, so a change of the source that the emulation does not follow
    fails here instead of testing something else."""
    os.makedirs(dst, exist_ok=True)
    for name in ("kernels.h", "philox.h"):
        with open(os.path.join(CSRC, name)) as f, open(os.path.join(dst, name), "w") as g:
            g.write(f.read())

    msc = open(os.path.join(CSRC, "msc_device.cuh")).read()
    for line in ("#include <cooperative_groups.h>\n", "namespace cg = cooperative_groups;\n"):
        assert msc.count(line) == 1
        msc = msc.replace(line, "")
    inc = '#include "../../include/ising_b200.h"'
    assert msc.count(inc) == 1
    msc = msc.replace(inc, '#include "%s"' % os.path.join(ROOT, "include", "ising_b200.h"))
    # the launch helpers call cudaLaunchKernelEx
    msc = _cut(msc, "// Launch with programmatic stream serialisation", "static inline uint32_t pow2_ceil",
               "launch helpers of msc_device.cuh")
    assert "__shared__" not in msc
    open(os.path.join(dst, "msc_device.cuh"), "w").write(msc)

    rows = open(os.path.join(CSRC, "sweep_rows.cuh")).read()
    # the TMA-staged variant (opt-in, mbarrier / cp.async.bulk PTX) is not emulated
    tma = ("// ------------------------------------------------------------------------------------------\n"
           "// The same colour phase with the neighbour rows staged in shared memory by the TMA unit")
    rows = _cut(rows, tma, None, "TMA variant of sweep_rows.cuh") + "\n}  // namespace ising\n"
    # programmatic dependent launch orders kernels on a stream; the emulation runs them in order
    for ptx in ('    asm volatile("griddepcontrol.launch_dependents;");\n',
                '            asm volatile("griddepcontrol.wait;" ::: "memory");\n'):
        assert rows.count(ptx) == 1, ptx
        rows = rows.replace(ptx, "")
    assert "asm" not in re.sub(r"//.*", "", rows)
    dyn = "    extern __shared__ uint32_t sm[];\n"
    assert rows.count(dyn) == 2
    rows = rows.replace(dyn, "    uint32_t* sm = emu::dyn_smem;\n")
    assert rows.count("__shared__") == 1          # s_desc, the row geometry of a chunk of units
    rows = rows.replace("__shared__", "EMU_SHARED")
    open(os.path.join(dst, "sweep_rows.cuh"), "w").write(rows)

    launch = open(os.path.join(CSRC, "sweep_rows_launch.cuh")).read()
    launch = _cut(launch, "template <int DIM, bool PMJ, int K, int ROUNDS, int V, bool ACC, bool MULTIROW, bool COUNT,",
                  None, "launchers of sweep_rows_launch.cuh") + "\n}  // namespace ising\n"
    assert "rows_shape" in launch and "rows_partition" in launch
    open(os.path.join(dst, "sweep_rows_launch_shape.cuh"), "w").write(launch)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    build = str(tmp_path_factory.mktemp("host_emulation"))
    prepare_sources(os.path.join(build, "prepared"))
    so = os.path.join(build, "libemu_rows.so")
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", "-w", "-I", EMU, "-I", build,
           "-I", "/usr/local/cuda/include", os.path.join(EMU, "emu_rows.cpp"), "-o", so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    lib = C.CDLL(so)
    lib.emu_rows_phase.restype = C.c_int
    lib.emu_rows_phase.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                                   C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                                   C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int]
    return lib


# ---- the library's stencil layout, restated from csrc/kernels.h (Layout) and api_core.cu ------------
def torus(dims, rng, pmj, j0):
    """-> a, b, j of the periodic lattice, site = x + Lx (y + Ly z), bond d of site n: n -> n + e_d"""
    Lx, Ly, Lz = dims
    z, y, x = np.meshgrid(np.arange(Lz), np.arange(Ly), np.arange(Lx), indexing="ij")
    n = (x + Lx * (y + Ly * z)).ravel()
    fwd = [((x + 1) % Lx + Lx * (y + Ly * z)).ravel(), (x + Lx * ((y + 1) % Ly + Ly * z)).ravel()]
    if Lz > 1:
        fwd.append((x + Lx * (y + Ly * ((z + 1) % Lz))).ravel())
    a = np.concatenate([n] * len(fwd)).astype(np.uint64)
    b = np.concatenate(fwd).astype(np.uint64)
    j = np.full(len(a), j0)
    if pmj:
        j = j * rng.choice([-1.0, 1.0], size=len(a))
    return a, b, j


def layout_index(dims):
    """word row of natural site n in spins[2][rows][Lxh] (the replica words follow)"""
    Lx, Ly, Lz = dims
    n = np.arange(Lx * Ly * Lz)
    x, r = n % Lx, n // Lx
    y, z = r % Ly, r // Ly
    c = (x + y + z) & 1
    return (c * (Ly * Lz) + r) * (Lx // 2) + (x >> 1), c


def pack(states, dims, W):
    E, N = states.shape
    bits = np.zeros((N, W * 32), dtype=np.uint8)
    bits[:, :E] = states.T
    words = (bits.reshape(N, W, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(axis=2, dtype=np.uint32)
    idx, _ = layout_index(dims)
    out = np.zeros((N, W), dtype=np.uint32)
    out[idx] = words
    return out


def unpack(words, dims, E):
    idx, _ = layout_index(dims)
    w = words[idx]                                                    # natural order [N, W]
    bits = (w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1
    return bits.reshape(len(idx), -1)[:, :E].T.astype(bool)


def bond_masks(dims, a, b, j):
    """[2][halfN][8] as upload_stencil_masks (api_core.cu) lays them out: k = 0 the x neighbour stored
    at the same half-index, 1 the other x neighbour, 2 y-1, 3 y+1, 4 z-1, 5 z+1; all-ones iff J > 0"""
    Lx, Ly, Lz = dims
    N = Lx * Ly * Lz
    ndir = 3 if Lz > 1 else 2
    anti = (j > 0).reshape(ndir, N)                                   # anti[d, n]: bond n -> n + e_d
    idx, c = layout_index(dims)
    n = np.arange(N)
    x, r = n % Lx, n // Lx
    y, z = r % Ly, r // Ly
    p = x & 1                                                         # = (y + z + c) & 1
    xm = ((x - 1) % Lx) + Lx * r
    ym = x + Lx * ((y - 1) % Ly + Ly * z)
    zm = x + Lx * (y + Ly * ((z - 1) % Lz))
    m = np.zeros((N, 8), dtype=np.uint32)
    jxp, jxm = anti[0, n], anti[0, xm]
    m[:, 0] = np.where(p == 0, jxp, jxm)
    m[:, 1] = np.where(p == 0, jxm, jxp)
    m[:, 2] = anti[1, ym]
    m[:, 3] = anti[1, n]
    if ndir == 3:
        m[:, 4] = anti[2, zm]
        m[:, 5] = anti[2, n]
    out = np.zeros((N, 8), dtype=np.uint32)
    out[idx] = m * np.uint32(0xFFFFFFFF)
    return out


def satisfied_bonds(states, a, b, j):
    s = states.astype(np.int64) * 2 - 1
    return ((s[:, a.astype(np.int64)] * s[:, b.astype(np.int64)] * np.sign(j)[None, :]) < 0).sum(axis=1)


CASES = [
    # dims, E, V, pmj, small, grid blocks, betas
    ((8, 6, 4), 128, 4, True, False, 5, [0.1, 0.5, 1.2]),      # config 3's instantiation (V = 4, +-J, 3D)
    ((8, 6, 4), 128, 4, True, True, 3, [0.3, 0.9]),            # ... in the 128-thread shape (the 8-GPU split)
    ((4, 4, 6), 96, 1, True, False, 7, [0.2, 0.7]),            # ragged last word (96 = 3 words), V = 1
    ((6, 4, 2), 64, 2, False, False, 2, [0.25, 0.6]),          # uniform ferromagnet, V = 2, Lz = 2 (z-1 == z+1)
    ((16, 8, 1), 128, 4, False, False, 4, [0.3, 0.44, 0.6]),   # config 2's instantiation (2D, uniform)
    ((12, 6, 1), 40, 2, True, False, 3, [0.4, 1.0]),           # 2D +-J, E not a multiple of 32
    ((600, 2, 1), 32, 1, False, False, 6, [0.44]),             # several x tiles per row, single row per unit
]


@pytest.mark.parametrize("dims,E,V,pmj,small,grid,betas", CASES)
def test_row_walk_source_equals_the_mirror(emu, oracle, dims, E, V, pmj, small, grid, betas):
    rng = np.random.default_rng(hash((dims, E, V, pmj)) & 0xFFFF)
    dim = 3 if dims[2] > 1 else 2
    j0 = -1.0 if pmj or dims[0] != 6 else 1.0                    # one uniform case is antiferromagnetic
    a, b, j = torus(dims, rng, pmj, j0)
    N = dims[0] * dims[1] * dims[2]
    _, colors = layout_index(dims)
    W = (E + 31) // 32
    if W % V:
        W += V - W % V                                            # the launcher picks V from W; pad words are idle replicas
    init = rng.integers(0, 2, size=(E, N)).astype(bool)
    seed, sweep0, gw0 = 0x1234567887654321, 5, 3
    words = pack(init, dims, W)
    jm8 = bond_masks(dims, a, b, j) if pmj else None
    antiferro = 0xFFFFFFFF if (not pmj and j0 > 0) else 0
    nsat_per_sweep = []
    for s, beta in enumerate(betas):
        for colour in (0, 1):
            nsat = np.zeros(W * 32, dtype=np.uint64)
            rc = emu.emu_rows_phase(dim, dims[0], dims[1], dims[2], W, V,
                                    None if jm8 is None else jm8.ctypes.data, antiferro, words.ctypes.data,
                                    colour, seed, sweep0 + s, gw0, float(beta), 1.0, int(colour == 1),
                                    nsat.ctypes.data, int(small), grid)
            assert rc == 0, rc
        nsat_per_sweep.append(nsat[:E].copy())
    got = unpack(words, dims, E)
    en_ref, st_ref = oracle.msc_mirror(a, b, j, N, colors, E, seed, betas, replica_offset=32 * gw0,
                                       states=init, sweep0=sweep0, per_sweep=True)
    assert (got == st_ref).all(), "kernel source on the host differs from the mirror"
    assert (got != init).mean() > 0.05                             # (the sweeps did move the spins)
    # the accumulating phase: satisfied bonds after every sweep <-> the mirror's per-sweep energies
    nb = len(a)
    en = np.array([(nb - 2.0 * n.astype(np.float64)) for n in nsat_per_sweep]).T
    assert (en == en_ref).all()
    assert (nsat_per_sweep[-1] == satisfied_bonds(got, a, b, j)).all()
    # and the count-only pass (get_energy of the current configuration)
    nsat = np.zeros(W * 32, dtype=np.uint64)
    rc = emu.emu_rows_phase(dim, dims[0], dims[1], dims[2], W, V, None if jm8 is None else jm8.ctypes.data,
                            antiferro, words.ctypes.data, 0, seed, 0, gw0, 0.0, 1.0, 2, nsat.ctypes.data, 0, grid)
    assert rc == 0, rc
    assert (nsat[:E] == satisfied_bonds(got, a, b, j)).all()
    assert (unpack(words, dims, E) == got).all()                  # count only: no update
