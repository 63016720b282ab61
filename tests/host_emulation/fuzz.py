"""Random shapes through the sweep kernels' source on the host (tests/host_emulation/) against the mirror.

Not part of the test suite (tests/test_device_source_on_host.py holds the fixed cases): a development
tool for kernel changes.  `python tests/host_emulation/fuzz.py [cases per kernel family]` builds the
emulation once under a temporary directory and draws lattice sizes, replica counts, block counts,
plane / round counts, uniform and per-replica betas at random for the row walk, the per-phase /
cooperative / cluster checkerboard kernels, the general-graph kernel and the strip kernel; every case
must equal oracle/msc_mirror.c bit for bit (or be declined the way the library's launcher declines it).
Last run of the committed kernels: 500 + 300 + 150 + 150 cases, no mismatch.

EMU_SANITIZE=address (or thread) with the sanitizer runtime preloaded builds the emulation under that
sanitizer: `ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) EMU_SANITIZE=address
python tests/host_emulation/fuzz.py 80` is a memcheck over random shapes (last run: 320 cases, no report)."""
import ctypes as C
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))                        # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))       # repo root

import numpy as np

import oracle_lib
import test_device_source_on_host as T


import tempfile
build = tempfile.mkdtemp(prefix='ising_fuzz_')
NCASES = int(sys.argv[1]) if len(sys.argv) > 1 else 60
T.prepare_sources(os.path.join(build, 'prepared'))
so = os.path.join(build, 'libemu.so')
if True:
    san = ["-fsanitize=" + os.environ["EMU_SANITIZE"], "-g"] if os.environ.get("EMU_SANITIZE") else []   # with LD_PRELOAD of the runtime
    flags = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-pthread", "-w", "-I", T.EMU, "-I", build, "-I", "/usr/local/cuda/include"] + san
    units = ["emu_rows", "emu_stencil", "emu_general", "emu_strip"]
    ps = [subprocess.Popen(flags + ["-c", os.path.join(T.EMU, u + ".cpp"), "-o", os.path.join(build, u + ".o")]) for u in units]
    for p in ps: assert p.wait() == 0
    assert subprocess.run(["g++", "-shared", "-pthread"] + san[:1] + ["-o", so] + [os.path.join(build, u + ".o") for u in units]).returncode == 0
lib = C.CDLL(so)
lib.emu_rows_phase.restype = C.c_int
lib.emu_rows_phase.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                               C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32,
                               C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int]

def run(dims, E, pmj, small, grid, seed_i):
    rng = np.random.default_rng(seed_i)
    dim = 3 if dims[2] > 1 else 2
    a, b, j = T.torus(dims, rng, pmj, -1.0)
    N = dims[0]*dims[1]*dims[2]
    _, colors = T.layout_index(dims)
    W = (E + 31)//32
    V = 4 if W % 4 == 0 else (2 if W % 2 == 0 else 1)
    if small and V != 4: return 'skip'
    init = rng.integers(0, 2, size=(E, N)).astype(bool)
    words = T.pack(init, dims, W)
    jm8 = T.bond_masks(dims, a, b, j) if pmj else None
    betas = [0.3, 0.9]
    seed = 0x1234567887654321 + seed_i
    ens = []
    for s, beta in enumerate(betas):
        for colour in (0, 1):
            nsat = np.zeros(W*32, dtype=np.uint64)
            rc = lib.emu_rows_phase(dim, dims[0], dims[1], dims[2], W, V, None if jm8 is None else jm8.ctypes.data, 0,
                                    words.ctypes.data, colour, seed, 7 + s, 1, float(beta), 1.0, int(colour == 1),
                                    nsat.ctypes.data, int(small), grid)
            if rc != 0: return f'rc{rc}'
        ens.append(nsat[:E].copy())
    got = T.unpack(words, dims, E)
    en_ref, st_ref = oracle_lib.msc_mirror(a, b, j, N, colors, E, seed, betas, replica_offset=32, states=init, sweep0=7, per_sweep=True)
    ok = (got == st_ref).all() and (np.array([len(a) - 2.0*n for n in ens]).T == en_ref).all()
    return 'ok' if ok else 'MISMATCH'

shapes = [(2,2,1),(2,4,1),(4,2,1),(6,2,1),(2,2,2),(2,4,2),(4,2,4),(10,2,2),(14,6,1),(18,10,1),(34,4,1),(66,2,1),(130,2,1),
          (258,2,1),(514,2,1),(1030,2,1),(4,6,6),(6,10,2),(12,12,1),(8,14,2),(20,2,6),(6,6,6),(4,18,2)]
Es = [1, 32, 33, 64, 96, 128, 160, 256, 1300, 4100]
rng = np.random.default_rng(0)
t0 = time.time(); n = 0; bad = []
for i in range(NCASES):
    dims = shapes[rng.integers(len(shapes))]
    E = Es[rng.integers(len(Es))]
    if dims[0]*dims[1]*dims[2]*E > 300000: E = 64
    pmj = bool(rng.integers(2)); small = bool(rng.integers(2)); grid = int(rng.integers(1, 9))
    r = run(dims, E, pmj, small, grid, i)
    n += 1
    if r not in ('ok', 'skip', 'rc-2'): bad.append((dims, E, pmj, small, grid, r))
print('row walk:', n, 'cases, bad:', bad, '%.0f s' % (time.time() - t0), flush=True)
all_bad = list(bad)

# ---- per-phase / cooperative / cluster checkerboard kernels ----
lib.emu_stencil.restype = C.c_int
lib.emu_stencil.argtypes = [C.c_int, C.c_int] + [C.c_uint32] * 4 + [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p,
                            C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                            C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
# available (dim, pmj, V)
combos = [(3, True, 4), (3, True, 1), (3, False, 2), (2, False, 4), (2, False, 1), (2, True, 2)]
shapes3 = [(2,2,2),(4,2,2),(2,4,4),(4,4,4),(6,4,2),(8,2,6),(10,6,2),(16,4,2),(6,6,6),(34,2,2)]
shapes2 = [(2,2,1),(4,2,1),(2,6,1),(8,8,1),(32,32,1),(18,4,1),(66,2,1),(12,10,1),(130,4,1),(6,14,1)]
rng = np.random.default_rng(1)
bad = []; t0 = time.time(); cnt = {}
for i in range(NCASES):
    dim, pmj, V = combos[rng.integers(len(combos))]
    dims = (shapes3 if dim == 3 else shapes2)[rng.integers(10)]
    W = V * int(rng.integers(1, 4)) if V < 4 else 4 * int(rng.integers(1, 3))
    if V == 1 and W % 2 == 0: W += 1
    if V == 2 and W % 4 == 0: W += 2
    E = W * 32 - int(rng.integers(0, 32))
    mode = ['phase', 'coop', 'cluster'][rng.integers(3)]
    perbeta = bool(rng.integers(2)) and mode != 'coop'
    energies = bool(rng.integers(2))
    K, rounds = (6, 7) if (mode != 'phase' or perbeta) else [(6, 7), (5, 7), (7, 10)][rng.integers(3)]
    units = int(rng.integers(1, 9)) if mode != 'cluster' else [1, 2, 4, 8, 16][rng.integers(5)]
    N = dims[0]*dims[1]*dims[2]
    if mode == 'cluster' and N // 2 * W > 16384: continue
    a, b, j = T.torus(dims, rng, pmj, -1.0)
    _, colors = T.layout_index(dims)
    init = rng.integers(0, 2, size=(E, N)).astype(bool)
    words = T.pack(init, dims, W)
    jmask = None
    if pmj:
        jmask = np.ascontiguousarray(T.bond_masks(dims, a, b, j).reshape(2, N // 2, 8)[:, :, :2 * dim].transpose(0, 2, 1))
    seed, sweep0, gw0, cw, nsw = 0xFEEDFACE12345 + i, 9, 2, W * 32, 3
    if perbeta:
        betas_e = np.geomspace(0.1, 1.4, E)
        tplane, tlow = T.stencil_tables(betas_e, W, dim, 1.0)
        tp, tl = tplane.ctypes.data, tlow.ctypes.data
        betas = np.zeros(nsw)
    else:
        tp = tl = None
        betas = np.array([0.2, 0.5, 1.0])
    jm = None if jmask is None else jmask.ctypes.data
    def call(m, colour, sweep, n, bts, acc, hist):
        return lib.emu_stencil(m, dim, dims[0], dims[1], dims[2], W, V, jm, 0, words.ctypes.data, colour, seed, sweep, n,
                               gw0, K, rounds, bts.ctypes.data, 1.0, tp, tl, int(acc), hist.ctypes.data, cw, units)
    rc = 0
    if mode == 'phase':
        hist = np.zeros((nsw, cw), dtype=np.uint64)
        for t in range(nsw):
            for colour in (0, 1):
                rc |= call(0, colour, sweep0 + t, 1, betas[t:t+1], energies and colour == 1, hist[t])
    else:
        hist = np.zeros((1 if perbeta else nsw, cw), dtype=np.uint64)
        rc = call(1 if mode == 'coop' else 2, 0, sweep0, nsw, betas, energies, hist)
    tag = (mode, dims, E, V, pmj, K, rounds, perbeta, energies, units)
    if rc == 1:
        cnt['declined'] = cnt.get('declined', 0) + 1; continue
    if rc != 0:
        bad.append((tag, 'rc', rc)); print(tag, 'rc', rc); continue
    got = T.unpack(words, dims, E)
    kw = dict(replica_offset=32 * gw0, planes=K, rounds=rounds, states=init, sweep0=sweep0, per_sweep=True)
    if perbeta:
        en_ref, st_ref = oracle_lib.msc_mirror(a, b, j, N, colors, E, seed, None, per_replica_beta=betas_e, nsweeps=nsw, **kw)
    else:
        en_ref, st_ref = oracle_lib.msc_mirror(a, b, j, N, colors, E, seed, betas, **kw)
    ok = (got == st_ref).all()
    if energies:
        en = len(a) - 2.0 * hist[:, :E].astype(np.float64).T
        ok = ok and ((en[:, 0] == en_ref[:, -1]).all() if (mode != 'phase' and perbeta) else (en == en_ref).all())
    cnt[mode] = cnt.get(mode, 0) + 1
    if not ok:
        bad.append(tag); print('MISMATCH', tag, flush=True)
print('checkerboard kernels:', cnt, 'bad:', bad, '%.0f s' % (time.time() - t0), flush=True)
all_bad += bad

# ---- general graphs, strips ----
lib.emu_general_group.restype = C.c_int
lib.emu_general_group.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double,
                                  C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
lib.emu_strip_phase.restype = C.c_int
lib.emu_strip_phase.argtypes = [C.c_void_p] + [C.c_uint32] * 7 + [C.c_uint64, C.c_uint32, C.c_double, C.c_double,
                                                                  C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
rng = np.random.default_rng(3)
bad = []; t0 = time.time(); n_ok = 0
for i in range(NCASES):
    n = int(rng.integers(5, 90))
    maxdeg = int(rng.integers(1, 16))
    m = int(min(n * maxdeg // 2 * rng.uniform(0.3, 0.95), n * (n - 1) // 2 * 0.8))
    if m < 1: continue
    a, b = T.random_sparse(n, m, rng, maxdeg)
    jabs = float(rng.choice([0.5, 1.0, 2.0]))
    j = rng.choice([-jabs, jabs], size=len(a))
    nn = int(max(a.max(), b.max())) + 1
    col, adj = T.greedy_colouring(n, a, b)
    groups = T.colour_degree_groups(n, a, b, j, col, adj)
    W = int(rng.integers(1, 7)); E = W * 32 - int(rng.integers(0, 32))
    V = 2 if W % 2 == 0 else 1
    K, rounds = [(6, 7), (5, 7), (7, 10), (6, 10)][rng.integers(4)]
    perbeta = bool(rng.integers(2))
    spec = bool(rng.integers(2))
    init = rng.integers(0, 2, size=(E, n)).astype(bool)
    words = T.pack_natural(init, W)
    seed, sweep0, gw0, nsw = 0xC0FFEE1234 + i, 11, 2, 2
    if perbeta:
        betas_e = np.geomspace(0.05, 2.5, E)
        plane, low = T.per_replica_tables(betas_e, W, K, jabs)
        tp, tl = plane.ctypes.data, low.ctypes.data
        sb = [0.0] * nsw
    else:
        tp = tl = None; sb = [0.15, 1.3]
    rc = 0
    for s, beta in enumerate(sb):
        for d, sites, nbr, anti in groups:
            if d == 0: continue
            rc |= lib.emu_general_group(words.ctypes.data, W, V, sites.ctypes.data, nbr.ctypes.data, anti.ctypes.data,
                                        len(sites), d, sweep0 + s, seed, gw0, K, rounds, float(beta), jabs, tp, tl, int(spec), int(rng.integers(1, 6)))
    got = T.unpack_natural(words, E)
    kw = dict(replica_offset=32 * gw0, planes=K, rounds=rounds, states=init, sweep0=sweep0)
    if perbeta:
        _, ref = oracle_lib.msc_mirror(a, b, j, n, col, E, seed, None, per_replica_beta=betas_e, nsweeps=nsw, **kw)
    else:
        _, ref = oracle_lib.msc_mirror(a, b, j, n, col, E, seed, sb, **kw)
    tag = (n, m, maxdeg, W, E, K, rounds, perbeta, spec)
    if rc != 0 or not (got == ref).all():
        bad.append(tag); print('BAD', tag, rc, flush=True)
    else: n_ok += 1
print('general graphs:', n_ok, 'ok, bad:', bad, '%.0f s' % (time.time() - t0), flush=True)
all_bad += bad

bad = []; n_ok = 0
for i in range(NCASES):
    Wr = int(rng.integers(1, 10)); Lx = 64 * Wr
    nstrips = int(rng.integers(1, 4)); rows = 2 * int(rng.integers(1, 5)); Ly = rows * nstrips
    K, rounds = [(6, 7), (5, 7), (7, 10)][rng.integers(3)]
    j = float(rng.choice([-1.0, 1.0, 0.5]))
    init = rng.integers(0, 2, size=(Ly, Lx)).astype(bool)
    betas = [0.2, 0.44, 0.9][:int(rng.integers(1, 4))]
    seed = 0xABCDEF0123456789 ^ i
    full = T.strip_pack(init, Wr)
    strips = []
    for k in range(nstrips):
        buf = np.zeros((2, rows + 2, Wr), dtype=np.uint32); buf[:, 1:-1] = full[:, k*rows:(k+1)*rows]; strips.append(buf)
    def exchange(c):
        for k, buf in enumerate(strips):
            buf[c, 0] = strips[(k - 1) % nstrips][c, rows]; buf[c, rows + 1] = strips[(k + 1) % nstrips][c, 1]
    rc = 0
    for t, beta in enumerate(betas):
        for c in (0, 1):
            exchange(1 - c)
            for k, buf in enumerate(strips):
                rc |= lib.emu_strip_phase(buf.ctypes.data, Wr, rows, k * rows, Ly, 1, c, t, seed, 0xFFFFFFFF if j > 0 else 0,
                                          float(beta), abs(j), K, rounds, 1, rows, int(rng.integers(1, 5)))
    got = T.strip_unpack(np.concatenate([b_[:, 1:-1] for b_ in strips], axis=1), Lx)
    _, ref = oracle_lib.msc_mirror_single(Lx, Ly, j, seed, betas, planes=K, rounds=rounds, state=init)
    tag = (Wr, nstrips, rows, K, rounds, j, len(betas))
    if rc != 0 or not (got == ref).all():
        bad.append(tag); print('BAD strip', tag, rc, flush=True)
    else: n_ok += 1
print('strips:', n_ok, 'ok, bad:', bad, '%.0f s' % (time.time() - t0), flush=True)
all_bad += bad
sys.exit(1 if all_bad else 0)
