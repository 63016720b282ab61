"""Generates tests/golden/exact_2d_ising.json: exact finite-lattice (Kaufman 1949) and
infinite-lattice (Onsager) energies of the 2D ferromagnetic Ising model, J = -1 in the
reference's convention E = sum J s s.  These stand in for the golden vectors the reference does
not have.  Run:  python tests/golden/make_golden.py
"""
import json
import os

import mpmath as mp

mp.mp.dps = 50


def kaufman_lnz(K, L):
    K = mp.mpf(K)
    N = L

    def gamma(l):
        if l == 0:
            return 2 * K + mp.log(mp.tanh(K))
        return mp.acosh(mp.cosh(2 * K) * mp.coth(2 * K) - mp.cos(mp.pi * l / N))

    z1 = z2 = z3 = z4 = mp.mpf(1)
    for r in range(N):
        go, ge = gamma(2 * r + 1), gamma(2 * r)
        z1 *= 2 * mp.cosh(N * go / 2)
        z2 *= 2 * mp.sinh(N * go / 2)
        z3 *= 2 * mp.cosh(N * ge / 2)
        z4 *= 2 * mp.sinh(N * ge / 2)
    return mp.log(mp.mpf(1) / 2) + (N * N / mp.mpf(2)) * mp.log(2 * mp.sinh(2 * K)) + mp.log(z1 + z2 + z3 + z4)


def kaufman(beta, L):
    f = lambda b: kaufman_lnz(b, L)
    e = -mp.diff(f, beta) / (L * L)
    c = mp.mpf(beta) ** 2 * mp.diff(f, beta, 2) / (L * L)
    return float(e), float(c)


def onsager(beta):
    K = mp.mpf(beta)
    k = 2 * mp.sinh(2 * K) / mp.cosh(2 * K) ** 2
    e = -mp.coth(2 * K) * (1 + (2 / mp.pi) * (2 * mp.tanh(2 * K) ** 2 - 1) * mp.ellipk(k * k))
    m = (1 - mp.sinh(2 * K) ** -4) ** (mp.mpf(1) / 8) if K > mp.log(1 + mp.sqrt(2)) / 2 else mp.mpf(0)
    return float(e), float(m)


def main():
    out = {"kaufman": {}, "onsager": {}}
    for L in (4, 8, 16, 32, 64, 128):
        for beta in (0.3, 0.4, 0.44, 0.5):
            e, c = kaufman(beta, L)
            out["kaufman"][f"L{L}_b{beta}"] = {"L": L, "beta": beta, "e_per_site": e, "c_per_site": c}
    for beta in (0.30, 0.40, 0.42, 0.43, 0.44, 0.45, 0.46, 0.48, 0.50, 0.60):
        e, m = onsager(beta)
        out["onsager"][f"b{beta}"] = {"beta": beta, "e_per_site": e, "m": m}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exact_2d_ising.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path)


if __name__ == "__main__":
    main()
