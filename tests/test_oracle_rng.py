"""Known-answer tests pinning the oracle's RNG restatements.

The reference has no tests; these vectors come from the generators' authors (Vigna's
xoshiro256++ reference output, SplitMix64, Random123's kat_vectors for Philox4x32-10)."""
import os

import numpy as np
import pytest


def test_xoshiro256pp_reference_vector(oracle):
    rng = oracle.Rng(state=[1, 2, 3, 4])
    expect = [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
              9973669472204895162, 14011001112246962877, 12406186145184390807,
              15849039046786891736, 10450023813501588000]
    assert [rng.next_u64() for _ in range(10)] == expect


def test_seed_from_u64_is_splitmix64(oracle):
    rng = oracle.Rng(seed=0)
    assert [hex(int(x)) for x in rng.s] == ["0xe220a8397b1dcdaf", "0x6e789e6aa1b965f4",
                                            "0x6c45d188009454f", "0xf88bb8a8724c81ec"]


def test_make_seeds_golden(oracle):
    # Lattice(edges, seed_gen=0).make_seeds(4) / seed_gen=42 if rand 0.8 SmallRng is restated right
    assert list(oracle.make_seeds(0, 4)) == [5987356902031041503, 7051070477665621255,
                                             6633766593972829180, 211316841551650330]
    assert list(oracle.make_seeds(42, 4)) == [15021278609987233951, 5881210131331364753,
                                              18149643915985481100, 12933668939759105464]


def test_gen_range_widening_multiply(oracle):
    # restated independently in Python: zone = (n << lzcnt(n)) - 1, accept iff lo <= zone
    for n in (1, 2, 3, 7, 1024, 1000003, 2**40 + 17):
        a, b = oracle.Rng(seed=123), oracle.Rng(seed=123)
        zone = ((n << (64 - n.bit_length())) - 1) & (2**64 - 1)
        for _ in range(200):
            while True:
                m = b.next_u64() * n
                if (m & (2**64 - 1)) <= zone:
                    want = m >> 64
                    break
            got = a.gen_range(n)
            assert got == want and 0 <= got < n
        assert list(a.s) == list(b.s)


def test_gen_range_power_of_two_rejects_half(oracle):
    # n = 2^k: zone = 2^63 - 1 ... the conservative zone rejects ~50% of draws
    a, b = oracle.Rng(seed=5), oracle.Rng(seed=5)
    draws = 0
    for _ in range(2000):
        a.gen_range(1024)
    while list(b.s) != list(a.s):
        b.next_u64()
        draws += 1
        assert draws < 10000
    assert 3500 < draws < 4500


def test_gen_f64_and_bool(oracle):
    a, b = oracle.Rng(seed=9), oracle.Rng(seed=9)
    for _ in range(100):
        v = b.next_u64()
        assert a.gen_f64() == (v >> 11) * 2.0**-53
    for _ in range(100):
        v = b.next_u64()
        assert a.gen_bool() == bool(v >> 63)


def test_philox4x32_10_kat(oracle):
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert list(oracle.philox4x32(ctr, key, 10)) == want


def test_philox4x32_7_kat(oracle):
    # Random123 kat_vectors, philox4x32 7 rounds
    got = oracle.philox4x32([0, 0, 0, 0], [0, 0], 7)
    assert list(got) == [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]


def test_continuation_rounds_of_the_library_header(oracle, tmp_path):
    """The tie resolver's words beyond the calls a site update makes anyway are continuation
    rounds of its last Philox block (csrc/philox.h: philox4x32_more, oracle/msc_mirror.c:
    stream_word_tag).  The library's header is host + device code: compiled here for the host,
    one more round on a finished R-round block must equal the (R + 1)-round function of the same
    counter, for the library's own philox4x32 and for the mirror's."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = tmp_path / "more.cpp"
    src.write_text(r'''
#include <cstdio>
#include "philox.h"
using namespace ising;
int main() {
    const uint32_t c[4] = {0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x01000001u};
    const uint32_t k0 = 0xa4093822u, k1 = 0x299f31d0u;
    u32x4 s7 = philox4x32<7>(c[0], c[1], c[2], c[3], k0, k1);
    u32x4 m8 = philox4x32_more(s7, 7, k0, k1), m9 = philox4x32_more(m8, 8, k0, k1);
    u32x4 f8 = philox4x32<8>(c[0], c[1], c[2], c[3], k0, k1), f9 = philox4x32<9>(c[0], c[1], c[2], c[3], k0, k1);
    const PhiloxKeys pk = philox_round_keys(k0, k1);
    u32x4 p7 = philox4x32_keys<7>(c[0], c[1], c[2], c[3], pk);
    u32x4 q8 = philox4x32_more(p7, 7, pk.k[0], pk.k[1]);
    const int ok = m8.x == f8.x && m8.y == f8.y && m8.z == f8.z && m8.w == f8.w &&
                   m9.x == f9.x && m9.y == f9.y && m9.z == f9.z && m9.w == f9.w &&
                   q8.x == f8.x && q8.y == f8.y && q8.z == f8.z && q8.w == f8.w;
    std::printf("%d %08x %08x %08x %08x\n", ok, f8.x, f8.y, f8.z, f8.w);
    return ok ? 0 : 1;
}
''')
    exe = tmp_path / "more"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyisingmontecarlo_b200", "csrc")
    subprocess.run([gxx, "-std=c++17", "-O1", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "1"
    want = oracle.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x01000001], [0xa4093822, 0x299f31d0], 8)
    assert [int(x, 16) for x in out[1:]] == [int(v) for v in want]
