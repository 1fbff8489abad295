"""RNG known-answer vectors against the oracle, an independent pure-Python
restatement of the integer spec, and the backend's device functions (host compile)."""
import json
import os

import numpy as np

from conftest import REPO, ptr
from oracle import oracle

KAT = json.load(open(os.path.join(REPO, "tests", "golden", "rng_kat.json")))["vectors"]
M = 0xFFFFFFFF


def py_prng_seed(px, py, frame):  # main.glsl:176-181 with Python ints
    s = [(px * 0x9E3779B9 + frame) & M, (py * 0x9E3779B9 + frame) & M]
    s = [x ^ (x >> 16) for x in s]
    return [(x * 0x9E3779B9) & M for x in s]


def py_pcg2d(s):  # main.glsl:163-174
    x, y = [(1664525 * v + 1013904223) & M for v in s]
    x = (x + 1664525 * y) & M; y = (y + 1664525 * x) & M
    x ^= x >> 16; y ^= y >> 16
    x = (x + 1664525 * y) & M; y = (y + 1664525 * x) & M
    x ^= x >> 16; y ^= y >> 16
    r = (np.array([x, y], np.uint32).astype(np.float32) * np.float32(2.32830643654e-10)).view(np.uint32)
    return [x, y], [int(r[0]), int(r[1])]


def hexes(v):
    return [int(x, 16) for x in v]


def test_two_to_minus_32_is_exact():
    assert np.float32(2.32830643654e-10).view(np.uint32) == 0x2F800000


def test_python_spec_matches_kat():
    for k in KAT:
        seed = py_prng_seed(*k["pixel"], k["frame"])
        assert seed == hexes(k["seed"])
        st1, r = py_pcg2d(seed)
        assert st1 == hexes(k["state1"]) and r == hexes(k["r_bits"])
        st2, _ = py_pcg2d(st1)
        assert st2 == hexes(k["state2"])


def test_oracle_matches_kat():
    for k in KAT:
        seed = oracle.prng_seed(k["pixel"][0], k["pixel"][1], k["frame"])
        assert [int(x) for x in seed] == hexes(k["seed"])
        st1, r = oracle.pcg2d(seed)
        assert [int(x) for x in st1] == hexes(k["state1"])
        assert [int(x) for x in r.view(np.uint32)] == hexes(k["r_bits"])
        st2, _ = oracle.pcg2d(st1)
        assert [int(x) for x in st2] == hexes(k["state2"])


def test_device_functions_match_kat(devcheck):
    for k in KAT:
        seed, r, st = np.zeros(2, np.uint32), np.zeros(2, np.float32), np.zeros(2, np.uint32)
        devcheck.devcheck_rng(k["pixel"][0], k["pixel"][1], k["frame"], ptr(seed), ptr(r), ptr(st))
        assert [int(x) for x in seed] == hexes(k["seed"])
        assert [int(x) for x in r.view(np.uint32)] == hexes(k["r_bits"])
        assert [int(x) for x in st] == hexes(k["state1"])


def test_sincos_contract(devcheck):
    """Oracle and device sin/cos agree bit-for-bit on the whole argument range of the path
    ([0, 2*pi]) and stay within 2 ulp-ish (2.5e-7 abs) of libm."""
    xs = np.concatenate([np.linspace(0, 6.2831855, 20001, dtype=np.float32),
                         np.float32(6.2831853) * np.float32(0.25) * np.random.default_rng(0).random(5000, dtype=np.float32)])
    worst = 0.0
    for x in xs:
        a = oracle.sincos(x)
        b = np.zeros(2, np.float32)
        devcheck.devcheck_sincos(float(x), ptr(b))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        worst = max(worst, abs(float(a[0]) - np.sin(np.float64(x))), abs(float(a[1]) - np.cos(np.float64(x))))
    assert worst < 2.5e-7
