"""CPU tests of the kernel ALGORITHM: the phase functions of sdf_tools_b200/csrc/run_hull.cuh (the same source the
CUDA kernel compiles) run on the host as loops over the threads of a tile (tests/emul/emul_run.cpp), in natural and
shuffled thread order, against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_run.cpp")
LIB = os.path.join(HERE, "emul", "_build", "libemul_run.so")
LPCS = [4, 8, 16, 32]                                                    # lines per tile


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = os.path.join(HERE, "..", "sdf_tools_b200", "csrc", "run_hull.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    L = C.CDLL(LIB)
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    L.emul_run_signed_d2.argtypes = [u8p, C.c_int64, C.c_int64, C.c_int64, i32p, C.c_int, C.c_int, C.c_uint64]
    L.emul_run_pass.argtypes = [i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_uint64]
    return L


def line_oracle(s):
    """out(p) = min over q of (p-q)^2 + (|s(q)| if same polarity else 0): the definition, O(n^2)."""
    n = len(s)
    h = np.abs(s.astype(np.int64)); pol = s < 0; q = np.arange(n)
    out = np.empty(n, np.int64)
    for p in range(n):
        same = pol == pol[p]
        cost = (q - p) ** 2 + np.where(same, h, 0)
        cost = np.where(same & (h >= O.INF), 2 ** 40, cost)
        out[p] = min(cost.min(), O.INF)
    return np.where(pol, -out, out).astype(np.int32)


def test_full_transform_random_grids(lib):
    rng = np.random.default_rng(0)
    for trial in range(60):
        shape = tuple(int(v) for v in rng.integers(1, 70, size=3))
        p = rng.choice([0.001, 0.01, 0.05, 0.3, 0.5, 0.9, 0.99])
        m = (rng.random(shape) < p).astype(np.uint8)
        lpc = LPCS[trial % len(LPCS)]
        got = np.empty(shape, np.int32)
        assert lib.emul_run_signed_d2(m, *shape, got, lpc, trial % 2, trial) == 0
        assert np.array_equal(got, O.signed_d2(m)), (shape, p, lpc)


def test_line_passes_adversarial(lib):
    rng = np.random.default_rng(7)
    INF = O.INF
    for trial in range(60):
        n = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 100, 255, 256, 511, 513, 1000, 1024, 1025, 2047, 2048]))
        ncols = int(rng.integers(1, 5))
        lines = []
        for _ in range(ncols):
            kind = rng.integers(0, 5)
            if kind == 0: h = rng.integers(1, 50, n)
            elif kind == 1: h = rng.integers(1, 3 * n * n + 2, n)
            elif kind == 2: h = np.abs(np.arange(n) - rng.integers(0, n)) ** 2 + rng.integers(1, 20, n)
            elif kind == 3:
                h = np.full(n, INF); k = rng.integers(0, n, max(1, n // 50)); h[k] = rng.integers(1, 10000, len(k))
            else: h = np.where(rng.random(n) < 0.5, INF, rng.integers(1, 1000, n))
            h = np.minimum(h, INF)
            pp = rng.choice([0.0, 0.001, 0.02, 0.3, 0.5, 1.0])
            pol = np.zeros(n, bool)
            if pp > 0:
                if rng.random() < 0.5:
                    pol = rng.random(n) < pp
                else:
                    x = 0
                    while x < n:
                        L = int(rng.integers(1, max(2, n // 4))); pol[x:x + L] = rng.random() < pp; x += L
            lines.append(np.where(pol, -h, h).astype(np.int32))
        g = np.ascontiguousarray(np.stack(lines, axis=1))
        want = np.stack([line_oracle(l) for l in lines], axis=1)
        for lpc in LPCS:
            got = g.copy()
            assert lib.emul_run_pass(got, n, ncols, ncols, lpc, trial % 2, trial) == 0
            assert np.array_equal(got, want), (n, lpc, trial)


@pytest.mark.parametrize("name,shape", [("c2", (48, 48, 48)), ("c3", (64, 64, 64)), ("c5", (64, 64, 64))])
def test_structured_configs(lib, name, shape):
    m = synth.make_mask(name, shape)
    got = np.empty(shape, np.int32)
    assert lib.emul_run_signed_d2(m, *shape, got, 16, 1, 3) == 0
    assert np.array_equal(got, O.signed_d2(m))


def test_merge_walks_on_plateaus(lib):
    """Step / plateau height profiles (a box above a ground plane seen along y) make the seam merges drop hundreds of
    consecutive vertices, across whole bands, down to and including the run's boundary sites."""
    rng = np.random.default_rng(11)
    ctr = (C.c_longlong * 16)()
    lib.emul_run_counters(ctr, 1)
    for trial in range(30):
        n = int(rng.choice([256, 500, 1024, 1500, 2048]))
        lines = []
        for _ in range(4):
            h = np.full(n, int(rng.integers(100, 900)) ** 2, np.int64)
            for _k in range(int(rng.integers(1, 4))):                   # low plateaus, pits and ramps
                a = int(rng.integers(0, n)); w = int(rng.integers(1, max(2, n // 3)))
                kind = rng.integers(0, 3)
                if kind == 0: h[a:a + w] = int(rng.integers(1, 90)) ** 2
                elif kind == 1: h[a:a + 1] = 1
                else: h[a:a + w] = (np.arange(len(h[a:a + w])) + int(rng.integers(1, 50))) ** 2
            pol = np.zeros(n, bool)
            if rng.random() < 0.5:
                a = int(rng.integers(0, n)); pol[a:a + int(rng.integers(1, 40))] = True
            lines.append(np.where(pol, -h, h).astype(np.int32))
        g = np.ascontiguousarray(np.stack(lines, axis=1))
        want = np.stack([line_oracle(l) for l in lines], axis=1)
        for lpc in [16, 8, 4]:
            got = g.copy()
            assert lib.emul_run_pass(got, n, 4, 4, lpc, trial % 2, trial) == 0
            assert np.array_equal(got, want), (n, lpc)
    lib.emul_run_counters(ctr, 0)
    assert ctr[5] > 1000, list(ctr)                                    # the seam walks did drop vertices


def test_ties_and_collinear_profiles(lib):
    """Heights from tiny alphabets and exact parabolas make many hull tests come out EQUAL (collinear lifted points,
    equal candidates): the >= / <= conventions of build, seam walks and query must agree with the definition."""
    rng = np.random.default_rng(23)
    INF = O.INF
    for trial in range(120):
        n = int(rng.choice([7, 32, 33, 64, 96, 128, 129, 200, 256, 400]))
        kind = trial % 6
        q = np.arange(n, dtype=np.int64)
        if kind == 0:   h = rng.integers(1, 3, n)                                   # {1, 2}
        elif kind == 1: h = np.full(n, int(rng.integers(1, 5000)))                   # flat: every site is a vertex
        elif kind == 2: h = (q - int(rng.integers(-5, n + 5))) ** 2 + int(rng.integers(1, 4))      # one exact parabola
        elif kind == 3: h = np.minimum((q - int(rng.integers(0, n))) ** 2, (q - int(rng.integers(0, n))) ** 2) + 1
        elif kind == 4: h = np.where(rng.random(n) < 0.5, 1, (q % 7 + 1) ** 2)
        else:           h = np.where(rng.random(n) < 0.2, INF, rng.integers(1, 4, n))
        h = np.minimum(h, INF)
        pol = np.zeros(n, bool)
        if trial % 3 == 0:
            x = 0
            while x < n:
                L = int(rng.integers(1, max(2, n // 3))); pol[x:x + L] = rng.random() < 0.4; x += L
        s = np.where(pol, -h, h).astype(np.int32)
        want = line_oracle(s)
        g = np.ascontiguousarray(np.stack([s, s[::-1].copy()], axis=1))
        want2 = np.stack([want, line_oracle(s[::-1].copy())], axis=1)
        for lpc in [4, 16]:
            got = g.copy()
            assert lib.emul_run_pass(got, n, 2, 2, lpc, trial % 2, trial) == 0
            assert np.array_equal(got, want2), (n, kind, lpc)


def test_runs_with_boundary_sites(lib):
    """Short runs of both polarities with mixed finite / not-yet-seen heights: boundary sites open and close chains
    inside bands, at band ends, at seams and at the line ends."""
    rng = np.random.default_rng(5)
    INF = O.INF
    for trial in range(150):
        n = int(rng.choice([5, 31, 32, 33, 63, 64, 65, 96, 127, 128, 160, 257, 1024]))
        pol = np.zeros(n, bool)
        x = 0
        maxrun = int(rng.choice([1, 2, 3, 8, 33, 70]))
        cur = bool(rng.integers(0, 2))
        while x < n:
            L = int(rng.integers(1, maxrun + 1)); pol[x:x + L] = cur; cur = not cur; x += L
        kind = trial % 4
        if kind == 0:   h = np.full(n, INF)
        elif kind == 1: h = np.where(rng.random(n) < 0.7, INF, rng.integers(1, 2000, n))
        elif kind == 2: h = rng.integers(1, 40, n)
        else:           h = rng.integers(1, 100000, n)
        s = np.where(pol, -h, h).astype(np.int32)
        g = np.ascontiguousarray(np.stack([s, s[::-1].copy()], axis=1))
        want = np.stack([line_oracle(s), line_oracle(s[::-1].copy())], axis=1)
        for lpc in [4, 32]:
            got = g.copy()
            assert lib.emul_run_pass(got, n, 2, 2, lpc, trial % 2, trial) == 0
            assert np.array_equal(got, want), (n, kind, lpc, trial)
