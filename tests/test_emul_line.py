"""CPU tests of the line-per-thread kernel ALGORITHM: sdf_tools_b200/csrc/line_hull.cuh (the same source the CUDA
kernel compiles) run on the host one line at a time (tests/emul/emul_line.cpp), against the oracle and the definition."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import synth
from test_emul_band import line_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul_line.cpp")
LIB = os.path.join(HERE, "emul", "_build", "libemul_line.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr = os.path.join(HERE, "..", "sdf_tools_b200", "csrc", "line_hull.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
                               "-o", LIB, SRC])
    L = C.CDLL(LIB)
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    L.emul_line_signed_d2.argtypes = [u8p, C.c_int64, C.c_int64, C.c_int64, i32p, C.c_int]
    L.emul_line_pass.argtypes = [i32p, i32p, C.c_int, C.c_int64, C.c_int64, C.c_int]
    L.emul_line_entry_format.argtypes = [C.c_int64, C.c_int64]
    return L


def test_full_transform_random_grids(lib):
    rng = np.random.default_rng(0)
    for trial in range(60):
        shape = tuple(int(v) for v in rng.integers(1, 70, size=3))
        p = rng.choice([0.0, 0.001, 0.01, 0.05, 0.3, 0.5, 0.9, 0.99, 1.0])
        m = (rng.random(shape) < p).astype(np.uint8)
        got = np.empty(shape, np.int32)
        assert lib.emul_line_signed_d2(m, *shape, got, trial % 3) == 0
        assert np.array_equal(got, O.signed_d2(m)), (shape, p)


def adversarial_lines(rng, n, ncols, hmax):
    INF = O.INF
    lines = []
    for _ in range(ncols):
        kind = rng.integers(0, 7)
        if kind == 0: h = rng.integers(1, 50, n)
        elif kind == 1: h = rng.integers(1, hmax, n)
        elif kind == 2: h = np.abs(np.arange(n) - rng.integers(0, n)) ** 2 + rng.integers(1, 20, n)
        elif kind == 3:
            h = np.full(n, INF); k = rng.integers(0, n, max(1, n // 50)); h[k] = rng.integers(1, 10000, len(k))
        elif kind == 4: h = np.where(rng.random(n) < 0.5, INF, rng.integers(1, 1000, n))
        elif kind == 5: h = (np.arange(n) - rng.integers(0, n)) ** 2 % 7 + 1            # tiny alphabet: ties
        else:                                                                          # exact parabolas: collinear lifted points
            h = np.minimum(hmax - 1, (int(rng.integers(1, 30)) + np.abs(np.arange(n) - rng.integers(0, n))) ** 2)
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
    return lines


def test_line_passes_adversarial(lib):
    rng = np.random.default_rng(7)
    for trial in range(80):
        n = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 65, 100, 255, 256, 511, 513, 1000, 1024, 1025, 2047, 2048]))
        ncols = int(rng.integers(1, 5))
        for wide in (0, 1, 2):                          # entry formats {pos:10,h:22}, {pos:11,h:21}, 64-bit
            if (wide == 0 and n > 1024) or (wide == 1 and n > 2048):
                continue
            hmax = (3 * n * n + 2) if wide == 2 else min(3 * n * n + 2, (1 << (22 if wide == 0 else 21)) - 1)
            lines = adversarial_lines(rng, n, ncols, hmax)
            g = np.ascontiguousarray(np.stack(lines, axis=1))
            want = np.stack([line_oracle(l) for l in lines], axis=1)
            got = np.empty_like(g)
            assert lib.emul_line_pass(g, got, n, ncols, ncols, wide) == 0
            assert np.array_equal(got, want), (n, wide, trial)


@pytest.mark.parametrize("name,shape", [("c2", (48, 48, 48)), ("c3", (64, 64, 64)), ("c5", (64, 64, 64))])
def test_structured_configs(lib, name, shape):
    m = synth.make_mask(name, shape)
    got = np.empty(shape, np.int32)
    assert lib.emul_line_signed_d2(m, *shape, got, 0) == 0
    assert np.array_equal(got, O.signed_d2(m))


def test_entry_format_choice(lib):
    """The narrowest stack entry that holds a pass: positions < n, heights <= hmax."""
    f = lib.emul_line_entry_format
    assert f(1024, (1 << 22) - 2) == 0 and f(1024, (1 << 22) - 1) == 2 and f(1025, 100) == 1      # all-ones = terminator
    assert f(2048, (1 << 21) - 2) == 1 and f(2048, (1 << 21) - 1) == 2 and f(2049, 1) == 2
    assert f(1024, 2 * 1023 ** 2) == 0            # x pass of a 1024^3 grid
    assert f(2048, 1023 ** 2) == 1                # y pass of 2048 x 2048 x 1024
    assert f(2048, 1023 ** 2 + 2047 ** 2) == 2    # its x pass


def test_paths_are_exercised(lib):
    """The adversarial lines must actually reach the data-dependent paths of the sweeps: pops (incl. chains long enough to
    reload the third chain element from the stack), run changes, query advances -- and agree with the banded kernel's
    emulator on the same lines (two independent implementations of the same pass)."""
    import test_emul_band as TB
    L = C.CDLL(LIB)
    L.emul_line_counters.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    ctr = (C.c_longlong * 8)()
    L.emul_line_counters(ctr, 1)
    rng = np.random.default_rng(23)
    n, ncols = 700, 6
    lines = adversarial_lines(rng, n, ncols, 1 << 21)
    # a staircase that ends in a deep pit: the pit pops hundreds of vertices in one chain
    h = (np.arange(n)[::-1] // 2 + 5) ** 2
    h[-1] = 1
    lines[0] = h.astype(np.int32)
    g = np.ascontiguousarray(np.stack(lines, axis=1))
    want = np.stack([line_oracle(l) for l in lines], axis=1)
    got = np.empty_like(g)
    assert lib.emul_line_pass(g, got, n, ncols, ncols, 0) == 0
    assert np.array_equal(got, want)
    L.emul_line_counters(ctr, 0)
    sites, finite, pops, runs, adv = ctr[0], ctr[1], ctr[2], ctr[3], ctr[4]
    assert sites == n * ncols and 0 < finite <= sites and pops > 300 and adv > 0, list(ctr)
    # the banded emulator on the same lines (in place)
    if os.path.exists(TB.LIB):
        B = C.CDLL(TB.LIB)
        i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        B.emul_band_pass.argtypes = [i32p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64]
        gb = g.copy()
        assert B.emul_band_pass(gb, n, ncols, ncols, 5, 16, 0, 1) == 0
        assert np.array_equal(gb, got)
