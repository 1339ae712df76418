"""The final pass's distance arithmetic (sdf_dist_f32 in sdf_tools_b200/csrc/edt_kernels.cu) replaces sqrt.f64 by an
rsqrt seed, one coupled Goldschmidt step and one Markstein correction in FP64.  This CPU test restates the sequence
with C's fma() (tests/emul/sqrt_seq.c) and checks that it returns the correctly rounded double sqrt for EVERY
m < 2^25 under seed errors larger than the hardware's (|rel| <= 2^-21 against MUFU.RSQ's 2^-22.4); the device itself
is checked value by value in tests/test_gpu_parity.py::test_distance_arithmetic_matches_ieee_sqrt."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "sqrt_seq.c")
LIB = os.path.join(HERE, "emul", "_build", "libsqrt_seq.so")


def test_newton_sequence_is_correctly_rounded():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    lib = C.CDLL(LIB)
    lib.sqrt_seq_count_bad.restype = C.c_long
    lib.sqrt_seq_count_bad.argtypes = [C.c_int, C.c_int]
    assert lib.sqrt_seq_count_bad(1, 1 << 25) == 0
