"""sdf_tools_b200/build.py -- compiles libsdf_b200.so (sm_100a) in-tree with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsdf_b200.so")
SOURCES = ["edt_kernels.cu", "line_kernels.cu", "capi.cu"]
HEADERS = ["edt_kernels.cuh", "band_hull.cuh", "line_hull.cuh", os.path.join("..", "..", "include", "sdf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source of the package for sm_100a into sdf_tools_b200/lib/libsdf_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH + ".tmp"] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def pybind_path() -> str:
    import sysconfig
    return os.path.join(_HERE, "pysdf_tools" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pybind(force: bool = False) -> str:
    """Compile the pybind11 module `pysdf_tools` (C++ face + Python face of the reference API) with g++."""
    import sysconfig
    import pybind11
    out = pybind_path()
    src = os.path.join(CSRC, "pybind_module.cpp")
    inc = os.path.join(os.path.dirname(_HERE), "include")
    deps = [src, LIB_PATH] + [os.path.join(inc, "sdf_tools", f) for f in ("sdf.hpp", "sdf_generation.hpp", "collision_map.hpp")] \
        + [os.path.join(inc, "arc_utilities", "voxel_grid.hpp"), os.path.join(inc, "sdf_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = "g++" if os.path.exists("/usr/bin/g++") else (os.environ.get("CXX") or "g++")
    cmd = [cxx if cxx != "g++" else "/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I", inc, "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], src, "-o", out + ".tmp",
           "-L", LIB_DIR, "-lsdf_b200", "-Wl,-rpath,$ORIGIN/lib"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ (pybind module) failed:\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
