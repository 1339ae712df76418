"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers of the hot path
(`sdf_generation::ExtractSignedDistanceField`, /root/reference/include/sdf_tools/sdf_generation.hpp:209-271):

* ``exact``  -- oracle/edt_oracle.c, a plain-C exact separable EDT x2 + the reference's combine loop;
* ``ref``    -- oracle/_ref/libsdf_ref.so, the reference's OWN text compiled unmodified against
                oracle/ref_shim.hpp (built by oracle/build_ref.sh where /root/reference exists, shipped
                prebuilt to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs import this module.
The product package (sdf_tools_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_EXACT_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libsdf_ref.so")

INF = 0x40000000

_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the C oracle (always possible) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(_EXACT_PATH) or \
            os.path.getmtime(_EXACT_PATH) < os.path.getmtime(os.path.join(_HERE, "edt_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    if force or not os.path.exists(_REF_PATH):
        subprocess.call(["bash", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL,
                        stderr=subprocess.DEVNULL)


_exact = None
_ref = None


def _lib_exact():
    global _exact
    if _exact is None:
        build()
        lib = C.CDLL(_EXACT_PATH)
        i64 = C.c_int64
        lib.oracle_edt_d2.argtypes = [_u8p, i64, i64, i64, _i32p, C.c_int]
        lib.oracle_signed_d2.argtypes = [_u8p, i64, i64, i64, _i32p, C.c_int]
        lib.oracle_build_sdf.argtypes = [_u8p, i64, i64, i64, C.c_double, _f32p,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        lib.oracle_pack_cells.argtypes = [C.c_void_p, i64, C.c_int, _u8p]
        _exact = lib
    return _exact


def have_ref() -> bool:
    if not os.path.exists(_REF_PATH):
        build()
    return os.path.exists(_REF_PATH)


def _lib_ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libsdf_ref.so is not built (needs /root/reference at build time)")
        lib = C.CDLL(_REF_PATH)
        i64 = C.c_int64
        lib.ref_build_sdf.restype = C.c_double
        lib.ref_build_sdf.argtypes = [_u8p, i64, i64, i64, C.c_double, C.c_float, _f32p,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.ref_build_d2.restype = C.c_double
        lib.ref_build_d2.argtypes = [_u8p, i64, i64, i64, C.c_int, C.c_int, _f64p]
        lib.ref_sizeof_bucket_cell.restype = C.c_int
        _ref = lib
    return _ref


def _mask3(mask):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    if m.ndim != 3:
        raise ValueError("mask must be (nx, ny, nz)")
    return (m != 0).astype(np.uint8)


# ----------------------------------------------------------------------------------------------- exact oracle
def edt_d2(site_mask, nthreads: int = 0) -> np.ndarray:
    """Exact squared EDT to the nearest voxel with site_mask != 0 (INF where there is none)."""
    m = _mask3(site_mask)
    out = np.empty(m.shape, dtype=np.int32)
    _lib_exact().oracle_edt_d2(m, *m.shape, out, nthreads)
    return out


def signed_d2(mask, nthreads: int = 0) -> np.ndarray:
    """+d2(nearest filled) at free voxels, -d2(nearest free) at filled voxels."""
    m = _mask3(mask)
    out = np.empty(m.shape, dtype=np.int32)
    rc = _lib_exact().oracle_signed_d2(m, *m.shape, out, nthreads)
    if rc:
        raise MemoryError("oracle_signed_d2")
    return out


def build_sdf(mask, res: float, nthreads: int = 0):
    """(sdf f32 (nx,ny,nz), max f64, min f64) -- sdf_generation.hpp:209-271 with the exact EDT."""
    m = _mask3(mask)
    out = np.empty(m.shape, dtype=np.float32)
    mx, mn = C.c_double(), C.c_double()
    rc = _lib_exact().oracle_build_sdf(m, *m.shape, float(res), out, C.byref(mx), C.byref(mn), nthreads)
    if rc:
        raise MemoryError("oracle_build_sdf")
    return out, mx.value, mn.value


def pack_cells(cells: np.ndarray, unknown_is_filled: bool) -> np.ndarray:
    """COLLISION_CELL[] -> uint8 mask with the rule of collision_map.hpp:686-709."""
    c = np.ascontiguousarray(cells)
    assert c.dtype.itemsize == 8
    out = np.empty(c.shape, dtype=np.uint8)
    _lib_exact().oracle_pack_cells(c.ctypes.data, c.size, int(unknown_is_filled), out.reshape(-1))
    return out


def brute_d2(site_mask) -> np.ndarray:
    """O(N * sites) brute force, tiny grids only -- pins the exact oracle itself."""
    m = _mask3(site_mask)
    pts = np.argwhere(m != 0).astype(np.int64)
    out = np.full(m.shape, INF, dtype=np.int64)
    if len(pts):
        gx, gy, gz = np.meshgrid(*[np.arange(s, dtype=np.int64) for s in m.shape], indexing="ij")
        for p in pts:
            d = (gx - p[0]) ** 2 + (gy - p[1]) ** 2 + (gz - p[2]) ** 2
            np.minimum(out, d, out=out)
    return out.astype(np.int32)


# ------------------------------------------------------------------------------------------- reference itself
def ref_build_sdf(mask, res: float, oob: float = 0.0):
    """The reference's own ExtractSignedDistanceField<T> -> (sdf, max, min, wall_ms)."""
    m = _mask3(mask)
    out = np.empty(m.shape, dtype=np.float32)
    mx, mn = C.c_double(), C.c_double()
    ms = _lib_ref().ref_build_sdf(m, *m.shape, float(res), float(oob), out, C.byref(mx), C.byref(mn))
    if ms < 0:
        raise RuntimeError("reference threw")
    return out, mx.value, mn.value, ms


def ref_build_d2(mask, want_filled: bool, reverse_seeds: bool = False):
    """The reference's own BuildDistanceField -> (distance_square f64, wall_ms)."""
    m = _mask3(mask)
    out = np.empty(m.shape, dtype=np.float64)
    ms = _lib_ref().ref_build_d2(m, *m.shape, int(want_filled), int(reverse_seeds), out)
    if ms < 0:
        raise RuntimeError("reference threw")
    return out, ms


def ref_sizeof_bucket_cell() -> int:
    return _lib_ref().ref_sizeof_bucket_cell()
