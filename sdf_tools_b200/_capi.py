"""sdf_tools_b200/_capi.py -- ctypes declarations of include/sdf_b200.h (the C ABI of libsdf_b200.so)."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

OK = 0
ERR_CUDA, ERR_ARG, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4
MAX_AXIS = 2048
INF = 0x40000000

# every symbol include/sdf_b200.h declares: (name, restype, argtypes)
_i64, _dbl, _vp, _int = C.c_int64, C.c_double, C.c_void_p, C.c_int
_pd = C.POINTER(C.c_double)
SYMBOLS = [
    ("sdf_b200_version", C.c_char_p, []),
    ("sdf_b200_device_count", _int, []),
    ("sdf_b200_create", _int, [C.POINTER(_vp), _int]),
    ("sdf_b200_create_multi", _int, [C.POINTER(_vp), C.POINTER(C.c_int), _int]),
    ("sdf_b200_num_devices", _int, [_vp]),
    ("sdf_b200_exchange_kind", C.c_char_p, [_vp]),
    ("sdf_b200_destroy", None, [_vp]),
    ("sdf_b200_last_error", C.c_char_p, [_vp]),
    ("sdf_b200_set_stream", _int, [_vp, _vp, _int]),
    ("sdf_b200_synchronize", _int, [_vp]),
    ("sdf_b200_set_envelope_kernel", _int, [_vp, _int]),
    ("sdf_b200_host_alloc", _int, [C.POINTER(_vp), C.c_size_t]),
    ("sdf_b200_host_free", _int, [_vp]),
    ("sdf_b200_build_from_mask", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _vp, _pd, _pd]),
    ("sdf_b200_build_from_cells", _int, [_vp, _vp, _int, _i64, _i64, _i64, _dbl, _vp, _pd, _pd]),
    ("sdf_b200_build_device", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _vp]),
    ("sdf_b200_get_extrema", _int, [_vp, _pd, _pd]),
    ("sdf_b200_pack_cells_device", _int, [_vp, _vp, _i64, _int, _vp]),
    ("sdf_b200_signed_d2_device", _int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    ("sdf_b200_signed_d2_from_mask", _int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    ("sdf_b200_slab_summary_device", _int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    ("sdf_b200_slab_build_device", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _vp, C.POINTER(C.c_int32), _int, _int, _int, _vp]),
    ("sdf_b200_slab_summary_scatter_device", _int, [_vp, _vp, _i64, _i64, _i64, C.POINTER(_vp), _int, _int]),
    ("sdf_b200_get_extrema_raw", _int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("sdf_b200_extrema_from_raw", _int, [C.c_int32, C.c_int32, _dbl, _pd, _pd]),
    ("sdf_b200_build_with_border", _int, [_vp, _vp, _vp, _i64, _i64, _i64, _dbl, _vp, _pd, _pd]),
    ("sdf_b200_build_object_sdfs", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _vp, _int, _vp, _pd, _pd]),
    ("sdf_b200_gradient_device", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _int, _vp]),
    ("sdf_b200_gradient_host", _int, [_vp, _vp, _i64, _i64, _i64, _dbl, _int, _int, _vp]),
    ("sdf_b200_last_timing", _int, [_vp, C.POINTER(C.c_float)]),
    ("sdf_b200_last_launches", _int, [_vp]),
    ("sdf_b200_last_envelope_kernel", _int, [_vp]),
    ("sdf_b200_debug_phase_clocks", _int, [_vp, _int, C.POINTER(C.c_longlong), _int]),
]

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libsdf_b200.so.  Fails loudly if it has not been built: there is no CPU or PyTorch fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("SDF_B200_LIB", LIB_PATH)      # override: A/B runs of kernel variants (tools/)
        if not os.path.exists(path):
            raise LibraryMissing(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sdf_tools_b200 has no CPU fallback.")
        lib = C.CDLL(path)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
