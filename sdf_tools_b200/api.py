"""sdf_tools_b200/api.py -- host-side calls into the C ABI (numpy in / numpy out, or raw device pointers).

Mirrors the one operation the reference exposes for this path,
``CollisionMapGrid.ExtractSignedDistanceField(oob, unknown_is_filled, add_virtual_border) -> (sdf, (max, min))``
(src/sdf_tools/bindings.cpp:81, include/sdf_tools/collision_map.hpp:680-712), on bulk arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


class SdfB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sdf_b200 error {code}: {msg}")
        self.code = code


COLLISION_CELL_DTYPE = np.dtype([("occupancy", np.float32), ("component", np.uint32)])


class Context:
    """One device + stream + staging buffers (sdf_b200_ctx).  Not thread-safe."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_void_p()
        rc = self._lib.sdf_b200_create(C.byref(h), int(device))
        if rc != _capi.OK:
            raise SdfB200Error(rc, self._lib.sdf_b200_last_error(None).decode())
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sdf_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != _capi.OK:
            raise SdfB200Error(rc, self._lib.sdf_b200_last_error(self._h).decode())

    # ---------------------------------------------------------------------------------------- host buffers
    def build_from_mask(self, mask: np.ndarray, resolution: float, out: np.ndarray | None = None):
        """mask (nx,ny,nz) uint8/bool, non-zero = filled -> (sdf f32 (nx,ny,nz), (max, min))."""
        m = np.ascontiguousarray(mask)
        if m.dtype != np.uint8:
            m = (m != 0).astype(np.uint8)
        if m.ndim != 3:
            raise ValueError("mask must have shape (nx, ny, nz)")
        if out is None:
            out = np.empty(m.shape, dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == m.shape
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_build_from_mask(self._h, m.ctypes.data, *m.shape, float(resolution),
                                                       out.ctypes.data, C.byref(mx), C.byref(mn)))
        return out, (mx.value, mn.value)

    def build_from_cells(self, cells: np.ndarray, resolution: float, unknown_is_filled: bool = False,
                         out: np.ndarray | None = None):
        """cells (nx,ny,nz) of COLLISION_CELL_DTYPE (8 B, as CollisionMapGrid stores them)."""
        c = np.ascontiguousarray(cells)
        if c.dtype.itemsize != 8 or c.ndim != 3:
            raise ValueError("cells must be (nx,ny,nz) of 8-byte COLLISION_CELL records")
        if out is None:
            out = np.empty(c.shape, dtype=np.float32)
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_build_from_cells(self._h, c.ctypes.data, int(bool(unknown_is_filled)),
                                                        *c.shape, float(resolution), out.ctypes.data,
                                                        C.byref(mx), C.byref(mn)))
        return out, (mx.value, mn.value)

    # -------------------------------------------------------------------------------------- device buffers
    def set_stream(self, cuda_stream_ptr: int | None):
        """Launch on the given CUDA stream handle (0 = the default stream); None = the context's own stream."""
        if cuda_stream_ptr is None:
            self._check(self._lib.sdf_b200_set_stream(self._h, C.c_void_p(0), 1))
        else:
            self._check(self._lib.sdf_b200_set_stream(self._h, C.c_void_p(cuda_stream_ptr), 0))

    def build_device(self, d_mask_ptr: int, nx: int, ny: int, nz: int, resolution: float, d_sdf_ptr: int):
        self._check(self._lib.sdf_b200_build_device(self._h, C.c_void_p(d_mask_ptr), nx, ny, nz,
                                                    float(resolution), C.c_void_p(d_sdf_ptr)))

    def signed_d2_device(self, d_mask_ptr: int, nx: int, ny: int, nz: int, d_out_ptr: int):
        self._check(self._lib.sdf_b200_signed_d2_device(self._h, C.c_void_p(d_mask_ptr), nx, ny, nz,
                                                        C.c_void_p(d_out_ptr)))

    def pack_cells_device(self, d_cells_ptr: int, n: int, unknown_is_filled: bool, d_mask_ptr: int):
        self._check(self._lib.sdf_b200_pack_cells_device(self._h, C.c_void_p(d_cells_ptr), n,
                                                         int(bool(unknown_is_filled)), C.c_void_p(d_mask_ptr)))

    # ------------------------------------------------------------------------------------------ slab builds
    def slab_summary_device(self, d_mask_ptr: int, nx: int, ny: int, nz_local: int, d_summary_ptr: int):
        self._check(self._lib.sdf_b200_slab_summary_device(self._h, C.c_void_p(d_mask_ptr), nx, ny, nz_local,
                                                           C.c_void_p(d_summary_ptr)))

    def slab_build_device(self, d_mask_ptr: int, nx: int, ny: int, nz_local: int, resolution: float,
                          d_all_summaries_ptr: int, slab_widths, rank: int, d_sdf_ptr: int, keep_extrema: bool = False):
        w = (C.c_int32 * len(slab_widths))(*[int(v) for v in slab_widths])
        self._check(self._lib.sdf_b200_slab_build_device(self._h, C.c_void_p(d_mask_ptr), nx, ny, nz_local,
                                                         float(resolution), C.c_void_p(d_all_summaries_ptr), w,
                                                         len(slab_widths), int(rank), int(bool(keep_extrema)),
                                                         C.c_void_p(d_sdf_ptr)))

    def get_extrema_raw(self):
        a, b = C.c_int32(), C.c_int32()
        self._check(self._lib.sdf_b200_get_extrema_raw(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def extrema_from_raw(self, max_free_d2: int, max_filled_d2: int, resolution: float):
        mx, mn = C.c_double(), C.c_double()
        self._lib.sdf_b200_extrema_from_raw(int(max_free_d2), int(max_filled_d2), float(resolution),
                                            C.byref(mx), C.byref(mn))
        return mx.value, mn.value

    def get_extrema(self):
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_get_extrema(self._h, C.byref(mx), C.byref(mn)))
        return mx.value, mn.value

    def synchronize(self):
        self._check(self._lib.sdf_b200_synchronize(self._h))

    def last_timing(self):
        arr = (C.c_float * 4)()
        self._check(self._lib.sdf_b200_last_timing(self._h, arr))
        return {"pass1_z_ms": arr[0], "pass2_y_ms": arr[1], "pass3_x_ms": arr[2], "total_ms": arr[3]}

    def last_launches(self) -> int:
        return int(self._lib.sdf_b200_last_launches(self._h))


def build_sdf(mask: np.ndarray, resolution: float, device: int = 0):
    """One-shot helper: (sdf, (max, min)) for a host mask."""
    with Context(device) as ctx:
        return ctx.build_from_mask(mask, resolution)
