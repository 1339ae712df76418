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
TAGGED_CELL_DTYPE = np.dtype([("occupancy", np.float32), ("component", np.uint32), ("object_id", np.uint32),
                              ("convex_segment", np.uint32)])            # tagged_object_collision_map.hpp:22-43
QUERY_OBJECT, QUERY_ANY, QUERY_NAMED = 0, 1, 2


def _check_out(out, shape):
    """The C ABI writes prod(shape) floats through the pointer it is given: refuse anything else."""
    if out is None:
        return np.empty(shape, dtype=np.float32)
    if not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags.c_contiguous or \
            tuple(out.shape) != tuple(shape) or not out.flags.writeable:
        raise ValueError(f"out must be a writeable C-contiguous float32 array of shape {tuple(shape)}")
    return out


class Context:
    """One device + stream + staging buffers (sdf_b200_ctx).  Not thread-safe."""

    def __init__(self, device=0):
        """device: a CUDA ordinal, or a sequence of ordinals for a multi-device context (sdf_b200_create_multi: the host
        entry points then shard the grid into z slabs over the devices; device-buffer calls act on the first one)."""
        self._lib = _capi.load()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            devs = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self._lib.sdf_b200_create_multi(C.byref(h), devs, len(device))
            self.devices = [int(d) for d in device]
        else:
            rc = self._lib.sdf_b200_create(C.byref(h), int(device))
            self.devices = [int(device)]
        if rc != _capi.OK:
            raise SdfB200Error(rc, self._lib.sdf_b200_last_error(None).decode())
        self._h = h
        self.device = self.devices[0]

    @property
    def exchange_kind(self) -> str:
        return self._lib.sdf_b200_exchange_kind(self._h).decode()

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
        out = _check_out(out, m.shape)
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_build_from_mask(self._h, m.ctypes.data, *m.shape, float(resolution),
                                                       out.ctypes.data, C.byref(mx), C.byref(mn)))
        return out, (mx.value, mn.value)

    def build_from_cells(self, cells: np.ndarray, resolution: float, unknown_is_filled: bool = False,
                         out: np.ndarray | None = None):
        """cells (nx,ny,nz) of COLLISION_CELL_DTYPE (8 B, as CollisionMapGrid stores them)."""
        c = np.ascontiguousarray(cells)
        if c.ndim != 3 or c.dtype.itemsize != 8 or (c.dtype.names is not None and (
                c.dtype.names[0] != "occupancy" or c.dtype.fields["occupancy"][0] != np.float32
                or c.dtype.fields["occupancy"][1] != 0)):
            raise ValueError("cells must be (nx,ny,nz) of 8-byte COLLISION_CELL records {float32 occupancy; uint32 component}")
        out = _check_out(out, c.shape)
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_build_from_cells(self._h, c.ctypes.data, int(bool(unknown_is_filled)),
                                                        *c.shape, float(resolution), out.ctypes.data,
                                                        C.byref(mx), C.byref(mn)))
        return out, (mx.value, mn.value)

    def build_with_border(self, mask: np.ndarray, resolution: float, mask_second: np.ndarray | None = None):
        """add_virtual_border = true (sdf_b200_build_with_border): (sdf (nx,ny,nz) f32, (max, min)).  mask_second: the
        predicate as the reference's SECOND extraction would see it (stateful predicates); default: the same mask."""
        m = np.ascontiguousarray(mask)
        if m.dtype != np.uint8:
            m = (m != 0).astype(np.uint8)
        if m.ndim != 3:
            raise ValueError("mask must have shape (nx, ny, nz)")
        m2 = m if mask_second is None else np.ascontiguousarray((np.asarray(mask_second) != 0).astype(np.uint8))
        if m2.shape != m.shape:
            raise ValueError("mask_second must have the shape of mask")
        out = np.empty(m.shape, dtype=np.float32)
        mx, mn = C.c_double(), C.c_double()
        self._check(self._lib.sdf_b200_build_with_border(self._h, m.ctypes.data, m2.ctypes.data, *m.shape, float(resolution),
                                                         out.ctypes.data, C.byref(mx), C.byref(mn)))
        return out, (mx.value, mn.value)

    def build_object_sdfs(self, tagged_cells: np.ndarray, resolution: float, queries, unknown_is_filled: bool = False):
        """Batched TaggedObjectCollisionMapGrid SDFs (sdf_b200_build_object_sdfs).  tagged_cells: (nx,ny,nz) of
        TAGGED_CELL_DTYPE; queries: object ids (ints: cells of that object) or (object_id, kind) pairs with kind in
        QUERY_OBJECT / QUERY_ANY / QUERY_NAMED.  Returns (sdfs (nq,nx,ny,nz) float32, [(max, min)] * nq)."""
        c = np.ascontiguousarray(tagged_cells)
        if c.ndim != 3 or c.dtype != TAGGED_CELL_DTYPE:
            raise ValueError("tagged_cells must be (nx,ny,nz) of TAGGED_CELL_DTYPE (16-byte records)")
        qs = [(int(q), QUERY_OBJECT) if np.isscalar(q) else (int(q[0]), int(q[1])) for q in queries]
        arr = np.array(qs, dtype=np.uint32).reshape(-1, 2)
        out = np.empty((len(qs),) + c.shape, dtype=np.float32)
        mx = (C.c_double * max(len(qs), 1))()
        mn = (C.c_double * max(len(qs), 1))()
        self._check(self._lib.sdf_b200_build_object_sdfs(self._h, c.ctypes.data, *c.shape, float(resolution),
                                                         int(bool(unknown_is_filled)), arr.ctypes.data, len(qs),
                                                         out.ctypes.data, mx, mn))
        return out, [(mx[k], mn[k]) for k in range(len(qs))]

    def gradient(self, sdf: np.ndarray, resolution: float, enable_edge_gradients: bool = True, dtype=np.float64):
        """(nx,ny,nz,3) grid-aligned gradient field of an f32 SDF (sdf_b200_gradient_host; reference sdf.hpp:341-358, :432-526)."""
        f = np.ascontiguousarray(sdf, dtype=np.float32)
        if f.ndim != 3:
            raise ValueError("sdf must have shape (nx, ny, nz)")
        if np.dtype(dtype) not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("dtype must be float64 or float32")
        out = np.empty(f.shape + (3,), dtype=dtype)
        self._check(self._lib.sdf_b200_gradient_host(self._h, f.ctypes.data, *f.shape, float(resolution),
                                                     int(bool(enable_edge_gradients)), int(np.dtype(dtype) == np.float64),
                                                     out.ctypes.data))
        return out

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

    def slab_summary_scatter_device(self, d_mask_ptr: int, nx: int, ny: int, nz_local: int, peer_gather_ptrs, rank: int):
        """Summary + exchange in one kernel: peer_gather_ptrs[k] = device pointer (valid on this device) of rank k's
        gather buffer of len(peer_gather_ptrs) * nx * ny uint32 words."""
        ptrs = (C.c_void_p * len(peer_gather_ptrs))(*[C.c_void_p(int(p)) for p in peer_gather_ptrs])
        self._check(self._lib.sdf_b200_slab_summary_scatter_device(self._h, C.c_void_p(d_mask_ptr), nx, ny, nz_local,
                                                                   ptrs, len(peer_gather_ptrs), int(rank)))

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

    ENVELOPE_KERNELS = {"auto": 0, "band": 1, "line": 2, "mixed": 3}

    def set_envelope_kernel(self, mode):
        """Kernels of the two strided envelope passes: "auto" (default), "band", "line" or "mixed" (sdf_b200.h); the
        result is identical in every mode."""
        self._check(self._lib.sdf_b200_set_envelope_kernel(self._h, self.ENVELOPE_KERNELS[mode] if isinstance(mode, str) else int(mode)))

    def last_timing(self):
        arr = (C.c_float * 4)()
        self._check(self._lib.sdf_b200_last_timing(self._h, arr))
        return {"pass1_z_ms": arr[0], "pass2_y_ms": arr[1], "pass3_x_ms": arr[2], "total_ms": arr[3]}

    def last_launches(self) -> int:
        return int(self._lib.sdf_b200_last_launches(self._h))

    def last_envelope_kernel(self) -> str:
        """Envelope kernels the last build used: "band", "line" or "mixed"."""
        return {1: "band", 2: "line", 3: "mixed"}.get(int(self._lib.sdf_b200_last_envelope_kernel(self._h)), "band")


def build_sdf(mask: np.ndarray, resolution: float, device: int = 0):
    """One-shot helper: (sdf, (max, min)) for a host mask."""
    with Context(device) as ctx:
        return ctx.build_from_mask(mask, resolution)
