"""sdf_tools_b200/distributed.py -- slab-sharded SDF build: one process per GPU, the grid cut along z.

The reference is single-process; this is the scale-out of its hot path (`ExtractSignedDistanceField`,
sdf_generation.hpp:209-271) that BASELINE.json asks for.  Only the first (binary) pass runs along z, so the only
thing that crosses a cut is, per z line, the distance from the slab face to the nearest voxel of the other
polarity.  Protocol per build (include/sdf_b200.h "Slab builds"):

  1. every rank summarises its slab: per (x,y) line  {polarity of first voxel, leading-run length,
     polarity of last voxel, trailing-run length}  packed in 4 bytes            (sdf_b200_slab_summary_device)
  2. ONE all-gather of the summaries (nx*ny*4 bytes per rank; NCCL over NVLink)  -- the only data-path collective
  3. every rank resolves its carries locally (also across slabs that are a single run) and runs the three passes
     on its own slab                                                             (sdf_b200_slab_build_device)
  4. all-reduce(max) of two integers gives the global (max, min) pair of sdf_generation.hpp:269.

The numpy functions below restate steps 1-3 for the world_size-2 gloo tests on CPU (host-side protocol logic);
the CUDA kernels in csrc/edt_kernels.cu are what runs in production.
"""
from __future__ import annotations

import numpy as np

FARB = 40000          # "no voxel of the other polarity on this side" (matches the kernels)
INF = 0x40000000


def slab_ranges(nz: int, nranks: int, align: int = 16):
    """Balanced z ranges [z0, z1) per rank, cut at multiples of `align` where possible (16-byte mask loads)."""
    if nranks < 1:
        raise ValueError("nranks must be >= 1")
    if nranks > nz:
        raise ValueError(f"cannot cut {nz} z planes into {nranks} non-empty slabs")
    units = nz // align
    if units >= nranks:
        base, extra = divmod(units, nranks)
        widths = [(base + (1 if r < extra else 0)) * align for r in range(nranks)]
        widths[-1] += nz - sum(widths)
    else:
        base, extra = divmod(nz, nranks)
        widths = [base + (1 if r < extra else 0) for r in range(nranks)]
    out, z = [], 0
    for w in widths:
        out.append((z, z + w))
        z += w
    return out


# ------------------------------------------------------------------------------ numpy restatement of the protocol
def summary_np(mask_slab: np.ndarray) -> np.ndarray:
    """uint32 (nx, ny): lead | first<<15 | trail<<16 | last<<31 per z line of a (nx, ny, nzl) slab."""
    m = mask_slab != 0
    nzl = m.shape[2]
    first, last = m[:, :, 0], m[:, :, -1]
    diff_f = m != first[:, :, None]
    diff_l = m != last[:, :, None]
    lead = np.where(diff_f.any(axis=2), diff_f.argmax(axis=2), nzl)
    trail = np.where(diff_l.any(axis=2), diff_l[:, :, ::-1].argmax(axis=2), nzl)
    return (lead.astype(np.uint32) | (first.astype(np.uint32) << 15) | (trail.astype(np.uint32) << 16)
            | (last.astype(np.uint32) << 31))


def resolve_carries_np(all_summaries: np.ndarray, widths, rank: int):
    """(carry_lo, carry_hi) int32 (nx, ny): distance from this rank's first / last voxel of every line to the
    nearest voxel of the other polarity in lower / higher slabs (FARB if none)."""
    me = all_summaries[rank]
    myfirst, mylast = (me >> 15) & 1, (me >> 31) & 1
    shape = me.shape
    cl = np.full(shape, FARB, np.int64)
    cr = np.full(shape, FARB, np.int64)
    acc = np.zeros(shape, np.int64)
    open_ = np.ones(shape, bool)
    for k in range(rank - 1, -1, -1):
        s = all_summaries[k]
        klast, ktrail, w = (s >> 31) & 1, (s >> 16) & 0x7FFF, int(widths[k])
        hit1 = open_ & (klast != myfirst)
        cl[hit1] = acc[hit1] + 1
        open_ &= ~hit1
        hit2 = open_ & (ktrail < w)
        cl[hit2] = acc[hit2] + ktrail[hit2] + 1
        open_ &= ~hit2
        acc[open_] += w
    acc[:] = 0
    open_[:] = True
    for k in range(rank + 1, len(widths)):
        s = all_summaries[k]
        kfirst, klead, w = (s >> 15) & 1, s & 0x7FFF, int(widths[k])
        hit1 = open_ & (kfirst != mylast)
        cr[hit1] = acc[hit1] + 1
        open_ &= ~hit1
        hit2 = open_ & (klead < w)
        cr[hit2] = acc[hit2] + klead[hit2] + 1
        open_ &= ~hit2
        acc[open_] += w
    return np.minimum(cl, FARB).astype(np.int32), np.minimum(cr, FARB).astype(np.int32)


def binary_z_np(mask_slab: np.ndarray, carry_lo=None, carry_hi=None) -> np.ndarray:
    """Signed squared distance along z to the nearest voxel of the other polarity (the binary pass), with the
    optional slab carries.  int32 (nx, ny, nzl), +-INF where there is none."""
    m = mask_slab != 0
    nzl = m.shape[2]
    idx = np.arange(nzl, dtype=np.int64)
    change = np.zeros(m.shape, bool)
    change[:, :, 1:] = m[:, :, 1:] != m[:, :, :-1]
    start = np.maximum.accumulate(np.where(change, idx, -1), axis=2)                       # run start or -1
    nxt = np.where(change, idx, 10 ** 9)
    nxt = np.minimum.accumulate(nxt[:, :, ::-1], axis=2)[:, :, ::-1]
    nxt_after = np.concatenate([nxt[:, :, 1:], np.full(m.shape[:2] + (1,), 10 ** 9)], axis=2)   # first run start > z
    lo = FARB if carry_lo is None else carry_lo[:, :, None].astype(np.int64)
    hi = FARB if carry_hi is None else carry_hi[:, :, None].astype(np.int64)
    dl = np.where(start >= 1, idx - start + 1, np.where(np.asarray(lo) < FARB, idx + lo, FARB))
    dr = np.where(nxt_after < nzl, nxt_after - idx, np.where(np.asarray(hi) < FARB, nzl - 1 - idx + hi, FARB))
    d = np.minimum(dl, dr)
    d2 = np.where(d >= FARB, INF, d * d)
    return np.where(m, -d2, d2).astype(np.int32)


# ------------------------------------------------------------------------------------------------ device path
class SlabBuilder:
    """One rank of a slab-sharded build.  `group` is a torch.distributed process group (NCCL on GPUs).

    The grid is cut into world * parts z slabs and dealt to the ranks cyclically (slab k belongs to rank k % world).
    parts > 1 (SDF_B200_SLAB_PARTS) interleaves several thinner slabs per rank to even out the data-dependent cost of
    the envelope passes; measured on 8 B200 it did not pay (1024^3: 2.77 ms with 1 slab per rank vs 3.11 ms with 2;
    2048x2048x1024: 16.8 vs 17.3 ms), so the default is one slab per rank.  The exchange is one all-gather per part
    index (the summaries of world consecutive slabs); carries are resolved over all world * parts slabs."""

    def __init__(self, ctx, shape, rank: int, world: int, group=None, parts: int | None = None):
        import os
        import torch
        self.torch = torch
        self.ctx = ctx
        self.nx, self.ny, self.nz = (int(v) for v in shape)
        self.rank, self.world, self.group = rank, world, group
        if parts is None:
            parts = int(os.environ.get("SDF_B200_SLAB_PARTS", "1")) if world > 1 else 1
        while parts > 1 and self.nz // (world * parts) < 32:
            parts -= 1
        self.parts = parts
        self.nslabs = world * parts
        if self.nz > 2048:
            raise ValueError("the z axis (all slabs together) is limited to SDF_B200_MAX_AXIS = 2048 cells")
        self.ranges = slab_ranges(self.nz, self.nslabs)
        self.widths = [b - a for a, b in self.ranges]
        self.slab_ids = [p * world + rank for p in range(parts)]          # global slab index of each local part
        self.local_ranges = [self.ranges[k] for k in self.slab_ids]
        # single-part conveniences (the common N == 1 / parts == 1 case)
        self.z0, self.z1 = self.local_ranges[0]
        self.nzl = self.z1 - self.z0
        dev = torch.device("cuda", ctx.device)
        nl = self.nx * self.ny
        self.summary = torch.empty(parts * nl, dtype=torch.int32, device=dev)
        self.all_summaries = torch.empty(self.nslabs * nl, dtype=torch.int32, device=dev)
        # Exchange of the summaries: "p2p" = the summary kernel stores into every rank's gather buffer (symmetric memory
        # over NVLink, one signal-pad barrier); "nccl" = all_gather_into_tensor.  SDF_B200_EXCHANGE=nccl|p2p forces one.
        self.exchange = "nccl" if world > 1 else "none"
        self._symm = None
        want = os.environ.get("SDF_B200_EXCHANGE", "p2p")
        if world > 1 and parts == 1 and want != "nccl":
            import torch.distributed as dist
            grp = group if group is not None else dist.group.WORLD
            try:
                import torch.distributed._symmetric_memory as symm_mem
                # two gather buffers used alternately: build i+1 may store while a slow peer still reads build i
                buf = symm_mem.empty(2 * self.nslabs * nl, dtype=torch.int32, device=dev)
                hdl = symm_mem.rendezvous(buf, grp)
                self._symm_buf = buf
                self._symm = hdl
                self._peer_ptrs = [int(p) for p in hdl.buffer_ptrs]
                self._parity = 0
            except Exception as e:               # no symmetric memory on this build / topology: NCCL all-gather
                self._symm = None
                self._symm_error = repr(e)
            # every rank must take the same path: one that could not set up symmetric memory sends all of them to NCCL
            agree = torch.tensor([1 if self._symm is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=grp)
            if int(agree.item()) == 1:
                self.exchange = "p2p"
            else:
                self._symm = None

    def alloc_outputs(self):
        torch = self.torch
        dev = torch.device("cuda", self.ctx.device)
        return [torch.empty((self.nx, self.ny, b - a), dtype=torch.float32, device=dev) for a, b in self.local_ranges]

    def build(self, d_masks, resolution: float, d_sdfs):
        """d_masks / d_sdfs: uint8 / float32 device tensors (nx, ny, width) of this rank's slabs, in self.slab_ids order
        (a single tensor is accepted when the rank owns one slab)."""
        torch = self.torch
        import torch.distributed as dist
        if not isinstance(d_masks, (list, tuple)):
            d_masks, d_sdfs = [d_masks], [d_sdfs]
        ctx = self.ctx
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        nl = self.nx * self.ny
        if self._symm is not None:
            # summary + exchange in one kernel over peer memory; the barriers order (previous build's readers) ->
            # (stores of every rank) -> (this build's readers) on the current stream
            # One signal-pad barrier per build: it orders (stores of every rank into buffer `parity`) before (the
            # carry kernels that read it).  The buffer of build i is rewritten by build i+2, i.e. after the barrier of
            # build i+1, which no rank passes before its own build i has finished reading.
            off = self._parity * self.nslabs * nl * 4
            self._parity ^= 1
            ctx.slab_summary_scatter_device(d_masks[0].data_ptr(), self.nx, self.ny, self.widths[self.rank],
                                            [p + off for p in self._peer_ptrs], self.rank)
            self._symm.barrier(channel=0)
            self.all_summaries = self._symm_buf[off // 4: off // 4 + self.nslabs * nl]
            ctx.slab_build_device(d_masks[0].data_ptr(), self.nx, self.ny, self.widths[self.rank], resolution,
                                  self._symm_buf.data_ptr() + off, self.widths, self.rank, d_sdfs[0].data_ptr())
            return d_sdfs
        for p, m in enumerate(d_masks):
            w = self.widths[self.slab_ids[p]]
            ctx.slab_summary_device(m.data_ptr(), self.nx, self.ny, w, self.summary[p * nl:].data_ptr())
        if self.world > 1:
            for p in range(self.parts):     # slabs p*world .. p*world+world-1 are consecutive in the gathered buffer
                dist.all_gather_into_tensor(self.all_summaries[p * self.world * nl:(p + 1) * self.world * nl],
                                            self.summary[p * nl:(p + 1) * nl], group=self.group)
            src = self.all_summaries
        else:
            src = self.summary
        for p, (m, o) in enumerate(zip(d_masks, d_sdfs)):
            k = self.slab_ids[p]
            ctx.slab_build_device(m.data_ptr(), self.nx, self.ny, self.widths[k], resolution, src.data_ptr(),
                                  self.widths, k, o.data_ptr(), keep_extrema=(p > 0))
        return d_sdfs

    def extrema(self, resolution: float):
        """Global (max, min): all-reduce(max) of the two raw integers, then the f64 conversion."""
        torch = self.torch
        import torch.distributed as dist
        a, b = self.ctx.get_extrema_raw()              # maxima accumulated over this rank's slabs
        if self.world > 1:
            t = torch.tensor([a, b], dtype=torch.int32, device=torch.device("cuda", self.ctx.device))
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            a, b = (int(v) for v in t.tolist())
        return self.ctx.extrema_from_raw(a, b, resolution)
