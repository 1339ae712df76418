"""bench_multi.py -- the N > 1 leg of bench.py: one process per GPU (torchrun), the grid cut into z slabs.

Strong scaling: the workload grid is the same for every N (BASELINE.json: "Mvox/s SDF build (1024^3 grid) at
1/2/4/8 B200"); rank r holds the z range slab_ranges(nz, N)[r].  The one data-path collective is the all-gather of
the 4-byte per-line slab summaries between the summary kernel and the binary pass (sdf_tools_b200/distributed.py).
Timing: barrier + synchronize on both sides of exactly K builds, CUDA events on the launching stream, MAX over ranks.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np


def run_multi(args, world: int, rank: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from bench import (ALGO_BYTES_PER_VOXEL, PASS_BYTES, ClockSampler, cpu_reference_leg, measured_peak)
    from sdf_tools_b200 import Context, _capi, synth
    from sdf_tools_b200.distributed import SlabBuilder

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synth.CONFIGS[args.workload]
    nx, ny, nz = cfg["shape"]
    res = cfg["res"]
    nvox = nx * ny * nz

    ctx = Context(local_rank)
    sb = SlabBuilder(ctx, (nx, ny, nz), rank, world)
    masks_h = [synth.make_mask(args.workload, zrange=r) for r in sb.local_ranges]   # this rank's slabs (cyclic deal)
    nloc = sum(m.size for m in masks_h)
    d_masks = [torch.from_numpy(m).cuda() for m in masks_h]
    d_sdfs = sb.alloc_outputs()
    stream = torch.cuda.current_stream()

    def step():
        sb.build(d_masks, res, d_sdfs)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    per_pass = {k: [] for k in PASS_BYTES}
    ctx.set_stream(stream.cuda_stream)
    for _ in range(min(args.steps, 5)):                  # per-kernel times: build the rank's slabs one by one
        acc = {k: 0.0 for k in per_pass}
        sb.build(d_masks, res, d_sdfs)                   # (fills the gathered summaries)
        for p_i, (m_, o_) in enumerate(zip(d_masks, d_sdfs)):
            k_ = sb.slab_ids[p_i]
            ctx.slab_build_device(m_.data_ptr(), nx, ny, sb.widths[k_], res, sb.all_summaries.data_ptr() if world > 1
                                  else sb.summary.data_ptr(), sb.widths, k_, o_.data_ptr(), keep_extrema=(p_i > 0))
            tm = ctx.last_timing()
            for k in per_pass:
                acc[k] += tm[k]
        for k in per_pass:
            per_pass[k].append(acc[k])
    clocks = sampler.stop() if sampler else None
    # per-rank kernel time (sum of the three passes): shows the data-dependent load imbalance between slabs
    mine = torch.tensor([float(np.mean([sum(per_pass[k][j] for k in per_pass) for j in range(len(per_pass["pass1_z_ms"]))]))],
                        dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    per_rank_ms = [round(float(t.item()), 4) for t in allr]
    mx, mn = sb.extrema(res)
    launches = (ctx.last_launches() + 1) * sb.parts * args.steps   # per slab: summary + carry + 3 passes

    # ------------------------------------------------------------------------------------------- end to end
    e2e = None
    if not args.no_e2e:
        lib = _capi.load()
        h_in, h_out = C.c_void_p(), C.c_void_p()
        assert lib.sdf_b200_host_alloc(C.byref(h_in), nloc) == 0
        assert lib.sdf_b200_host_alloc(C.byref(h_out), nloc * 4) == 0
        flat_in = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(nloc,))
        flat_out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(nloc,))
        pin_masks, pin_outs, off = [], [], 0
        for m in masks_h:                                         # pinned host views, one per local slab
            pin_masks.append(torch.from_numpy(flat_in[off:off + m.size].reshape(m.shape)))
            pin_outs.append(torch.from_numpy(flat_out[off:off + m.size].reshape(m.shape)))
            pin_masks[-1].copy_(torch.from_numpy(m))
            off += m.size

        def e2e_step():
            for dm, pm in zip(d_masks, pin_masks):
                dm.copy_(pm, non_blocking=True)                   # H2D of this rank's slabs
            sb.build(d_masks, res, d_sdfs)
            for po, ds in zip(pin_outs, d_sdfs):
                po.copy_(ds, non_blocking=True)                   # D2H of this rank's slabs
            torch.cuda.synchronize()
            return sb.extrema(res)

        e2e_step()
        dist.barrier()
        n_e2e = max(1, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": nvox / e2e_ms / 1e3, "unit": "Mvox/s", "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_bytes_per_step": nvox, "d2h_bytes_per_step": nvox * 4 + 8 * world,
               "api": "sdf_tools_b200.distributed.SlabBuilder.build (pinned host slab per rank -> H2D -> summary -> "
                      "all-gather -> 3 passes -> D2H)"}
        lib.sdf_b200_host_free(h_in)
        lib.sdf_b200_host_free(h_out)

    verified = None
    if args.verify:                                       # not timed: every rank's slabs against the CPU oracle
        from oracle import oracle as O
        step()
        torch.cuda.synchronize()
        full = synth.make_mask(args.workload)
        want, wmx, wmn = O.build_sdf(full, res)
        ok = all(np.array_equal(o.cpu().numpy().view(np.uint32), want[:, :, a:b_].view(np.uint32))
                 for o, (a, b_) in zip(d_sdfs, sb.local_ranges)) and (mx, mn) == (wmx, wmn)
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified = bool(t.item())
        assert verified, f"rank {rank}: slab build differs from the oracle"

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = nvox / ms_per_step / 1e3
        peak, peak_src = measured_peak()
        means = {k: float(np.mean(v)) for k, v in per_pass.items()}
        dom = max(means, key=lambda k: means[k])
        vox_rank = nloc
        dom_gbs = PASS_BYTES[dom] * vox_rank / (means[dom] * 1e-3) / 1e9
        whole_gbs = ALGO_BYTES_PER_VOXEL * nvox / (ms_per_step * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            full = synth.make_mask(args.workload, zrange=(0, min(args.cpu_edge, nz)))
            cpu = cpu_reference_leg(full, res, args.cpu_edge)
        out = {
            "metric": "Mvox/s SDF build", "value": value, "unit": "Mvox/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32 d2 -> f32 (exact f64 sqrt table)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['desc']}", "grid": [nx, ny, nz], "resolution": res,
                       "sharding": f"{sb.nslabs} z slabs of {sb.widths[0]} planes dealt cyclically, {sb.parts} per rank",
                       "collective": f"{sb.parts} all-gather(s) of 4 B per z line per build "
                       f"({nx * ny * 4} B per rank each) + all-reduce of 2 ints for the extrema",
                       "l2": "per-rank inputs larger than L2; no flush needed", "extrema": [mx, mn]},
            "roofline": {"bound": "hbm", "kernel": {"pass1_z_ms": "binary_z_kernel", "pass2_y_ms": "band_strided_kernel (y)",
                                                    "pass3_x_ms": "band_strided_kernel (x, final)"}[dom],
                         "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "traffic": None,
                         "peak_source": peak_src, "per_gpu": True, "algorithmic_bytes_per_voxel": PASS_BYTES[dom],
                         "kernel_ms": means[dom],
                         "whole_build": {"achieved": whole_gbs, "peak": peak * world, "frac": whole_gbs / (peak * world),
                                         "algorithmic_bytes_per_voxel": ALGO_BYTES_PER_VOXEL},
                         "per_pass_ms": means, "per_rank_kernel_ms": per_rank_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "verified_vs_oracle": verified,
        }
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
