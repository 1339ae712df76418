"""bench_multi.py -- the N > 1 leg of bench.py: one process per GPU (torchrun), the grid cut into z slabs.

Strong scaling: the workload grid is the same for every N (BASELINE.json: "Mvox/s SDF build (1024^3 grid) at
1/2/4/8 B200"); rank r holds the z range slab_ranges(nz, N)[r].  The one data-path exchange is that of the 4-byte
per-line slab summaries before the binary pass (sdf_tools_b200/distributed.py): by default the summary kernel stores
them straight into every rank's gather buffer over NVLink (symmetric memory) and one signal-pad barrier orders the
builds behind it; SDF_B200_EXCHANGE=nccl uses an NCCL all-gather instead.
Timing: barrier + synchronize on both sides of exactly K builds, CUDA events on the launching stream, MAX over ranks.
Every run checks the assembled result against a single-GPU build of the same grid on rank 0 (not timed;
`verified_vs_single_gpu`), and against the CPU oracle when the grid is small enough or --verify is given.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import time

import numpy as np


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and the pinned host buffers it will allocate) to the NUMA node of its GPU: with 8 ranks on a
    two-socket host, unbound ranks push half of their PCIe traffic across the socket link."""
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = vis.split(",")[local_rank].strip() if vis else str(local_rank)
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", phys],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]                                   # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def time_slab_builds(torch, dist, sb, d_masks, d_sdfs, res, steps, warmup, stream, sampler=None):
    def step():
        sb.build(d_masks, res, d_sdfs)
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    if sampler:
        sampler.start()                                   # clocks are sampled during the timed builds only
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def run_multi(args, world: int, rank: int, local_rank: int):
    numa = bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    from bench import (ALGO_BYTES_PER_VOXEL, PASS_BYTES, PASS_KERNEL, ClockSampler, cpu_reference_leg, crop_for_cpu,
                       measured_peak)
    from sdf_tools_b200 import Context, _capi, synth
    from sdf_tools_b200.distributed import SlabBuilder

    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = synth.CONFIGS[args.workload]
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else cfg["shape"]
    nx, ny, nz = shape
    res = cfg["res"]
    nvox = nx * ny * nz

    ctx = Context(local_rank)
    sb = SlabBuilder(ctx, (nx, ny, nz), rank, world)
    masks_h = [synth.make_mask(args.workload, shape if args.shape else None, zrange=r) for r in sb.local_ranges]
    nloc = sum(m.size for m in masks_h)
    d_masks = [torch.from_numpy(m).cuda() for m in masks_h]
    d_sdfs = sb.alloc_outputs()
    stream = torch.cuda.current_stream()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step = time_slab_builds(torch, dist, sb, d_masks, d_sdfs, res, args.steps, args.warmup, stream, sampler)
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
    sb.build(d_masks, res, d_sdfs)                       # leave a complete build behind for the checks below
    torch.cuda.synchronize()
    # per-rank kernel time (sum of the three passes): shows the data-dependent load imbalance between slabs
    mine = torch.tensor([float(np.mean([sum(per_pass[k][j] for k in per_pass) for j in range(len(per_pass["pass1_z_ms"]))]))],
                        dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    per_rank_ms = [round(float(t.item()), 4) for t in allr]
    mx, mn = sb.extrema(res)
    launches = (ctx.last_launches() + 1) * sb.parts * args.steps   # per slab: summary(+exchange) + carry + 3 passes

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
               "h2d_bytes_per_step": nvox, "d2h_bytes_per_step": nvox * 4 + 8 * world, "numa_node_of_rank0": numa,
               "api": "sdf_tools_b200.distributed.SlabBuilder.build (pinned host slab per rank -> H2D -> summary + exchange "
                      "-> 3 passes -> D2H), every rank bound to its GPU's NUMA node"}
        lib.sdf_b200_host_free(h_in)
        lib.sdf_b200_host_free(h_out)

    # ------------------------------------------------------------------------------- checks (not timed)
    verified_single = None
    if not args.no_verify:
        # every rank's slabs against ONE single-GPU build of the whole grid (itself checked against the CPU oracle at this
        # size by tests/test_gpu_parity.py): rank 0 builds it and scatters the reference slabs
        ok = True
        if rank == 0:
            full = torch.from_numpy(synth.make_mask(args.workload, shape if args.shape else None)).cuda()
            want = torch.empty((nx, ny, nz), dtype=torch.float32, device="cuda")
            c1 = Context(local_rank)
            c1.set_stream(stream.cuda_stream)
            c1.build_device(full.data_ptr(), nx, ny, nz, res, want.data_ptr())
            torch.cuda.synchronize()
            ok = c1.get_extrema() == (mx, mn)
            del full
        for src in range(world):                                   # slab of rank `src`: rank 0 -> src, compared there
            for p_i, (a, b_) in enumerate([sb.ranges[p * world + src] for p in range(sb.parts)]):
                if rank == 0:
                    ref = want[:, :, a:b_].contiguous()
                    if src == 0:
                        ok = ok and bool(torch.equal(ref.view(torch.int32), d_sdfs[p_i].view(torch.int32)))
                    else:
                        dist.send(ref, dst=src)
                elif rank == src:
                    ref = torch.empty_like(d_sdfs[p_i])
                    dist.recv(ref, src=0)
                    ok = ok and bool(torch.equal(ref.view(torch.int32), d_sdfs[p_i].view(torch.int32)))
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified_single = bool(t.item())
        if rank == 0:
            del want
        assert verified_single, f"rank {rank}: slab build differs from the single-GPU build"
    verified = None
    if args.verify or nvox <= 1 << 28:                             # the CPU oracle itself (all cores), small grids by default
        from oracle import oracle as O
        full = synth.make_mask(args.workload, shape if args.shape else None)
        want_h, wmx, wmn = O.build_sdf(full, res)
        ok = all(np.array_equal(o.cpu().numpy().view(np.uint32), want_h[:, :, a:b_].view(np.uint32))
                 for o, (a, b_) in zip(d_sdfs, sb.local_ranges)) and (mx, mn) == (wmx, wmn)
        t = torch.tensor([int(ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified = bool(t.item())
        del full, want_h
        assert verified, f"rank {rank}: slab build differs from the oracle"

    # ------------------------------------------------------------------ the north_star's scaling config, same ranks
    c4 = None
    if not args.no_c4 and args.workload == "c3" and not args.shape:
        del d_masks, d_sdfs
        torch.cuda.empty_cache()
        shape4 = synth.CONFIGS["c4"]["shape"]
        sb4 = SlabBuilder(ctx, shape4, rank, world)
        m4 = [torch.from_numpy(synth.make_mask("c4", zrange=r)).cuda() for r in sb4.local_ranges]
        o4 = sb4.alloc_outputs()
        ms4 = time_slab_builds(torch, dist, sb4, m4, o4, synth.CONFIGS["c4"]["res"], 5, 3, stream)
        n4 = int(np.prod(shape4))
        peak4, _ = measured_peak()
        c4 = {"workload": f"c4: {synth.CONFIGS['c4']['desc']}", "grid": list(shape4), "n_gpus": world, "ms_per_step": ms4,
              "value": n4 / ms4 / 1e3, "unit": "Mvox/s", "exchange": sb4.exchange,
              "whole_build_frac_per_gpu": ALGO_BYTES_PER_VOXEL * n4 / (ms4 * 1e-3) / 1e9 / (peak4 * world)}
        del m4, o4

    if rank == 0:
        value = nvox / ms_per_step / 1e3
        peak, peak_src = measured_peak()
        means = {k: float(np.mean(v)) for k, v in per_pass.items()}
        dom = max(means, key=lambda k: means[k])
        dom_gbs = PASS_BYTES[dom] * nloc / (means[dom] * 1e-3) / 1e9
        whole_gbs = ALGO_BYTES_PER_VOXEL * nvox / (ms_per_step * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            full = synth.make_mask(args.workload, shape if args.shape else None, zrange=(0, min(args.cpu_edge, nz)))
            cpu = cpu_reference_leg(crop_for_cpu(full, args.cpu_edge), res,
                                    "corner crop of the workload grid (the reference cannot hold the full grid)")
        out = {
            "metric": "Mvox/s SDF build", "value": value, "unit": "Mvox/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32 d2 -> f32 (exact f64 sqrt table)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['desc']}", "grid": [nx, ny, nz], "resolution": res,
                       "sharding": f"{sb.nslabs} z slabs of {sb.widths[0]} planes dealt cyclically, {sb.parts} per rank",
                       "exchange": sb.exchange,
                       "collective": ("summary kernel stores 4 B per z line into every rank's gather buffer over NVLink "
                                      "(symmetric memory) + 1 signal-pad barrier per build" if sb.exchange == "p2p" else
                                      f"{sb.parts} NCCL all-gather(s) of 4 B per z line per build ({nx * ny * 4} B per rank each)")
                       + "; all-reduce of 2 ints for the extrema (outside the timed builds)",
                       "l2": "per-rank inputs larger than L2; no flush needed", "extrema": [mx, mn]},
            "roofline": {"bound": "hbm", "kernel": PASS_KERNEL[dom],
                         "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak, "traffic": None,
                         "peak_source": peak_src, "per_gpu": True, "algorithmic_bytes_per_voxel": PASS_BYTES[dom],
                         "kernel_ms": means[dom],
                         "whole_build": {"achieved": whole_gbs, "peak": peak * world, "frac": whole_gbs / (peak * world),
                                         "algorithmic_bytes_per_voxel": ALGO_BYTES_PER_VOXEL},
                         "per_pass_ms": means, "per_rank_kernel_ms": per_rank_ms},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "verified_vs_single_gpu": verified_single, "verified_vs_oracle": verified, "extra": {"c4": c4},
        }
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
