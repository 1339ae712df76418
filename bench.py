#!/usr/bin/env python
"""bench.py -- Mvox/s of the SDF build (BASELINE.json metric) on N B200s, one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]

A "step" is one full build of the workload grid: uint8 occupancy mask -> f32 signed distance field + (max, min),
i.e. one call of sdf_generation::ExtractSignedDistanceField (reference sdf_generation.hpp:209-271).
  value     device-resident: mask already in HBM, SDF left in HBM; CUDA-event time of exactly K builds.
  e2e       the same builds through the host-buffer C-ABI entry point (sdf_b200_build_from_mask): pinned host mask
            -> H2D -> 3 kernels -> D2H of the f32 field + extrema, every step.
  e2e_api   the call the reference's users make: pysdf_tools.CollisionMapGrid.ExtractSignedDistanceField (8-byte cells
            in, SignedDistanceField object out) on the C5 grid (and on C3 with --e2e-api-c3).
  roofline  dominant kernel's algorithmic bytes / its mean CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs; `traffic`
            only when the committed ncu capture was taken from the same kernel sources (hash), else null.
  cpu_baseline  the reference's own code (oracle/_ref) on a bounded crop of the same workload, 1 thread (the reference
            is single-threaded), plus the exact separable EDT (oracle port, OpenMP, all host cores) on a 512^3 crop.
  same_config   ours on the C2 256^3 sphere grid -- the config `--impl reference` runs WHOLE (the reference cannot
            hold 1024^3: ~108 B/voxel), so one same-grid GPU / reference ratio exists.
  extra.c4  the north_star's scaling config (2048x2048x1024) on this number of GPUs (skipped with --no-c4).
Inputs (1 GiB mask, 4 GiB field at 1024^3) are far larger than the 126 MB L2, so no L2 flush is needed.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_VOXEL = 21.0          # 1+4 (pass 1) + 4+4 (pass 2) + 4+4 (pass 3)
PASS_BYTES = {"pass1_z_ms": 5.0, "pass2_y_ms": 8.0, "pass3_x_ms": 8.0}
PASS_KERNEL = {"pass1_z_ms": "binary_z_kernel", "pass2_y_ms": "band_strided_kernel (y)",
               "pass3_x_ms": "band_strided_kernel (x, final)"}


def kernel_source_hash() -> str:
    """sha256 over the CUDA sources of the build kernels with comments and white space removed: ties an ncu capture to the
    CODE it was taken from (editing a comment does not invalidate a measurement, editing a statement does)."""
    import re
    h = hashlib.sha256()
    for f in ("edt_kernels.cu", "band_hull.cuh", "line_kernels.cu", "line_hull.cuh"):
        with open(os.path.join(ROOT, "sdf_tools_b200", "csrc", f), "r") as fh:
            src = fh.read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # block comments
        src = re.sub(r"//[^\n]*", "", src)                        # line comments (no string literal in these files holds "//")
        h.update(re.sub(r"\s+", "", src).encode())
    return h.hexdigest()[:16]


def ncu_traffic(pass_key: str, nvox: int):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture -- only for the same
    grid AND the same kernel sources; otherwise None (a stale profile must not pose as a measurement)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        if d.get("kernel_source_hash") != kernel_source_hash() or nvox != 1024 ** 3:
            return None
        return float(d[pass_key]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled every 20 ms DURING the timed region: NVML in a thread (the timed region of a
    15 ms build x K steps is too short for `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []          # (sm_mhz, max_mhz, set of reasons)
        self.idx = gpu_index
        self.stop_flag = False
        self.thread = None
        self.proc = None
        self.how = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        try:
            return int(vis.split(",")[self.idx]) if vis else self.idx
        except Exception:
            return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}

            def loop():
                while not self.stop_flag:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((float(sm), float(mx), {k for k, b in bits.items() if r & b}))
                    except Exception:
                        pass
                    time.sleep(0.02)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.how = "nvml, 20 ms"
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi -lms 50"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                reasons = {n for n, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                            ("sw_power_cap", 8)) if len(r) > col and r[col].lower().startswith("active")}
                self.rows.append((float(r[1]), float(r[2]), reasons))
            except Exception:
                continue

    def stop(self):
        if self.thread is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML / nvidia-smi unavailable)"]}
        self.stop_flag = True
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        elif self.thread:
            self.thread.join(timeout=1)
        sm = [r[0] for r in self.rows]
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None,
                "samples": len(sm), "source": self.how, "reasons": sorted(reasons)}


def crop_for_cpu(mask: np.ndarray, edge: int) -> np.ndarray:
    """Bounded sample of the workload for the CPU legs: the edge^3 corner that holds the ground plane."""
    e = [min(edge, s) for s in mask.shape]
    return np.ascontiguousarray(mask[:e[0], :e[1], :e[2]])


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_reference_leg(sample: np.ndarray, res: float, what: str):
    """Times the reference's own ExtractSignedDistanceField<T> (oracle/_ref; oracle port if _ref is absent)."""
    from oracle import oracle as O
    nvox = sample.size
    if O.have_ref():
        _, _, _, ms = O.ref_build_sdf(sample, res)
        kind, cores = "reference", 1
    else:
        t0 = time.perf_counter()
        O.build_sdf(sample, res, nthreads=0)
        ms = (time.perf_counter() - t0) * 1e3
        kind, cores = "port", os.cpu_count() or 1
    return {"value": nvox / ms / 1e3, "unit": "Mvox/s", "cores": cores, "kind": kind, "ms": ms,
            "sample": f"{'x'.join(map(str, sample.shape))} {what}, 1 build", "host_cpu": _cpu_model(),
            "host_cores": os.cpu_count()}


def cpu_exact_edt_leg(sample: np.ndarray, res: float, what: str):
    """The same ALGORITHM on the CPU: exact separable EDT x2 + combine (oracle/edt_oracle.c), OpenMP over all cores."""
    from oracle import oracle as O
    O.build_sdf(crop_for_cpu(sample, 64), res, nthreads=0)                 # thread pool warm-up
    t0 = time.perf_counter()
    O.build_sdf(sample, res, nthreads=0)
    ms = (time.perf_counter() - t0) * 1e3
    return {"value": sample.size / ms / 1e3, "unit": "Mvox/s", "cores": os.cpu_count() or 1, "kind": "port", "ms": ms,
            "sample": f"{'x'.join(map(str, sample.shape))} {what}, 1 build, exact separable EDT (OpenMP)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, rank 0 only.  Default: the C2 256^3 sphere grid WHOLE
    (BASELINE.json configs[1]; the reference needs ~108 B/voxel and ~35 s for it), which `ours` also reports under
    `same_config`.  --workload c3 times a corner crop of the C3 grid instead (the full grid does not fit)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sdf_tools_b200 import synth
    wl = args.ref_workload
    cfg = synth.CONFIGS[wl]
    if int(np.prod(cfg["shape"])) <= 320 ** 3:
        sample, what = synth.make_mask(wl), "whole grid"
    else:
        sample, what = crop_for_cpu(synth.make_mask(wl), args.cpu_edge), "corner crop of the workload grid (the reference cannot hold the full grid)"
    times, leg = [], None
    for i in range(args.warmup + args.steps):
        leg = cpu_reference_leg(sample, cfg["res"], what)
        if i >= args.warmup:
            times.append(leg["ms"])
    ms = float(np.mean(times))
    val = sample.size / ms / 1e3
    leg["value"], leg["ms"] = val, ms
    print(json.dumps({
        "impl": "reference", "metric": "Mvox/s SDF build", "value": val, "unit": "Mvox/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32 d2 (f64 in the reference) -> f32", "data": "synthetic",
        "config": {"workload": f"{wl}: {cfg['desc']}", "grid": list(sample.shape), "sample": leg["sample"],
                   "note": "compare with the `same_config` block of the GPU arm (same grid, whole)"},
        "cpu_baseline": leg,
        "e2e": {"value": val, "unit": "Mvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_device_builds(ctx, torch, stream, mask_h, res, steps, warmup, sampler=None):
    """CUDA-event time of `steps` device-resident builds; returns (ms_per_step, per-pass means, extrema, launches, d_sdf)."""
    nx, ny, nz = mask_h.shape
    d_mask = torch.from_numpy(mask_h).cuda()
    d_sdf = torch.empty((nx, ny, nz), dtype=torch.float32, device="cuda")

    def step():
        ctx.build_device(d_mask.data_ptr(), nx, ny, nz, res, d_sdf.data_ptr())

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()                                   # clocks are sampled during the timed builds only
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    per_pass = {k: [] for k in PASS_BYTES}
    for _ in range(steps):
        step()
        t = ctx.last_timing()
        for k in per_pass:
            per_pass[k].append(t[k])
    return total_ms / steps, {k: float(np.mean(v)) for k, v in per_pass.items()}, ctx.get_extrema(), ctx.last_launches(), d_mask, d_sdf


def e2e_host_leg(ctx, torch, mask_h, res, steps):
    """Host-buffer C ABI, pinned buffers, H2D + build + D2H + extrema inside the timed region."""
    from sdf_tools_b200 import _capi
    import ctypes as C
    lib = _capi.load()
    nx, ny, nz = mask_h.shape
    nvox = mask_h.size
    h_in, h_out = C.c_void_p(), C.c_void_p()
    assert lib.sdf_b200_host_alloc(C.byref(h_in), nvox) == 0
    assert lib.sdf_b200_host_alloc(C.byref(h_out), nvox * 4) == 0
    pin_mask = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(nx, ny, nz))
    pin_out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(nx, ny, nz))
    pin_mask[...] = mask_h
    ctx.build_from_mask(pin_mask, res, out=pin_out)           # warm-up (allocates the staging buffers)
    torch.cuda.synchronize()
    n = max(1, min(steps, 5))
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.build_from_mask(pin_mask, res, out=pin_out)       # returns after D2H + extrema readback
    ms = (time.perf_counter() - t0) * 1e3 / n
    lib.sdf_b200_host_free(h_in)
    lib.sdf_b200_host_free(h_out)
    return {"value": nvox / ms / 1e3, "unit": "Mvox/s", "ms_per_step": ms, "steps": n, "h2d_bytes_per_step": nvox,
            "d2h_bytes_per_step": nvox * 4 + 8, "api": "sdf_b200_build_from_mask (C ABI, pinned host buffers)"}


def e2e_api_leg(workload: str, steps: int):
    """pysdf_tools.CollisionMapGrid.ExtractSignedDistanceField: 8-byte COLLISION_CELLs of a C++ std::vector in, a
    SignedDistanceField (std::vector<float>) out -- pageable memory on both sides, as the reference's callers have it."""
    from sdf_tools_b200 import pysdf_tools as P, synth
    cfg = synth.CONFIGS[workload]
    nx, ny, nz = cfg["shape"]
    occ = synth.make_mask(workload).astype(np.float32)
    tf = P.Isometry3d([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    grid = P.CollisionMapGrid(tf, "world", cfg["res"], nx, ny, nz, P.COLLISION_CELL(0.0))
    grid.SetOccupancy(occ)
    del occ
    grid.ExtractSignedDistanceField(0.0, False, False)                    # warm-up
    n = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        sdf, ext = grid.ExtractSignedDistanceField(0.0, False, False)
    ms = (time.perf_counter() - t0) * 1e3 / n
    nvox = nx * ny * nz
    return {"workload": workload, "value": nvox / ms / 1e3, "unit": "Mvox/s", "ms_per_step": ms, "steps": n,
            "h2d_bytes_per_step": nvox * 8, "d2h_bytes_per_step": nvox * 4 + 8,
            "api": "pysdf_tools.CollisionMapGrid.ExtractSignedDistanceField (pageable std::vector in / out)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--shape", default=None, help="nx,ny,nz: run the workload's generator at another grid size (tests)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-workload", default="c2", choices=["c1", "c2", "c3", "c5"],
                    help="--impl reference: the grid the reference's CPU code is timed on (default: C2 256^3, whole)")
    ap.add_argument("--cpu-edge", type=int, default=320, help="edge of the CPU-baseline crop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the extra 2048x2048x1024 timing")
    ap.add_argument("--no-same-config", action="store_true", help="skip the C2 block that matches --impl reference")
    ap.add_argument("--e2e-api-c3", action="store_true", help="also time the pybind API on the 1024^3 grid (8.6 GB of cells)")
    ap.add_argument("--verify", action="store_true", help="N > 1: compare every rank's slabs with the CPU oracle (any grid size)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the default check against a single-GPU build")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        if args.steps > 2:
            args.steps = 2            # each step is a bounded CPU run of ~10-40 s
        args.warmup = min(args.warmup, 1)
        return run_reference(args)

    import torch
    from sdf_tools_b200 import Context, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from bench_multi import run_multi   # slab-sharded build (one process per GPU)
        return run_multi(args, world, rank, local_rank)

    torch.cuda.set_device(0)
    cfg = synth.CONFIGS[args.workload]
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else cfg["shape"]
    nx, ny, nz = shape
    res = cfg["res"]
    nvox = nx * ny * nz
    mask_h = synth.make_mask(args.workload, shape if args.shape else None)
    fill = float(mask_h.mean())

    ctx = Context(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    sampler = ClockSampler(0)
    ms_per_step, means, (mx, mn), launches, d_mask, d_sdf = time_device_builds(ctx, torch, stream, mask_h, res, args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    launches *= args.steps
    del d_mask, d_sdf
    torch.cuda.empty_cache()

    value = nvox / ms_per_step / 1e3          # Mvox/s
    peak, peak_src = measured_peak()
    dom = max(means, key=lambda k: means[k])
    dom_gbs = PASS_BYTES[dom] * nvox / (means[dom] * 1e-3) / 1e9
    whole_gbs = ALGO_BYTES_PER_VOXEL * nvox / (ms_per_step * 1e-3) / 1e9
    env_kernel = ctx.last_envelope_kernel()
    pass_kernel = dict(PASS_KERNEL)
    if env_kernel in ("line", "mixed"): pass_kernel["pass2_y_ms"] = "line_hull_kernel (y)"
    if env_kernel == "line": pass_kernel["pass3_x_ms"] = "line_hull_kernel (x, final)"
    roofline = {"bound": "hbm", "kernel": pass_kernel[dom], "envelope_kernels": env_kernel,
                "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": ncu_traffic(dom, nvox), "algorithmic_bytes_per_launch": PASS_BYTES[dom] * nvox,
                "peak_source": peak_src, "algorithmic_bytes_per_voxel": PASS_BYTES[dom],
                "kernel_ms": means[dom], "kernel_source_hash": kernel_source_hash(),
                "whole_build": {"achieved": whole_gbs, "frac": whole_gbs / peak,
                                "algorithmic_bytes_per_voxel": ALGO_BYTES_PER_VOXEL},
                "per_pass_ms": means}

    e2e = None if args.no_e2e else e2e_host_leg(ctx, torch, mask_h, res, args.steps)

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_leg(crop_for_cpu(mask_h, args.cpu_edge), res,
                                "corner crop of the workload grid (the reference needs ~108 B/voxel and cannot hold the full grid)")
        cpu["exact_edt_openmp"] = cpu_exact_edt_leg(crop_for_cpu(mask_h, 512), res, "corner crop of the workload grid")
    del mask_h

    extra = {}
    if not args.no_e2e:
        try:
            extra["e2e_api"] = [e2e_api_leg("c5", args.steps)]
            if args.e2e_api_c3:
                extra["e2e_api"].append(e2e_api_leg("c3", args.steps))
        except Exception as e:                                   # the pybind module is optional for the bench
            extra["e2e_api"] = {"error": repr(e)}
    if not args.no_same_config and args.workload != "c2":
        m2 = synth.make_mask("c2")
        r2 = synth.CONFIGS["c2"]["res"]
        ms2, means2, _, _, dm, ds = time_device_builds(ctx, torch, stream, m2, r2, max(args.steps, 10), 3)
        del dm, ds
        e2 = None if args.no_e2e else e2e_host_leg(ctx, torch, m2, r2, args.steps)
        extra["same_config"] = {"workload": f"c2: {synth.CONFIGS['c2']['desc']}", "grid": list(m2.shape), "ms_per_step": ms2,
                                "value": m2.size / ms2 / 1e3, "unit": "Mvox/s", "per_pass_ms": means2, "e2e": e2,
                                "note": "the grid `bench.py --impl reference` runs whole"}
    if not args.no_c4 and args.workload == "c3" and not args.shape:
        m4 = synth.make_mask("c4")
        ms4, means4, _, _, dm, ds = time_device_builds(ctx, torch, stream, m4, synth.CONFIGS["c4"]["res"], 5, 3)
        del dm, ds
        extra["c4"] = {"workload": f"c4: {synth.CONFIGS['c4']['desc']}", "grid": list(m4.shape), "n_gpus": 1,
                       "ms_per_step": ms4, "value": m4.size / ms4 / 1e3, "unit": "Mvox/s", "per_pass_ms": means4,
                       "whole_build_frac": ALGO_BYTES_PER_VOXEL * m4.size / (ms4 * 1e-3) / 1e9 / peak}
        del m4

    out = {
        "metric": "Mvox/s SDF build", "value": value, "unit": "Mvox/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32 d2 -> f32 (exact f64 sqrt table)", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['desc']}", "grid": [nx, ny, nz], "resolution": res,
                   "fill": fill, "l2": "inputs (1 B + 4 B per voxel) larger than L2; no flush needed",
                   "extrema": [mx, mn]},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extra": extra,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
