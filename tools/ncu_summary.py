"""tools/ncu_summary.py -- turn an `ncu --set full` report into the markdown summary kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/r1_prof.ncu-rep [extra_counts.csv ...] > profiles/r1_ncu_summary.md
Extra CSVs are `ncu --metrics ... --csv` logs (tools/ncu_counts.sh) for kernels the full capture skipped."""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"]
ALGO = {"binary_z_kernel": 5, "band_strided_kernel": 8, "line_hull_kernel": 8}          # algorithmic bytes per voxel
NVOX = 1 << 30


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    print(f"# ncu summary (`ncu --set full --clock-control none --import-source on`, one capture per kernel of one 1024^3 build, B200)\n")
    print(f"Source: {rep} (scratch; regenerate with tools/final_profile.sh + tools/ncu_summary.py).")
    print("Times under ncu are cold-cache and serialised; bench values come from CUDA events in bench.py (profiles/*_bench_n1.json).\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        print(f"## {name}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEEP:
            if k in d:
                print(f"| {k} | {d[k]} | {u[k]} |")
        rd = to_bytes(d["dram__bytes_read.sum"].replace(",", ""), u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"].replace(",", ""), u["dram__bytes_write.sum"])
        per = next((v for k, v in ALGO.items() if k in name), None)
        if per:
            algo = per * NVOX
            print(f"| traffic (dram read+write) | {(rd + wr) / 1e9:.3f} | GB |")
            print(f"| algorithmic bytes ({per} B/voxel x 2^30) | {algo / 1e9:.3f} | GB |")
            print(f"| traffic / algorithmic | {(rd + wr) / algo:.3f} | |")
        stalls = {k.split("smsp__pcsamp_warps_issue_stalled_")[1]: float(v.replace(",", "")) for k, v in d.items()
                  if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "n/a")}
        tot = sum(stalls.values()) or 1
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:8]
        print("\nwarp stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top) + "\n")
        traffic[name] = rd + wr
    for extra in sys.argv[2:]:
        rows = [r for r in csv.reader(open(extra)) if len(r) > 10]
        h = rows[0]; ki, mi, ui, vi, ii = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
        per_k = {}
        for r in rows[1:]:
            per_k.setdefault((r[ii], r[ki]), []).append((r[mi], r[vi], r[ui]))
        print(f"## counters from {extra} (`ncu --metrics ...`, tools/ncu_counts.sh)\n")
        for (_, name), ms in per_k.items():
            print(f"### {name}\n\n| metric | value | unit |\n|---|---|---|")
            for m, v, un in ms:
                print(f"| {m} | {v} | {un} |")
            print()
            dd = {m: v for m, v, _ in ms}
            if "dram__bytes_read.sum" in dd and name not in traffic:
                traffic[name] = float(dd["dram__bytes_read.sum"].replace(",", "")) + float(dd["dram__bytes_write.sum"].replace(",", ""))
    # traffic file read by bench.py: only valid for the kernel sources it was measured on (hash)
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_hash
    out = {"kernel_source_hash": kernel_source_hash(), "grid": "1024x1024x1024", "source": f"{rep} (ncu --set full, one launch per kernel)"}
    for name, b in traffic.items():
        if "binary_z" in name: key = "pass1_z_ms"
        elif "line_hull" in name:                    # line_hull_kernel<WIDE, FINAL, TPB>
            flags = name.split("<")[1].split(">")[0].split(",")
            key = "pass3_x_ms" if flags[1].strip() in ("(bool)1", "1", "true") else "pass2_y_ms"
        else: key = "pass3_x_ms" if ("(bool)1" in name or ", 1>" in name) else "pass2_y_ms"    # band_strided_kernel<.., FINAL>
        out[key] = {"kernel": name.split("(")[0], "dram_bytes_per_launch": b}
    json.dump(out, open("/tmp/ncu_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main()
