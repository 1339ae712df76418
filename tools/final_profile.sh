#!/bin/bash
# Round-end measurement set (single GPU): bench line, reference arm, ncu launch list of the bench command, counters,
# one full capture of the three build kernels, per-phase source counters, traffic file tied to the kernel source hash.
R=${1:-r2}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-c4 --no-same-config > gpurun_out/b_ncu.log 2>&1
timeout 300 tools/ncu_counts.sh
cp gpurun_out/counts_libsdf_b200.csv gpurun_out/${R}_counters.csv
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:'binary_z|band_strided|line_hull' -s 3 -c 3 -o gpurun_out/${R}_prof -f python tools/one_build.py c3 > gpurun_out/${R}_ncu_full.log 2>&1
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.csv
lscpu | head -20 > gpurun_out/${R}_lscpu.txt
python - <<P
import json
d=json.load(open("gpurun_out/${R}_bench_n1.json"))
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
P
ls -la gpurun_out | tail -8
