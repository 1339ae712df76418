#!/bin/bash
# Round-end measurement set (single GPU): bench line, ncu launch list of the same command, counters, one full capture.
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 200 tools/ncu_counts.sh
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name regex:band_strided -c 2 -o gpurun_out/r1_prof -f python tools/one_build.py c3 > gpurun_out/r1_ncu_full.log 2>&1
python - <<'P'
import json
d=json.load(open("gpurun_out/r1_bench_n1.json"))
print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
P
ls -la gpurun_out | tail -8
