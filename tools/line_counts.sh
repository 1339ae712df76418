#!/bin/bash
# Counters of the line-per-thread envelope kernels on the 1024^3 workload (SDF_B200_HULL=line forces them for both passes);
# SDF_B200_LIB selects a variant library.
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
mkdir -p gpurun_out
OUT=${1:-gpurun_out/line_counts.csv}
python tools/one_build.py c3 > /dev/null 2>&1
SDF_B200_HULL=line timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:'line_hull' -s 2 -c 2 --csv --log-file $OUT python tools/one_build.py c3 > /dev/null 2>&1
python - "$OUT" <<'PY'
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
for r in rows[1:]:
    print(r[iid], r[ik][:34], r[im], r[iv])
PY
