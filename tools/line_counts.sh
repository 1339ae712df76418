M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__thread_inst_executed_per_inst_executed.ratio
mkdir -p gpurun_out
python tools/one_build.py c3 > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none --kernel-name regex:'line_hull' -s 2 -c 2 --csv --log-file gpurun_out/line_counts.csv python tools/one_build.py c3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/line_counts.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value'); iid=h.index('ID')
for r in rows[1:]:
    print(r[iid], r[ik][:40], r[im], r[iv])
PY
