#!/bin/bash
# Instruction / stall / DRAM counters of the three build kernels on the 1024^3 workload, per library:
# every library under sdf_tools_b200/lib/variants/ if that directory exists (A/B runs), else the product library.
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
mkdir -p gpurun_out
python tools/one_build.py c3 > /dev/null 2>&1   # caches the mask
if [ -d sdf_tools_b200/lib/variants ]; then LIBS=$(ls sdf_tools_b200/lib/variants/*.so); else LIBS=sdf_tools_b200/lib/libsdf_b200.so; fi
for lib in $LIBS; do
  n=$(basename $lib .so)
  SDF_B200_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:'binary_z|band_strided|line_hull' -c 3 --csv --log-file gpurun_out/counts_$n.csv python tools/one_build.py c3 > /dev/null 2>&1
done
