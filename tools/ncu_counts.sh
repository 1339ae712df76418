#!/bin/bash
# instruction / stall counters of the two envelope passes for every library under sdf_tools_b200/lib/variants
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
python tools/one_build.py c3 > /dev/null 2>&1   # caches the mask
for lib in sdf_tools_b200/lib/variants/*.so; do
  n=$(basename $lib .so)
  SDF_B200_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:band_strided -c 2 --csv --log-file gpurun_out/counts_$n.csv python tools/one_build.py c3 > /dev/null 2>&1
done
