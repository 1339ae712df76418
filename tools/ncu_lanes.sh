#!/bin/bash
# lane efficiency / instruction counts of the envelope kernels on c3 at 512^3, 32-line vs 16-line tiles
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
cat > /tmp/one512.py <<'P'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from sdf_tools_b200 import Context, synth
mh = synth.make_mask("c3", (512, 512, 512))
m = torch.from_numpy(mh).cuda(); out = torch.empty((512,)*3, dtype=torch.float32, device="cuda")
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    ctx.build_device(m.data_ptr(), 512, 512, 512, 0.01, out.data_ptr()); torch.cuda.synchronize()
print(ctx.last_timing())
P
for v in 0 1; do
  SDF_B200_LPC16=$v timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:'run_strided' -s 2 -c 2 --csv --log-file gpurun_out/lanes_lpc16_$v.csv python /tmp/one512.py > /dev/null 2>&1
done
