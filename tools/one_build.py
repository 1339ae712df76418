"""tools/one_build.py -- one device-resident build of a named workload (for ncu captures of single launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdf_tools_b200 import Context, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[wl]; nx, ny, nz = cfg["shape"]
cache = "/tmp/mask_%s.npy" % wl
if os.path.exists(cache): mh = np.load(cache)
else: mh = synth.make_mask(wl); np.save(cache, mh)
m = torch.from_numpy(mh).cuda()
out = torch.empty((nx, ny, nz), dtype=torch.float32, device="cuda")
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for _ in range(2):                                   # the second build is the warm one (ncu: -s 3 -c 3 skips the first)
    ctx.build_device(m.data_ptr(), nx, ny, nz, cfg["res"], out.data_ptr())
    torch.cuda.synchronize()
print("ok", ctx.last_timing())
