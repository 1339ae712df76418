"""tools/time_signed.py -- per-pass times of the integer-only build (no sqrt/combine epilogue) vs the full build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdf_tools_b200 import Context, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[wl]; nx, ny, nz = cfg["shape"]
m = torch.from_numpy(synth.make_mask(wl)).cuda()
out = torch.empty((nx, ny, nz), dtype=torch.int32, device="cuda")
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for name, fn in (("signed_d2 (FINAL=0)", lambda: ctx.signed_d2_device(m.data_ptr(), nx, ny, nz, out.data_ptr())),
                 ("build     (FINAL=1)", lambda: ctx.build_device(m.data_ptr(), nx, ny, nz, cfg["res"], out.data_ptr()))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    print(name, {k: round(v, 3) for k, v in ctx.last_timing().items()})
