"""tools/ab_variants.py -- A/B timing of kernel variants: each library under sdf_tools_b200/lib/variants/ (built by hand
from a modified copy of csrc/) is loaded in its own process through SDF_B200_LIB and timed on the same workload."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from sdf_tools_b200 import Context, synth
wl = sys.argv[1]
cfg = synth.CONFIGS[wl]; nx, ny, nz = cfg["shape"]
cache = "/tmp/mask_%%s.npy" %% wl
if os.path.exists(cache): mh = np.load(cache)
else: mh = synth.make_mask(wl); np.save(cache, mh)
m = torch.from_numpy(mh).cuda()
out = torch.empty((nx, ny, nz), dtype=torch.float32, device="cuda")
ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
acc = {}
for it in range(8):
    ctx.build_device(m.data_ptr(), nx, ny, nz, cfg["res"], out.data_ptr())
    torch.cuda.synchronize()
    if it >= 3:
        for k, v in ctx.last_timing().items(): acc.setdefault(k, []).append(v)
csum = float(out[torch.isfinite(out)].double().sum().item())
print(json.dumps({"lib": os.path.basename(os.environ.get("SDF_B200_LIB", "default")), **{k: round(float(np.median(v)), 3) for k, v in acc.items()}, "checksum": csum}))
''' % ROOT
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
for lib in sorted(glob.glob(os.path.join(ROOT, "sdf_tools_b200", "lib", "variants", "*.so"))):
    env = dict(os.environ, SDF_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD, wl], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + lib + " " + r.stderr[-400:], flush=True)
