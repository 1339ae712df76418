"""tools/ab_env.py -- per-kernel timing of the product library on named workloads, optionally against variants selected by
environment variables (`name:ENV=value,...`; e.g. an alternative library through SDF_B200_LIB): each variant runs in its own
process on the same workloads; prints per-kernel medians and a checksum of the result.
  python tools/ab_env.py c3 c5 c4@2048,2048,128 exp:SDF_B200_LIB=/path/to/libvariant.so"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
from sdf_tools_b200 import Context, synth
res = {}
for wl in sys.argv[1:]:
    cfg = synth.CONFIGS[wl.split("@")[0]]; nx, ny, nz = cfg["shape"]
    if "@" in wl:
        dims = [int(v) for v in wl.split("@")[1].split(",")]
        nx, ny, nz = dims if len(dims) == 3 else dims * 3
    cache = "/tmp/mask_%%s.npy" %% wl
    if os.path.exists(cache): mh = np.load(cache)
    else: mh = synth.make_mask(wl.split("@")[0], (nx, ny, nz)); np.save(cache, mh)
    m = torch.from_numpy(mh).cuda()
    out = torch.empty((nx, ny, nz), dtype=torch.float32, device="cuda")
    ctx = Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    acc = {}
    for it in range(9):
        ctx.build_device(m.data_ptr(), nx, ny, nz, cfg["res"], out.data_ptr())
        torch.cuda.synchronize()
        if it >= 3:
            for k, v in ctx.last_timing().items(): acc.setdefault(k, []).append(v)
    fin = torch.isfinite(out)
    csum = float(out[fin].double().sum().item())
    res[wl] = {**{k: round(float(np.median(v)), 3) for k, v in acc.items()}, "checksum": csum}
    del m, out, ctx
print(json.dumps(res))
''' % ROOT
# name=ENV1=a,ENV2=b ... on the command line adds a variant; without any, the product library is timed as it is
variants = [("product", {})]
args = [a for a in sys.argv[1:] if "=" not in a]
for spec in [a for a in sys.argv[1:] if "=" in a]:
    name, _, rest = spec.partition(":")
    variants.append((name, dict(kv.split("=", 1) for kv in rest.split(",") if kv)))
wls = args or ["c3", "c5", "c2"]
for name, env in variants:
    r = subprocess.run([sys.executable, "-c", CHILD] + wls, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    print(name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-600:], flush=True)
