"""Multi-GPU paths on real hardware (skipped below 2 visible GPUs):
  * the single-process multi-device context of the C ABI (sdf_b200_create_multi): z slabs, summaries exchanged over
    peer memory by the summary kernel itself, or through NCCL -- must reproduce the single-device build bit for bit;
  * the one-process-per-GPU slab build (bench_multi.py under torch.distributed.run, real NCCL / symmetric memory),
    checked against the CPU oracle at BASELINE.json's C5 grid and on a 512x512x1024 copy of the C4 grid.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    try:
        from sdf_tools_b200 import _capi
        return _capi.load().sdf_b200_device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_multi_device_context_matches_single_device(exchange):
    code = r'''
import os, sys, numpy as np
sys.path.insert(0, %r)
from sdf_tools_b200 import Context, synth, _capi
from sdf_tools_b200.api import COLLISION_CELL_DTYPE
n = min(_capi.load().sdf_b200_device_count(), 8)
one = Context(0)
multi = Context(list(range(n)))
print("exchange", multi.exchange_kind, "devices", n)
cases = [synth.make_mask("c3", (96, 80, 128)), synth.iid((40, 33, 64), 0.3, 5), synth.iid((17, 9, 50), 0.02, 6),
         np.zeros((8, 8, 32), np.uint8), synth.iid((5, 4, 3), 0.4, 2), synth.make_mask("c5", (128, 128, 128))]
uni = synth.iid((24, 20, 96), 0.05, 9); uni[:, :, 32:64] = uni[:, :, 32:33]; cases.append(uni)
for m in cases:
    a, ea = one.build_from_mask(m, 0.05)
    b, eb = multi.build_from_mask(m, 0.05)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), m.shape
    assert ea == eb, (ea, eb)
occ = synth.octree((64, 64, 64), 0.3, 5)
cells = np.zeros(occ.shape, dtype=COLLISION_CELL_DTYPE); cells["occupancy"] = occ
for unk in (False, True):
    a, ea = one.build_from_cells(cells, 0.02, unk)
    b, eb = multi.build_from_cells(cells, 0.02, unk)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and ea == eb
print("OK")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SDF_B200_EXCHANGE=exchange), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    if exchange == "nccl":
        assert "exchange nccl" in r.stdout, r.stdout


def _torchrun_bench(extra, nproc, exchange="p2p"):
    port = 29500 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(nproc), "--steps", "2", "--warmup", "1",
           "--no-e2e", "--no-cpu-baseline", "--no-c4", "--verify"] + extra
    r = subprocess.run(cmd, env=dict(os.environ, SDF_B200_EXCHANGE=exchange), capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@needs2
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_torchrun_slab_build_c5_512_vs_oracle(exchange):
    """BASELINE.json configs[4]: 512^3 octree-style grid on all visible GPUs, every rank's slab against the CPU oracle."""
    n = min(_ndev(), 8)
    out = _torchrun_bench(["--workload", "c5"], n, exchange)
    assert out["n_gpus"] == n and out["verified_vs_oracle"] is True and out["verified_vs_single_gpu"] is True
    if exchange == "nccl":
        assert out["config"]["exchange"] == "nccl"


@needs2
def test_torchrun_slab_build_c4_copy_vs_oracle():
    """A 512x512x1024 copy of the C4 grid (BASELINE.json configs[3]: z-sharded, one exchange per build)."""
    n = min(_ndev(), 8)
    out = _torchrun_bench(["--workload", "c4", "--shape", "512,512,1024"], n)
    assert out["verified_vs_oracle"] is True and out["verified_vs_single_gpu"] is True
