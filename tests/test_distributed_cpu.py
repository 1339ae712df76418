"""Host-side logic of the N > 1 path on CPU: slab partitioning, the per-line summary / carry protocol
(numpy restatement in sdf_tools_b200/distributed.py) and its exchange over torch.distributed (gloo, world_size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from sdf_tools_b200 import distributed as D
from sdf_tools_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_ranges():
    assert D.slab_ranges(1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]
    for nz, nr in [(100, 3), (17, 4), (2048, 8), (33, 2), (16, 1)]:
        r = D.slab_ranges(nz, nr)
        assert r[0][0] == 0 and r[-1][1] == nz and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(b > a for a, b in r)


def test_binary_pass_matches_oracle_on_lines():
    rng = np.random.default_rng(3)
    for _ in range(30):
        n = int(rng.integers(1, 70))
        line = (rng.random((1, 1, n)) < rng.choice([0.0, 0.1, 0.5, 1.0])).astype(np.uint8)
        assert np.array_equal(D.binary_z_np(line), O.signed_d2(line))


@pytest.mark.parametrize("nr", [1, 2, 3, 5, 8])
def test_slab_protocol_numpy(nr):
    rng = np.random.default_rng(nr)
    for trial in range(8):
        shape = (int(rng.integers(1, 20)), int(rng.integers(1, 20)), int(rng.integers(nr, 70)))
        m = (rng.random(shape) < rng.choice([0.0, 0.01, 0.3, 0.9, 1.0])).astype(np.uint8)
        if trial % 3 == 0 and shape[2] >= 6:
            a, b = shape[2] // 3, 2 * shape[2] // 3
            m[:, :, a:b] = m[:, :, a:a + 1]                    # slabs that are a single run
        rs = D.slab_ranges(shape[2], nr)
        widths = [b - a for a, b in rs]
        sums = np.stack([D.summary_np(m[:, :, a:b]) for a, b in rs])
        got = np.concatenate([D.binary_z_np(m[:, :, a:b], *D.resolve_carries_np(sums, widths, r))
                              for r, (a, b) in enumerate(rs)], axis=2)
        assert np.array_equal(got, D.binary_z_np(m))


_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["SDF_ROOT"])
from sdf_tools_b200 import distributed as D, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shape = (12, 10, 48)
rs = D.slab_ranges(shape[2], world)
widths = [b - a for a, b in rs]
mask = synth.make_mask("c3", shape, zrange=rs[rank])                      # each rank generates only its slab
summary = torch.from_numpy(D.summary_np(mask).astype(np.int64))
gathered = [torch.empty_like(summary) for _ in range(world)]
dist.all_gather(gathered, summary)                                         # the one data-path collective
allsum = np.stack([g.numpy().astype(np.uint32) for g in gathered])
lo, hi = D.resolve_carries_np(allsum, widths, rank)
mine = D.binary_z_np(mask, lo, hi)
full = synth.make_mask("c3", shape)
want = D.binary_z_np(full)[:, :, rs[rank][0]:rs[rank][1]]
ok = torch.tensor([int(np.array_equal(mine, want))])
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
# extrema: all-reduce(max) of the per-rank raw maxima
raw = torch.tensor([int(mine.max(initial=0)), int((-mine).max(initial=0))])
dist.all_reduce(raw, op=dist.ReduceOp.MAX)
fullb = D.binary_z_np(full)
ok2 = int(raw[0]) == int(fullb.max(initial=0)) and int(raw[1]) == int((-fullb).max(initial=0))
if rank == 0:
    print("RESULT", int(ok.item()), int(ok2))
dist.destroy_process_group()
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, SDF_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "RESULT 1 1" in res.stdout, res.stdout[-500:] + res.stderr[-1500:]
