"""CPU suite for the multi-GPU path: frames are sharded over ranks with no data-path collective.
The world_size-2 test runs two real processes over the gloo backend (rendezvous on 127.0.0.1)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
import __graft_entry__ as ge
sh = ge.load_package().sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = sh.strided_frames(1800, rank, world)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
# the bench's per-step frame choice: the same 20-frame list on every rank, rotated by the rank
avail = list(range(14))
steps = [sh.bench_frame(s, rank, 20, avail) for s in range(20)]
all_steps = [None] * world
dist.all_gather_object(all_steps, steps)
# timing protocol of bench.py: max over ranks of the per-rank wall time, sum of work
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
w = torch.tensor([float(len(mine))], dtype=torch.float64)
dist.all_reduce(w, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"ok": sh.check_partition(gathered, 1800), "tmax": t.item(), "frames": w.item(),
                      "steps": [sorted(s) for s in all_steps], "order": all_steps}))
dist.destroy_process_group()
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_strided_partition(pkg):
    sh = pkg.sharding
    for world in (1, 2, 4, 8):
        parts = [sh.strided_frames(1800, r, world) for r in range(world)]
        assert sh.check_partition(parts, 1800)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert sh.strided_frames(10, 1, 4) == [1, 5, 9]
    assert sh.strided_frames(1800, 0, 8, begin=370, step=1)[:2] == [370, 378]
    with pytest.raises(ValueError):
        sh.strided_frames(10, 4, 4)
    assert sh.animation_seconds(0.1, 1800, 8) == pytest.approx(22.5)


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] is True
    assert res["tmax"] == 2.0 and res["frames"] == 1800.0
    want = sorted([j % 14 for j in range(20)])
    assert res["steps"] == [want, want]            # both ranks render the same multiset of frames ...
    assert res["order"][0] != res["order"][1]      # ... in a different order (rotated by the rank)


def test_bench_frame_multisets(pkg):
    sh = pkg.sharding
    avail = [0, 100, 200, 330, 420, 520, 660, 800, 1000, 1100, 1250, 1400, 1600, 1750]
    for steps in (1, 3, 14, 20, 33):
        ref = sorted(sh.bench_frame(i, 0, steps, avail) for i in range(steps))
        for world in (2, 4, 8):
            for rank in range(world):
                assert sorted(sh.bench_frame(i, rank, steps, avail) for i in range(steps)) == ref
    with pytest.raises(ValueError):
        sh.bench_frame(0, 0, 0, avail)
