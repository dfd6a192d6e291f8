"""CPU suite, part 3: the N>1 host logic with world_size-2 gloo (no GPU needed)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import medseg_b200
from medseg_b200.sharding import shard_range, max_over_ranks, sum_over_ranks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(257, world, rank)
owned = torch.zeros(257, dtype=torch.int32); owned[lo:hi] = 1
dist.all_reduce(owned)
t = max_over_ranks(10.0 + rank)
n = sum_over_ranks(hi - lo)
dist.barrier()
if rank == 0:
    print(json.dumps({"cover": bool((owned == 1).all()), "tmax": t, "n": n, "world": world}))
dist.destroy_process_group()
'''


def test_shard_range_partitions():
    from medseg_b200.sharding import shard_range
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(256, 8, r) for r in (0, 7)] == [(0, 32), (224, 256)]


def _torchrun(args, script_text=None, script_path=None, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_gloo_world2_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d == {"cover": True, "tmax": 11.0, "n": 257.0, "world": 2}


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` under torchrun: rank 0 alone runs and prints, the other rank exits 0."""
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--cpu-sample", "1"], timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
