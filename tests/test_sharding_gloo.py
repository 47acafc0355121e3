"""world_size-2 gloo test (CPU) of the N>1 host logic: shard partitioning and the optional all-gather of the output."""
import os
import subprocess
import sys

import numpy as np

from dali_b200.sharding import shard_range, shard_of, sharded_source

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_follow_reader_convention():
    # loader.cc:78-81: start = size * shard_id / num_shards
    assert [shard_range(10, s, 3) for s in range(3)] == [(0, 3), (3, 6), (6, 10)]
    assert [shard_range(256, s, 8) for s in range(8)] == [(32 * s, 32 * s + 32) for s in range(8)]
    items = list(range(10))
    assert sum((shard_of(items, s, 4) for s in range(4)), []) == items
    src = sharded_source(items, 4, 1, 2)
    assert src(0) == [5, 6, 7, 8] and src(1) == [9, 5, 6, 7]


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from dali_b200.sharding import all_gather_output, shard_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(8, r, w)
local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous().to(torch.float16)
full = all_gather_output(local)
assert full.shape == (8, 3, 2, 2) and full.dtype == torch.float16
assert torch.equal(full[:, 0, 0, 0].float(), torch.arange(8, dtype=torch.float32)), full[:, 0, 0, 0]
from dali_b200.sharding import GatherBuffer
gb = GatherBuffer((hi - lo, 3, 2, 2), torch.float16, torch.device("cpu"))
gb.local.copy_(local)                                   # on the GPU the last operator writes here directly
gb.all_gather()
assert torch.equal(gb.full, full)
dist.barrier()
if r == 0:
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
