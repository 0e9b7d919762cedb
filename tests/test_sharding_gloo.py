"""N>1 host logic on CPU: two gloo ranks shard a 7-frame batch round-robin, process their frames
(with the oracle standing in for the device here — this test is about the sharding, ordering and
the max-over-ranks timing reduction, not about pixels), and the merged result must equal the
single-process result."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from gst_plugins_rs_b200 import sharding, frames
import oracle

assert sharding.init_process_group("gloo")
rank, _, size = sharding.world()
n_frames, w, h = 7, 96, 8
mine = sharding.frames_for_rank(n_frames, rank, size)
out = {}
for i in mine:
    src = frames.frame_rand(w, h, 4, i)
    out[i] = int(oracle.hsvfilter(src, w, h, "RGBA", (37.5, 1.2, 0.05, 0.9, 0.02)).astype(np.uint64).sum())
sharding.barrier()
slowest = sharding.max_over_ranks(10.0 + rank)        # rank-dependent "time"
total = sharding.sum_over_ranks(len(mine))
gathered = [None] * size
dist.all_gather_object(gathered, [out[i] for i in mine])
if rank == 0:
    merged = sharding.merge_in_order(gathered, n_frames)
    print(json.dumps({"merged": merged, "slowest": slowest, "total": total}))
dist.destroy_process_group()
"""


def test_two_rank_round_robin(orc, tmp_path):
    from gst_plugins_rs_b200 import frames, sharding
    assert sharding.frames_for_rank(7, 0, 2) == [0, 2, 4, 6]
    assert sharding.frames_for_rank(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((sharding.frames_for_rank(64, r, 8) for r in range(8)), [])) == list(range(64))
    assert sharding.merge_in_order([["a", "c"], ["b"]], 3) == ["a", "b", "c"]
    assert sharding.row_bands(2160, 8) == [(270 * i, 270) for i in range(8)]
    bands = sharding.row_bands(1081, 4)
    assert bands == [(0, 271), (271, 270), (541, 270), (811, 270)] and sum(n for _, n in bands) == 1081
    assert sharding.row_bands(3, 8)[3:] == [(3, 0)] * 5

    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    want = [int(orc.hsvfilter(frames.frame_rand(96, 8, 4, i), 96, 8, "RGBA",
                              (37.5, 1.2, 0.05, 0.9, 0.02)).astype(np.uint64).sum())
            for i in range(7)]
    assert res["merged"] == want          # every frame once, back in stream order
    assert res["slowest"] == 11.0         # max over ranks, not rank 0's own time
    assert res["total"] == 7.0
