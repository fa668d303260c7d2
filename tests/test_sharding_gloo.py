"""N > 1 host logic on CPU: two gloo ranks shard a batch of frame-pair problems and fold their timings."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import __graft_entry__ as g
import importlib
pkg = g.package()
sh = importlib.import_module(pkg.__name__ + ".sharding")
rank, world, local = sh.rank_world()
dist.init_process_group("gloo", init_method="env://")
mine = sh.shard(37, world, rank)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
times, counts = sh.fold(dist, "cpu", [10.0 * (rank + 1), 5.0], [len(mine), 3])
# the one collective of the batched path: per-problem result rows -> the table in global problem order on every rank
import numpy as np
rows = np.array([[100.0 + p, 2.0 * p] for p in mine])
table = sh.gather_by_problem(dist, "cpu", 37, world, rank, rows)
ok = bool(np.array_equal(table[:, 0], 100.0 + np.arange(37)) and np.array_equal(table[:, 1], 2.0 * np.arange(37)))
# the one use of torch.distributed by the point-sharded pair: the all-gather of the ranks' 64-byte arena handles
handle = (np.arange(64) * (rank + 3) % 251).astype(np.uint8)
table = sh.exchange_handles(dist, "cpu", handle, world, rank)
ok = ok and table.shape == (world, 64) and all(np.array_equal(table[r], (np.arange(64) * (r + 3) % 251).astype(np.uint8)) for r in range(world))
oks = [None] * world
dist.all_gather_object(oks, ok)
if rank == 0:
    print(json.dumps(dict(shards=gathered, times=times, counts=counts, rate=sh.throughput(counts[0], times[0]), gather_ok=oks)))
dist.destroy_process_group()
"""


def test_two_rank_sharding_and_fold(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-2000:] for o in outs]
    import json
    js = json.loads(outs[0][0].strip().splitlines()[-1])
    a, b = js["shards"]
    assert sorted(a + b) == list(range(37)) and not set(a) & set(b)        # every problem exactly once
    assert js["times"] == [20.0, 5.0]                                     # max over ranks
    assert js["counts"] == [37.0, 6.0]                                    # sum over ranks
    assert js["rate"] == 37 / 0.02
    assert js["gather_ok"] == [True, True]                                # both ranks hold the table in problem order


def test_shard_edge_cases(pkg):
    import importlib
    import pytest
    sh = importlib.import_module(pkg.__name__ + ".sharding")
    assert sh.shard(0, 4, 1) == []
    assert sh.shard(3, 8, 5) == []
    assert sh.shard(3, 8, 2) == [2]
    assert sum(len(sh.shard(4096, 8, r)) for r in range(8)) == 4096
    with pytest.raises(ValueError):
        sh.shard(4, 2, 2)
    assert [sh.owner(p, 4) for p in (0, 3, 4, 9)] == [(0, 0), (3, 0), (0, 1), (1, 2)]
    import numpy as np
    t = sh.gather_by_problem(None, "cpu", 5, 1, 0, np.arange(10.0).reshape(5, 2))        # single process: identity
    assert np.array_equal(t, np.arange(10.0).reshape(5, 2))
    with pytest.raises(ValueError):
        sh.gather_by_problem(None, "cpu", 5, 2, 0, np.zeros((2, 1)))                       # rank 0 of 2 owns 3 of 5


def test_point_shard_partition_is_tile_aligned_balanced_and_complete(pkg):
    """dsc_shard_partition (pure host function of the C ABI): the row partition of a point-sharded pair"""
    import numpy as np
    import pytest
    rng = np.random.default_rng(0)
    for nslices, world in ((1, 1), (5, 2), (16, 4), (1000, 8), (31250, 8), (31250, 3)):
        width = rng.integers(6, 14, nslices)
        sp = np.concatenate([[0], np.cumsum(width)]).astype(np.int32)
        rb = pkg.shard_partition(sp, world)
        assert rb[0] == 0 and rb[-1] == nslices * 32 and np.all(np.diff(rb) >= 0)
        assert np.all((rb % 512 == 0) | (rb == nslices * 32))                  # tile boundaries (or the end: a rank may be empty)
        if nslices >= 64 * world:                                              # enough tiles: the ranks' work differs by less than 2 tiles
            cost = lambda a, b: (sp[min(nslices, b // 32)] - sp[a // 32]) + 3.37 * (min(nslices, b // 32) - a // 32)
            work = [cost(rb[r], rb[r + 1]) for r in range(world)]
            tile = 16 * (width.mean() + 3.37)
            assert max(work) - min(work) <= 2.5 * tile
    with pytest.raises(pkg.DscError):
        pkg.shard_partition(np.zeros(3, np.int32), 9)                          # more ranks than DSC_SHARD_MAX_RANKS
