"""ONE frame pair over two GPUs (SURVEY.md 8e second row; dsc.h dsc_shard_*): one process per GPU, halo rows and PCG
scalars exchanged by the kernels over NVLink peer memory.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise.

Checked: the LM trace and the refined pair of the sharded run against the single-GPU path of the same library (decisions
identical, values to the rounding of differently partitioned sums), against the C oracle, both ranks bit-identical to each
other, and bit-identical from run to run (fixed reduction order)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch, torch.distributed as dist
import importlib
import __graft_entry__ as g
pkg = g.package()
sh = importlib.import_module(pkg.__name__ + ".sharding")
from oracle import scenes, edges
rank, world, local = sh.rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, iters = {n}, {iters}
sc = scenes.tube_scene(n, seed=31, depth_sigma=0.0003)
p, keep = scenes.problem_from_scene(sc, "knn", 8, min_cos=0.99999)
w = pkg.make_weights(1.0, 1.0e7, 0.0003)
def upload(ctx):
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, p.inv_sigma2_1, p.inv_sigma2_2, scale1=p.s1, scale2=p.s2, Tg7=p.Tg.as7())
    gph = p.graph
    ctx.set_graph(gph.rowptr, gph.col, gph.w, gph.area, gph.n_triangles, 1)
    ctx.compute_rotations()
sp = sh.ShardedPair(pkg, local, p.n, dist, world, rank)
ctx = sp.ctx
upload(ctx)
ctx.set_pcg(rtol=1e-12, max_iters=40000, check_every=64)
c0, parts0 = ctx.cost(w)
runs = []
for rep in range(2):
    ctx.reset_state()
    recs, st = ctx.optimize(w, iters)
    out = ctx.download()
    sig = ctx.pixel_sigma()
    runs.append(dict(trace=[(r.chi2_before, r.chi2_after, r.lam, r.trials, r.accepted, r.pcg_iters) for r in recs], final=st.final_chi2,
                     X1=out["X1d"].copy(), X2=out["X2d"].copy(), scales=list(out["scales"]), Tg=list(out["Tg"]), update=out["update"],
                     sigma=list(sig), launches=st.kernel_launches, ms=st.device_ms))
info = ctx.shard_info()
res = dict(rank=rank, cost0=c0, info=info, n=p.n,
           traces=[r["trace"] for r in runs], finals=[r["final"] for r in runs], updates=[r["update"] for r in runs],
           same_run_to_run=bool(np.array_equal(runs[0]["X1"], runs[1]["X1"]) and np.array_equal(runs[0]["X2"], runs[1]["X2"]) and runs[0]["trace"] == runs[1]["trace"]),
           scales=runs[0]["scales"], Tg=runs[0]["Tg"], sigma=runs[0]["sigma"], ms=[r["ms"] for r in runs])
np.savez({out!r} + f".rank{{rank}}.npz", X1=runs[0]["X1"], X2=runs[0]["X2"])
if rank == 0:
    # the single-GPU path of the same library on the same pair
    with pkg.Context(local) as c1:
        upload(c1)
        c1.set_solver(1)
        c1.set_pcg(rtol=1e-12, max_iters=40000, check_every=64)
        os.environ["DSC_NO_CLUSTER_PCG"] = "1"
        s0, _ = c1.cost(w)
        recs, st = c1.optimize(w, iters)
        o1 = c1.download()
        res["single"] = dict(cost0=s0, trace=[(r.chi2_before, r.chi2_after, r.lam, r.trials, r.accepted, r.pcg_iters) for r in recs], final=st.final_chi2,
                             update=o1["update"], ms=st.device_ms)
        np.savez({out!r} + ".single.npz", X1=o1["X1d"], X2=o1["X2d"])
json.dump(res, open({out!r} + f".rank{{rank}}.json", "w"))
sp.close(dist)
dist.destroy_process_group()
"""


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(tmp_path, world, n, iters):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "shard")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, n=n, iters=iters, out=out))
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   DSC_SHARD_TIMEOUT_S="20")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-3000:] for o in outs]
    res = [json.load(open(out + f".rank{r}.json")) for r in range(world)]
    pts = [np.load(out + f".rank{r}.npz") for r in range(world)]
    return res, pts, np.load(out + ".single.npz")


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("n,iters", [(30011, 3), (3001, 4)])
def test_point_sharded_pair_on_two_gpus(tmp_path, n, iters):
    res, pts, single = _run(tmp_path, 2, n, iters)
    r0, r1 = res
    # the partition covers the pair, tile-aligned, and the two ranks really exchange halo rows
    assert r0["info"]["row_begin"] == 0 and r0["info"]["row_end"] == r1["info"]["row_begin"] and r1["info"]["row_end"] == r0["n"]
    assert r0["info"]["row_end"] % 512 == 0 and 0 < r0["info"]["row_end"] < r0["n"]
    assert r0["info"]["halo_rows"] > 0 and r1["info"]["halo_rows"] > 0
    # both ranks hold the same bits: trace, refined pair, globals
    assert r0["traces"] == r1["traces"] and r0["finals"] == r1["finals"] and r0["scales"] == r1["scales"] and r0["Tg"] == r1["Tg"]
    assert np.array_equal(pts[0]["X1"], pts[1]["X1"]) and np.array_equal(pts[0]["X2"], pts[1]["X2"])
    assert r0["updates"] == r1["updates"] and r0["sigma"] == r1["sigma"]
    # ... and the same bits from run to run (sums over blocks and ranks in a fixed order)
    assert r0["same_run_to_run"] and r1["same_run_to_run"]
    # against the single-GPU path: identical LM decisions, values to the rounding of differently partitioned sums
    s = r0["single"]
    assert r0["cost0"] == pytest.approx(s["cost0"], rel=1e-12)
    tr = r0["traces"][0]
    assert [t[3] for t in tr] == [t[3] for t in s["trace"]] and [t[4] for t in tr] == [t[4] for t in s["trace"]]
    for a, b in zip(tr, s["trace"]):
        assert a[0] == pytest.approx(b[0], rel=1e-7) and a[1] == pytest.approx(b[1], rel=1e-7) and a[2] == pytest.approx(b[2], rel=1e-6)
    assert r0["finals"][0] == pytest.approx(s["final"], rel=1e-7)
    scale = np.abs(single["X1"]).max()
    assert np.abs(pts[0]["X1"] - single["X1"]).max() <= 1e-6 * scale and np.abs(pts[0]["X2"] - single["X2"]).max() <= 1e-6 * scale
    assert r0["updates"][0] == pytest.approx(s["update"], rel=1e-5)
    print(f"[shard n={n}] 2 GPUs {r0['ms']} ms vs 1 GPU {s['ms']:.1f} ms; halo rows {r0['info']['halo_rows']} / {r1['info']['halo_rows']}")


def test_sharded_context_refuses_what_it_does_not_do(pkg):
    """no peers needed: a world of one rank; fp32 mode, the dense solver and the test hooks are refused, a missing attach is named"""
    from oracle import scenes, edges
    sc = scenes.sheet_scene(700, seed=3)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = pkg.make_weights(1.0, 2.0e5, 0.003)
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    with pkg.Context(0) as ctx:
        h = ctx.shard_init(0, 1, p.n)
        assert h.any()
        with pytest.raises(pkg.DscError):
            ctx.shard_init(0, 1, p.n)                                           # once
        ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, scale1=p.s1, scale2=p.s2)
        with pytest.raises(pkg.DscError) as e:
            ctx.set_graph(p.graph.rowptr, p.graph.col, p.graph.w, p.graph.area, p.graph.n_triangles, 1)
        assert e.value.status == -9                                             # DSC_ERR_SHARD: not attached
        ctx.shard_attach(h[None, :])
        ctx.set_graph(p.graph.rowptr, p.graph.col, p.graph.w, p.graph.area, p.graph.n_triangles, 1)
        ctx.compute_rotations()
        ctx.set_precision("f32")
        with pytest.raises(pkg.DscError):
            ctx.optimize(w, 1)
        ctx.set_precision("f64")
        ctx.set_solver(2)
        with pytest.raises(pkg.DscError):
            ctx.optimize(w, 1)
        ctx.set_solver(0)
        with pytest.raises(pkg.DscError):
            ctx.debug_linearize(w)
        # a world of one rank runs the sharded kernels against itself: same result as the plain path
        ctx.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
        recs, st = ctx.optimize(w, 3)
        o = ctx.download()
        info = ctx.shard_info()
        assert info["row_begin"] == 0 and info["row_end"] == p.n and info["halo_rows"] == 0
    with pkg.Context(0) as c1:
        c1.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, scale1=p.s1, scale2=p.s2)
        c1.set_graph(p.graph.rowptr, p.graph.col, p.graph.w, p.graph.area, p.graph.n_triangles, 1)
        c1.compute_rotations()
        c1.set_solver(1)
        c1.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
        r1, s1 = c1.optimize(w, 3)
        o1 = c1.download()
    assert [r.trials for r in recs] == [r.trials for r in r1]
    assert st.final_chi2 == pytest.approx(s1.final_chi2, rel=1e-8)
    assert np.abs(o["X1d"] - o1["X1d"]).max() <= 1e-7 * np.abs(o1["X1d"]).max()
