#!/usr/bin/env python
"""Host-side breakdown of one end-to-end step of bench.py (wall clock per C-ABI call)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
sys.argv = [sys.argv[0]] + sys.argv[1:]
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
w = pkg.make_weights(**sc["weights"])
ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
ctx.set_early_reject(args.early_rtol, args.early_margin)
for rep in range(2):
    t = [time.perf_counter()]
    ctx.triangulate(prob["pair"], prob["prm"], prob["uv1"], prob["uv2"]); t.append(time.perf_counter())
    ctx.problem_upload(prob["pair"], prob["X1"], prob["X2"], prob["uv1"], prob["uv2"], prob["d1"], prob["d2"], scale1=prob["s1"], scale2=prob["s2"]); t.append(time.perf_counter())
    ctx.set_graph(prob["rowptr"], prob["col"], prob["w"], prob["area"], prob["ntri"], 3); t.append(time.perf_counter())
    ctx.compute_rotations(); ctx.synchronize(); t.append(time.perf_counter())
    ctx.optimize(w, sc["lm_iters"]); t.append(time.perf_counter())
    ctx.download(doubles=False); t.append(time.perf_counter())
    names = ["triangulate", "problem_upload", "set_graph", "rotations", "optimize", "download"]
    print(" ".join(f"{n}={1e3*(b-a):.1f}ms" for n, a, b in zip(names, t[:-1], t[1:])), f"total={1e3*(t[-1]-t[0]):.1f}ms")
