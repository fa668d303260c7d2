#!/usr/bin/env python
"""Launch every kernel of the refinement once on the bench workload (for `ncu --set full`, which replays each launch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["f64"]      # e.g. "f64,f32": the kernels of both precisions
sys.argv = [sys.argv[0], "--points", sys.argv[1] if len(sys.argv) > 1 else "1000000", "--k", sys.argv[2] if len(sys.argv) > 2 else "8"]
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
bench.upload(ctx, prob)
w = pkg.make_weights(**sc["weights"])
for mode in modes:
    ctx.set_precision(mode)
    k = ctx.profile_kernels(w, warm=1, reps=1)
    print(mode, {name: round(v["ms"] * 1e3, 1) for name, v in k.items()})
