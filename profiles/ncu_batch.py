#!/usr/bin/env python
"""One short launch of lm_batch_kernel (config 5: 10k-correspondence sheets) for `ncu --set full`.
usage: python profiles/ncu_batch.py [pairs] [lm_iters] [points]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
points = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
sys.argv = [sys.argv[0], "--points", str(points), "--workload", "sheet"]
args = bench.parse()
ctx = pkg.Context(0)
probs = []
for p in range(pairs):
    sc = bench.make_scene(pkg, args, p)
    probs.append(bench.prepare(pkg, ctx, sc, args))
w = pkg.make_weights(**sc["weights"])
with pkg.Batch(0) as b:
    b.upload(probs)
    b.set_pcg(rtol=1e-10, max_iters=6000)
    b.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
    recs, stats, ms = b.optimize(w, iters)
    print(dict(pairs=pairs, lm_iters=iters, ms=ms, pcg_iters=sum(s.total_pcg_iters for s in stats), trials=sum(s.total_trials for s in stats), **b.size()))
