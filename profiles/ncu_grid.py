#!/usr/bin/env python
"""A mid-size pair (default 50 000 correspondences, sheet scene) refined for 2 LM iterations: its linear solves are single
launches of pcg_grid_kernel (cooperative, one CTA per SM).  For ncu:  -k regex:pcg_grid -c 1 --set full"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
n = sys.argv[1] if len(sys.argv) > 1 else "50000"
sys.argv = [sys.argv[0], "--workload", "sheet", "--points", n]
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
bench.upload(ctx, prob)
w = pkg.make_weights(**sc["weights"])
ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
ctx.compute_rotations()
recs, st = ctx.optimize(w, 2)
print(f"{n} correspondences: {st.iterations} LM iterations, {st.total_trials} trials, {st.total_pcg_iters} PCG iterations, {st.kernel_launches} launches, {st.device_ms:.1f} ms")
