#!/usr/bin/env python
"""Per-LM-iteration record of one bench step (trials, PCG iterations, lambda, chi2)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
bench.upload(ctx, prob)
w = pkg.make_weights(**sc["weights"])
ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
if len(args.early_rtol) > 0:
    ctx.set_early_reject(args.early_rtol, args.early_margin)
recs, st = ctx.optimize(w, args.lm_iters or sc["lm_iters"])
for i, r in enumerate(recs):
    print(f"{i:2d} chi2 {r.chi2_before:.6e} -> {r.chi2_after:.6e} lambda {r.lam:.3e} trials {r.trials} pcg {r.pcg_iters}")
print(f"device {st.device_ms:.1f} ms: linearize {st.linearize_ms:.1f} pcg {st.pcg_ms:.1f} trial {st.trial_ms:.1f}; launches {st.kernel_launches}; early rejects {st.early_rejects}")
