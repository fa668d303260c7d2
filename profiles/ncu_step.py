#!/usr/bin/env python
"""ONE timed step of bench.py (config 3, 1 M correspondences, 30 LM iterations, early rejection as the bench uses it) between
cudaProfilerStart / cudaProfilerStop, for an ncu launch list of exactly that step:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python profiles/ncu_step.py
  python profiles/launch_shares.py launches.csv"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
sys.argv = [sys.argv[0]] + sys.argv[1:]
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
bench.upload(ctx, prob)
w = pkg.make_weights(**sc["weights"])
lm_iters = args.lm_iters or sc["lm_iters"]
ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
ctx.set_early_reject(args.early_rtol, args.early_margin)
ctx.compute_rotations()
ctx.optimize(w, lm_iters)                 # warm-up (not profiled)
ctx.reset_state()
try:
    rt = ctypes.CDLL("libcudart.so.12")
except OSError:
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
rt.cudaProfilerStart()
recs, st = ctx.optimize(w, lm_iters)
rt.cudaProfilerStop()
print(f"profiled step: {st.iterations} LM iterations, {st.total_trials} trials, {st.total_pcg_iters} PCG iterations, {st.kernel_launches} launches, "
      f"{st.device_ms:.1f} ms under the profiler")
