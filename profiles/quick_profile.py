#!/usr/bin/env python
"""Per-kernel CUDA-event times on the bench workload (no LM run): python profiles/quick_profile.py [points]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g

pkg = g.package()
args = bench.parse.__wrapped__() if hasattr(bench.parse, "__wrapped__") else None
sys.argv = [sys.argv[0], "--points", sys.argv[1] if len(sys.argv) > 1 else "1000000", "--k", sys.argv[2] if len(sys.argv) > 2 else "8"]
args = bench.parse()
ctx = pkg.Context(0)
sc = bench.make_scene(pkg, args, 0)
prob = bench.prepare(pkg, ctx, sc, args)
bench.upload(ctx, prob)
w = pkg.make_weights(**sc["weights"])
k = ctx.profile_kernels(w, warm=3, reps=20)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
for name, v in k.items():
    gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9
    print(f"{name:14s} {v['ms']*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/peak:6.3f} of peak")
