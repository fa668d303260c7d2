#!/usr/bin/env python
"""Classic bundle adjustment (dsc_ba_*) on two views x N points: a short LM run for `ncu` (launch list / --set full).
usage: python profiles/ncu_ba.py [points] [iterations]"""
import importlib
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.package()
wl = importlib.import_module(pkg.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = wl.ba_scene(n, 2, seed=0)
with pkg.BundleAdjuster(0) as b:
    b.upload(sc["poses7"], sc["pose_fixed"], sc["cams"], sc["X"], sc["obs_pose"], sc["obs_point"], sc["obs_uv"], sc["obs_isg"])
    recs, st = b.optimize(iters, float(np.float32(np.sqrt(5.99))))
    print(f"{n} points: {st.iterations} LM iterations, {st.total_trials} trials, {st.kernel_launches} launches, {st.device_ms:.2f} ms, "
          f"chi2 {recs[0].chi2_before:.4e} -> {st.final_chi2:.4e}")
