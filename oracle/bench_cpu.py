"""CPU baseline leg of bench.py (oracle = test infrastructure; this is the one place bench.py may
execute it).  Times the oracle's restatement of arapOptimization on a bounded sample of the bench workload and
extrapolates linearly in the number of correspondences to the bench size (optimistic for the CPU: both the sparse
factorisation and the PCG iteration count grow faster than linearly).  Two implementations:
  impl="c"     oracle/c/dsc_oracle.c: compiled C, OpenMP over all host cores, analytic Jacobians, matrix-free
               block-Jacobi PCG to the same tolerance as the CUDA path (default);
  impl="numpy" the numpy port: assembled sparse normal equations + SuperLU direct solve with the 8 global unknowns
               eliminated (the stand-in for g2o's LinearSolverEigen), one thread.
kind = "port" for both: the reference itself cannot be compiled in this image.
"""
import os
import time

import numpy as np

from . import camera, edges, lm
from .f32 import Pose
from .graph import knn_graph, compute_rotations
from .triangulate import triangulate_pairs, init_depth_scale_sim, GATE_SIM
from .se3 import SE3


def build_from_arrays(sc, n_sample, k):
    cam = (camera.KB8, np.asarray(sc["cam"], np.float32))
    T1, T2 = Pose.from34(sc["T1"]), Pose.from34(sc["T2"])
    uv1, uv2, d1, d2 = sc["uv1"], sc["uv2"], sc["d1"], sc["d2"]
    X1, X2, valid, _ = triangulate_pairs(uv1, uv2, cam, cam, T1, T2, "NRSLAM", "FarPoints", GATE_SIM, sc["min_cos"])
    idx = np.nonzero(valid)[0][:n_sample]
    X1, X2 = X1[idx], X2[idx]
    n = len(idx)
    s1 = init_depth_scale_sim(d1[idx], X1, T1, np.ones(n, bool))
    s2 = init_depth_scale_sim(d2[idx], X2, T2, np.ones(n, bool))
    g = knn_graph(X1.astype(np.float64), k, sc["area"])
    p = edges.Problem(cam1=cam, cam2=cam, T1=T1, T2=T2, uv1=uv1[idx], uv2=uv2[idx], inv_sigma2_1=np.ones(n),
                      inv_sigma2_2=np.ones(n), d1=d1[idx].astype(np.float64), d2=d2[idx].astype(np.float64), graph=g,
                      X1=X1.astype(np.float64), X2=X2.astype(np.float64), Tg=SE3(), s1=s1, s2=s2)
    p.R = compute_rotations(g, p.X1, p.X2)
    return p


def run(workload, n_sample, k, steps, warmup, n_full, lm_iters=2, sc=None, impl="c", pcg_rtol=1e-10):
    import importlib.util
    import sys
    if sc is None:
        import __graft_entry__ as g
        wl = importlib.import_module(g.package().__name__ + ".workloads")      # input synthesis only
        n_gen = int(n_sample * 1.2) + 64
        if workload == "drunkard":
            sc = wl.tube_scene(n_gen, seed=0, cam=wl.DRUNKARD_CAM, arap=1.0e7, depth_sigma=0.0003)
            name = "config3: Drunkard.yaml-shaped tube, 1 frame pair"
        elif workload == "realcolon":
            sc = wl.tube_scene(n_gen, seed=0, cam=wl.REALCOLON_CAM, arap=0.1, depth_sigma=1e-6, scales=(1.0, 1.0))
            name = "config4: Realcolon.yaml-shaped tube + border mask, 1 frame pair"
        else:
            sc = wl.sheet_scene(n_gen, seed=0)
            name = "config2: Simulation.yaml sheet, 1 frame pair"
    else:
        name = sc.get("name", workload)
    p = build_from_arrays(sc, n_sample, k)
    w = edges.Weights(**sc["weights"])
    cores = os.cpu_count() or 1
    if impl == "c":
        from . import cport
        cport.build()
        used = cport.threads()

        def one(iters):
            cp = cport.CProblem(p)
            t = time.perf_counter()
            cport.compute_rotations(cp)                      # computeR is part of arapOptimization
            tr = cport.optimize(cp, w, iters, threads=0, pcg_rtol=pcg_rtol)
            return len(tr["chi2"]), time.perf_counter() - t, sum(tr["pcg_iters"])
        how = (f"C oracle (oracle/c/dsc_oracle.c: analytic Jacobians, matrix-free block-Jacobi PCG to rtol {pcg_rtol:g}, "
               f"OpenMP), {used} threads of {cores} host cores (the reference is single-threaded)")
    else:
        used = 1

        def one(iters):
            t = time.perf_counter()
            _, tr = lm.optimize(p, w, iters)
            return len(tr.chi2), time.perf_counter() - t, 0
        how = f"numpy/scipy port with SuperLU direct solve, 1 thread of {cores} host cores (the reference is single-threaded)"
    for _ in range(warmup):
        one(1)
    its, dt, cg = 0, 0.0, 0
    for _ in range(max(1, steps)):
        a, b, c = one(lm_iters)
        its += a; dt += b; cg += c
    rate_sample = its / dt
    value = rate_sample * p.n / float(n_full)
    # triangulation (K1) on the CPU: the vectorised numpy oracle on the sample's matches, 2 map points per match
    cam = (camera.KB8, np.asarray(sc["cam"], np.float32))
    T1, T2 = Pose.from34(sc["T1"]), Pose.from34(sc["T2"])
    m = min(len(sc["uv1"]), 200000)
    t = time.perf_counter()
    triangulate_pairs(sc["uv1"][:m], sc["uv2"][:m], cam, cam, T1, T2, "NRSLAM", "FarPoints", GATE_SIM, sc["min_cos"])
    tri_rate = 2.0 * m / (time.perf_counter() - t)
    cpu = dict(value=value, unit="LM it/s", cores=used, kind="port", triangulated_points_per_s=tri_rate,
               sample=f"{its} LM iterations ({cg} PCG iterations) on {p.n} correspondences (k={k}) in {dt:.1f} s = "
                      f"{rate_sample:.3f} it/s, scaled linearly to {n_full} correspondences; {how}; triangulation: numpy "
                      f"oracle, {m} matches, 1 thread")
    return dict(value=value, ms_per_step=dt * 1e3 / max(1, steps), cpu_baseline=cpu,
                config=dict(workload=name, correspondences=n_full, k=k, sample_correspondences=p.n,
                            lm_iters_per_step=lm_iters))
