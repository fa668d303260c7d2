"""CPU legs of bench.py (oracle = test infrastructure; this is the one place bench.py may execute it).

The reference itself (g2o + Eigen + Sophus + Qhull + Open3D) cannot be compiled in this image, so both legs time the
plain-C oracle (oracle/c/dsc_oracle.c: analytic Jacobians, matrix-free block-Jacobi PCG, OpenMP over all host cores;
kind = "port").  The reference is single-threaded, differentiates the ARAP edges numerically (37 energy evaluations
per edge) and factorises the system: the port is the FASTER CPU program, i.e. the ratio reported against it is
conservative.

  run_reference   `bench.py --impl reference`: the SAME work as the GPU arm -- the same frame pair at its full size
                  (same generator, same number of correspondences, same k), the same number of LM iterations, every
                  linear solve to the same PCG tolerance (g2o's policy: no early rejection of trials).  One step is
                  timed, whatever --steps says (a CPU step takes minutes); if the time budget runs out first, the
                  iterations completed so far are reported and the record says so.
  components      bench.py's cpu_baseline on the GPU arm: a bounded sample (about 10-30 s) of the same full-size pair --
                  one linearisation, a few dozen PCG iterations, one cost evaluation -- composed with the iteration
                  counts the GPU arm measured under the policy-identical schedule (every solve to the tight tolerance).
"""
import os
import time

import numpy as np

from . import camera, edges
from .f32 import Pose
from .graph import Graph
from .se3 import SE3


def problem_from_arrays(sc, prob):
    """oracle Problem from the arrays bench.py's prepare() holds (triangulated points, graph, initial scales)"""
    cam = (camera.KB8, np.asarray(sc["cam"], np.float32))
    T1, T2 = Pose.from34(sc["T1"]), Pose.from34(sc["T2"])
    n = len(prob["X1"])
    g = Graph(np.ascontiguousarray(prob["rowptr"], np.int32), np.ascontiguousarray(prob["col"], np.int32),
              np.ascontiguousarray(prob["w"], np.float64), float(prob["area"]), int(prob["ntri"]))
    return edges.Problem(cam1=cam, cam2=cam, T1=T1, T2=T2, uv1=prob["uv1"], uv2=prob["uv2"], inv_sigma2_1=np.ones(n),
                         inv_sigma2_2=np.ones(n), d1=np.asarray(prob["d1"], np.float64), d2=np.asarray(prob["d2"], np.float64),
                         graph=g, X1=np.asarray(prob["X1"], np.float64), X2=np.asarray(prob["X2"], np.float64), Tg=SE3(),
                         s1=float(prob["s1"]), s2=float(prob["s2"]))


def _how(used, cores, rtol):
    return (f"C oracle (oracle/c/dsc_oracle.c: analytic Jacobians, matrix-free block-Jacobi PCG to rtol {rtol:g}, OpenMP), "
            f"{used} threads of {cores} host cores; the reference itself is single-threaded, differentiates numerically and "
            f"factorises (cannot be built here)")


def components(sc, prob, weights, counts, pcg_rtol, pcg_reps=24):
    """cpu_baseline of the GPU arm.  counts = dict(lm_iters, trials, pcg_iters) of ONE step under the full-solve policy."""
    from . import cport
    cport.build()
    p = problem_from_arrays(sc, prob)
    cp = cport.CProblem(p)
    w = edges.Weights(**weights)
    t0 = time.perf_counter()
    cport.compute_rotations(cp)
    t_rot = time.perf_counter() - t0
    tc = cport.time_components(cp, w, pcg_reps)
    sample_s = time.perf_counter() - t0
    step_s = (t_rot + counts["lm_iters"] * tc["linearize_s"] + counts["pcg_iters"] * tc["pcg_iter_s"]
              + counts["trials"] * (tc["cost_s"] + tc["precond_s"]))
    used, cores = cport.threads(), os.cpu_count() or 1
    return dict(value=counts["lm_iters"] / step_s, unit="LM it/s", cores=used, kind="port",
                seconds_per_step=step_s, components_s=dict(rotations=t_rot, **tc),
                sample=(f"the same {p.n}-correspondence pair (E = {p.graph.n_edges}): computeR + 1 linearisation + {pcg_reps} PCG "
                        f"iterations + 1 cost evaluation timed in {sample_s:.1f} s, composed with the step's counts under g2o's "
                        f"policy (every solve to rtol {pcg_rtol:g}): {counts['lm_iters']} LM iterations, {counts['trials']} trials, "
                        f"{counts['pcg_iters']} PCG iterations; {_how(used, cores, pcg_rtol)}"))


def run_reference(workload, n, k, lm_iters, pcg_rtol=1e-10, budget_s=900.0, seed=0):
    """`bench.py --impl reference`: one full step of the same work on the host cores."""
    import importlib
    import __graft_entry__ as g
    from . import cport
    from .triangulate import triangulate_pairs, init_depth_scale_sim, GATE_SIM
    wl = importlib.import_module(g.package().__name__ + ".workloads")      # input synthesis only
    cport.build()
    sc = wl.make_scene(workload, n, seed)
    cam = (camera.KB8, np.asarray(sc["cam"], np.float32))
    T1, T2 = Pose.from34(sc["T1"]), Pose.from34(sc["T2"])
    t0 = time.perf_counter()
    X1, X2, valid, _ = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, T1, T2, "NRSLAM", "FarPoints", GATE_SIM, sc["min_cos"])
    t_tri = time.perf_counter() - t0
    tri_rate = 2.0 * len(sc["uv1"]) / t_tri
    idx = np.nonzero(valid)[0][:n]
    if len(idx) < n:
        raise RuntimeError(f"only {len(idx)} valid correspondences of {n}")
    ones = np.ones(n, bool)
    prob = dict(X1=X1[idx], X2=X2[idx], uv1=sc["uv1"][idx], uv2=sc["uv2"][idx], d1=sc["d1"][idx], d2=sc["d2"][idx], area=sc["area"],
                ntri=2 * n, s1=init_depth_scale_sim(sc["d1"][idx], X1[idx], T1, ones), s2=init_depth_scale_sim(sc["d2"][idx], X2[idx], T2, ones))
    prob["rowptr"], prob["col"], prob["w"] = wl.knn_graph(prob["X1"][:, :2].astype(np.float64), k)      # untimed, as on the GPU arm
    p = problem_from_arrays(sc, prob)
    w = edges.Weights(**sc["weights"])
    iters = lm_iters or sc["lm_iters"]
    cp = cport.CProblem(p)
    t0 = time.perf_counter()
    cport.compute_rotations(cp)                              # computeR is part of arapOptimization
    tr = cport.optimize(cp, w, iters, threads=0, pcg_rtol=pcg_rtol, pcg_max=200000, budget_s=budget_s)
    dt = time.perf_counter() - t0
    done = len(tr["chi2"])
    used, cores = cport.threads(), os.cpu_count() or 1
    window = (f"all {iters} LM iterations of the step" if done == iters else
              f"the FIRST {done} of {iters} LM iterations (time budget {budget_s:.0f} s reached; the first iterations are the "
              f"expensive ones, so this window is pessimistic for the CPU)")
    cpu = dict(value=done / dt, unit="LM it/s", cores=used, kind="port", triangulated_points_per_s=tri_rate,
               sample=(f"{window} on the same {p.n}-correspondence pair (k = {k}, E = {p.graph.n_edges}): {sum(tr['trials'])} trials, "
                       f"{sum(tr['pcg_iters'])} PCG iterations in {dt:.1f} s; {_how(used, cores, pcg_rtol)}; triangulation: "
                       f"vectorised numpy oracle, {len(sc['uv1'])} matches in {t_tri:.1f} s, 1 thread"))
    return dict(value=done / dt, ms_per_step=dt * 1e3, cpu_baseline=cpu, lm_iters_done=done,
                trace=dict(chi2=tr["chi2"], trials=tr["trials"], pcg_iters=tr["pcg_iters"], final_chi2=tr["final_chi2"]),
                config=dict(workload=sc["name"], correspondences=p.n, directed_edges=int(p.graph.n_edges), k=k,
                            lm_iters_per_step=iters, pcg_rtol=pcg_rtol, pcg_iters_per_lm_iter=sum(tr["pcg_iters"]) / max(1, done),
                            lm_trials_per_step=sum(tr["trials"]), frame_pairs=1, solve_policy="every trial solved to pcg_rtol (g2o)",
                            steps_timed=1, lm_iters_timed=done))
