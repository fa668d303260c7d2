"""Full-size parity helpers (test infrastructure): the bench workloads (BASELINE.json configs 2-4) prepared entirely on
the CPU by the oracle, so that the C oracle can run the same frame pair the CUDA path gets.

  oracle_problem   scene -> oracle triangulation (bit-identical to K1) -> first n valid matches -> symmetrised k-NN
                   graph (k-d tree; bit-identical to dsc_knn_build) -> initial depth scales -> rotations (numpy SVD)
  run_c_oracle     LM trace of oracle/c/dsc_oracle.c on it
The golden traces of the 1M configurations (tests/golden/fullsize_*.json) are produced with these by
tests/golden/make_fullsize_traces.py; the 100k configuration is solved live inside the GPU test.
"""
import hashlib
import importlib

import numpy as np

from oracle import camera, edges, graph as ograph
from oracle.f32 import Pose
from oracle.se3 import SE3
from oracle.triangulate import triangulate_pairs, init_depth_scale_sim, GATE_SIM


def workloads():
    import __graft_entry__ as g
    return importlib.import_module(g.package().__name__ + ".workloads")


def oracle_problem(sc, n, k, rotations=True):
    wl = workloads()
    cam = (camera.KB8, np.asarray(sc["cam"], np.float32))
    T1, T2 = Pose.from34(sc["T1"]), Pose.from34(sc["T2"])
    X1, X2, valid, _ = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, T1, T2, "NRSLAM", "FarPoints", GATE_SIM, sc["min_cos"])
    idx = np.nonzero(valid)[0][:n]
    if len(idx) < n:
        raise RuntimeError(f"only {len(idx)} valid correspondences of {n}")
    X1, X2 = X1[idx], X2[idx]
    d1, d2 = sc["d1"][idx], sc["d2"][idx]
    s1 = init_depth_scale_sim(d1, X1, T1, np.ones(n, bool))
    s2 = init_depth_scale_sim(d2, X2, T2, np.ones(n, bool))
    rowptr, col, w = wl.knn_graph(X1[:, :2].astype(np.float64), k)
    g = ograph.Graph(rowptr, col, w, float(sc["area"]), 2 * n)
    p = edges.Problem(cam1=cam, cam2=cam, T1=T1, T2=T2, uv1=sc["uv1"][idx], uv2=sc["uv2"][idx], inv_sigma2_1=np.ones(n),
                      inv_sigma2_2=np.ones(n), d1=d1.astype(np.float64), d2=d2.astype(np.float64), graph=g,
                      X1=X1.astype(np.float64), X2=X2.astype(np.float64), Tg=SE3(), s1=float(s1), s2=float(s2))
    if rotations:
        p.R = ograph.compute_rotations(g, p.X1, p.X2)
    return p, idx


def fingerprint(p):
    """short hash of the problem's inputs: tells a changed generator from a changed solver when a golden trace fails"""
    h = hashlib.sha256()
    for a in (p.uv1, p.uv2, p.d1, p.d2, p.X1, p.X2, p.graph.rowptr, p.graph.col):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def run_c_oracle(p, w, iters, pcg_rtol=1e-12, pcg_max=60000, threads=0):
    from oracle import cport
    cp = cport.CProblem(p, rotations=p.R)
    c0, parts = cport.cost(cp, w)
    tr = cport.optimize(cp, w, iters, threads=threads, pcg_rtol=pcg_rtol, pcg_max=pcg_max)
    tr.update(chi2_parts=list(parts), X1=cp.X1, X2=cp.X2, scales=list(cp.scales()), Tg=cp.Tg7().tolist())
    return tr


def upload_gpu(pkg, ctx, p, reorder=1):
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, p.inv_sigma2_1, p.inv_sigma2_2, scale1=p.s1, scale2=p.s2,
                       Tg7=p.Tg.as7())
    g = p.graph
    ctx.set_graph(g.rowptr, g.col, g.w, g.area, g.n_triangles, reorder)
    ctx.compute_rotations()
