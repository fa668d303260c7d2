"""Parity at the sizes BASELINE.json names (configs 2, 3, 4), through the C ABI.

  config 2   100k-correspondence sheet, k = 8: the plain-C oracle solves the same pair live (3 LM iterations, PCG 1e-12)
  config 4   RealColon-shaped pair: distorted Kannala-Brandt camera + border mask + k = 16 together, 40k live, and the
             1M pair against a committed oracle trace
  config 3   Drunkard-shaped 1M pair against a committed oracle trace (cost, gradient, operator, 30 LM iterations)
The frame pair is prepared twice: by the CUDA path (K1 triangulation, k-NN graph, initial depth scales on the device) and
by the oracle on the CPU (tests/fullsize.py); both preparations must agree bit for bit before the solvers are compared.
Tolerance: north_star's fp64 bar, 1e-5 relative on per-iteration cost and final points.
"""
import json
import os

import numpy as np
import pytest

import fullsize
from oracle import edges

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _prepare_on_gpu(pkg, ctx, sc, n, k):
    """bench.py's prepare(): triangulate, keep the first n valid matches, k-NN graph, initial scales -- all on the device"""
    cam = (0, sc["cam"])
    pair = pkg.make_pair(cam, cam, sc["T1"], sc["T2"])
    prm = ctx.tri_params("NRSLAM", "FarPoints", 1, sc["min_cos"])
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
    idx = np.nonzero(valid)[0][:n]
    assert len(idx) == n
    rowptr, col, w = ctx.knn_graph(X1[idx], k)
    ctx.tri_upload(pair, sc["uv1"][idx], sc["uv2"][idx], sc["d1"][idx], sc["d2"][idx])
    ctx.tri_run(prm)
    s1, s2 = ctx.depth_scale_init(1), ctx.depth_scale_init(2)
    return dict(pair=pair, idx=idx, X1=X1[idx], X2=X2[idx], rowptr=rowptr, col=col, w=w, s1=s1, s2=s2)


def _assert_same_preparation(g, p, idx):
    assert np.array_equal(g["idx"], idx), "validity gates differ"
    assert np.array_equal(g["X1"].astype(np.float64), p.X1) and np.array_equal(g["X2"].astype(np.float64), p.X2), "triangulated points differ"
    assert np.array_equal(g["rowptr"], p.graph.rowptr) and np.array_equal(g["col"], p.graph.col), "neighbour graph differs"
    assert g["s1"] == pytest.approx(p.s1, rel=1e-12) and g["s2"] == pytest.approx(p.s2, rel=1e-12)


def _upload_prepared(ctx, sc, g, p):
    ctx.problem_upload(g["pair"], g["X1"], g["X2"], p.uv1, p.uv2, p.d1, p.d2, scale1=g["s1"], scale2=g["s2"])
    ctx.set_graph(g["rowptr"], g["col"], g["w"], sc["area"], 2 * p.n, 1)
    ctx.compute_rotations()


def _compare_with_live_oracle(pkg, ctx, workload, n, k, iters):
    wl = fullsize.workloads()
    sc = wl.make_scene(workload, n, 0)
    p, idx = fullsize.oracle_problem(sc, n, k, rotations=False)
    g = _prepare_on_gpu(pkg, ctx, sc, n, k)
    _assert_same_preparation(g, p, idx)
    _upload_prepared(ctx, sc, g, p)
    # rotations: the CUDA computeR against the C oracle's (the latter squares the condition number), then shared
    from oracle import cport, se3
    q = ctx.get_rotations()
    p.R = np.stack([se3.quat_to_rot(qi) for qi in q])
    cq = cport.CProblem(p)
    bad = cport.compute_rotations(cq)
    assert bad <= n // 1000 and np.median(np.abs(cq.R - p.R).reshape(n, -1).max(1)) < 1e-9
    w = edges.Weights(**sc["weights"])
    gw = pkg.make_weights(w.rep, w.arap, w.depth_sigma)
    cp = cport.CProblem(p, rotations=p.R)
    c0, parts = cport.cost(cp, w)
    g0, gparts = ctx.cost(gw)
    assert g0 == pytest.approx(c0, rel=1e-11)
    for a, b in zip(gparts, parts):
        assert a == pytest.approx(b, rel=1e-10)
    m = 8 + 6 * n
    x = ((np.arange(m) * 0.6180339887498949) % 1.0) - 0.5
    cb, chd, cy0, cchi = cport.debug_linearize(cp, w, 0.0, x)
    gb, ghd, gchi = ctx.debug_linearize(gw)
    lam = 1e-5 * np.abs(chd).max()
    gy = ctx.debug_matvec(gw, lam, x)
    np.testing.assert_allclose(gb, cb, rtol=1e-9, atol=1e-9 * np.abs(cb).max())
    np.testing.assert_allclose(ghd, chd, rtol=1e-9, atol=1e-12 * np.abs(chd).max())
    cy = cy0 + lam * x
    np.testing.assert_allclose(gy, cy, rtol=1e-9, atol=1e-10 * np.abs(cy).max())
    ctx.set_solver(1)
    ctx.set_pcg(rtol=1e-12, max_iters=100000, check_every=64)
    recs, st = ctx.optimize(gw, iters)
    out = ctx.download()
    tr = cport.optimize(cp, w, iters, pcg_rtol=1e-12, pcg_max=100000)
    assert st.iterations == len(tr["chi2"]) == iters
    for r, c, lam_o, tq in zip(recs, tr["chi2"], tr["lam"], tr["trials"]):
        assert r.chi2_before == pytest.approx(c, rel=1e-5)
        assert r.lam == pytest.approx(lam_o, rel=1e-4)
        assert r.trials == tq
    assert st.final_chi2 == pytest.approx(tr["final_chi2"], rel=1e-5)
    scale = np.abs(np.concatenate([cp.X1, cp.X2])).max()
    assert np.abs(out["X1d"] - cp.X1).max() <= 1e-5 * scale
    assert np.abs(out["X2d"] - cp.X2).max() <= 1e-5 * scale
    s = cp.scales()
    assert out["scales"][0] == pytest.approx(s[0], rel=1e-5) and out["scales"][1] == pytest.approx(s[1], rel=1e-5)
    np.testing.assert_allclose(out["Tg"], cp.Tg7(), atol=1e-5)
    return st


def test_config2_sheet_100k_against_the_c_oracle(pkg, ctx):
    """BASELINE.json configs[1] at its full size: 100k correspondences, k = 8, Simulation.yaml intrinsics."""
    _compare_with_live_oracle(pkg, ctx, "sheet", 100_000, 8, 3)


def test_config4_shape_mask_k16_distortion_40k_against_the_c_oracle(pkg, ctx):
    """Config 4's ingredients TOGETHER (Realcolon.yaml Kannala-Brandt distortion, border mask, k = 16, sigma_depth 1e-6)
    at a size the C oracle solves in seconds."""
    wl = fullsize.workloads()
    sc = wl.make_scene("realcolon", 40_000, 0)
    full = wl.tube_scene(int(40_000 * 1.08) + 64, seed=0, cam=wl.REALCOLON_CAM, arap=0.1, depth_sigma=1e-6, scales=(1.0, 1.0))
    assert len(sc["uv1"]) < len(full["uv1"])                       # the border mask removed matches
    _compare_with_live_oracle(pkg, ctx, "realcolon", 40_000, 16, 4)


def _compare_with_golden(pkg, ctx, name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated (tests/golden/make_fullsize_traces.py)")
    gold = json.load(open(path))
    n, k = gold["n"], gold["k"]
    wl = fullsize.workloads()
    sc = wl.make_scene(gold["workload"], n, gold["seed"])
    g = _prepare_on_gpu(pkg, ctx, sc, n, k)
    idx = g["idx"]
    assert len(g["col"]) == gold["directed_edges"]
    assert g["s1"] == pytest.approx(gold["s1"], rel=1e-9) and g["s2"] == pytest.approx(gold["s2"], rel=1e-9)
    ctx.problem_upload(g["pair"], g["X1"], g["X2"], sc["uv1"][idx], sc["uv2"][idx], sc["d1"][idx].astype(np.float64),
                       sc["d2"][idx].astype(np.float64), scale1=gold["s1"], scale2=gold["s2"])
    ctx.set_graph(g["rowptr"], g["col"], g["w"], sc["area"], 2 * n, 1)
    ctx.compute_rotations()
    gw = pkg.make_weights(**gold["weights"])
    chi, parts = ctx.cost(gw)
    assert chi == pytest.approx(gold["chi2_initial"], rel=1e-9)
    for a, b in zip(parts, gold["chi2_parts_initial"]):
        assert a == pytest.approx(b, rel=1e-8)
    # gradient, Hessian diagonal and operator on the committed sample rows + global checksums
    m = 8 + 6 * n
    x = ((np.arange(m) * 0.6180339887498949) % 1.0) - 0.5
    gb, ghd, gchi = ctx.debug_linearize(gw)
    rows = np.asarray(gold["sample_rows"])
    bs, hs, ys = np.asarray(gold["b_sample"]), np.asarray(gold["hdiag_sample"]), np.asarray(gold["y_sample"])
    assert float(np.linalg.norm(gb)) == pytest.approx(gold["b_norm"], rel=1e-8)
    assert float(gb @ x) == pytest.approx(gold["b_dot_probe"], rel=1e-6, abs=1e-9 * gold["b_norm"] * np.linalg.norm(x))
    assert float(ghd.sum()) == pytest.approx(gold["hdiag_sum"], rel=1e-8)
    np.testing.assert_allclose(gb[rows], bs, rtol=1e-8, atol=1e-9 * np.abs(bs).max())
    np.testing.assert_allclose(ghd[rows], hs, rtol=1e-8, atol=1e-12 * np.abs(hs).max())
    gy = ctx.debug_matvec(gw, gold["lambda_probe"], x)
    assert float(np.linalg.norm(gy)) == pytest.approx(gold["y_norm"], rel=1e-8)
    np.testing.assert_allclose(gy[rows], ys, rtol=1e-8, atol=1e-10 * np.abs(ys).max())
    # the LM trace, every solve to the oracle's tolerance
    iters = len(gold["chi2"])
    ctx.set_solver(1)
    ctx.set_pcg(rtol=gold["pcg_rtol"], max_iters=200000, check_every=64)
    recs, st = ctx.optimize(gw, iters)
    out = ctx.download()
    assert st.iterations == iters
    for r, c, lam_o, tq in zip(recs, gold["chi2"], gold["lam"], gold["trials"]):
        assert r.chi2_before == pytest.approx(c, rel=1e-5)
        assert r.lam == pytest.approx(lam_o, rel=1e-4)
        assert r.trials == tq
    assert st.final_chi2 == pytest.approx(gold["final_chi2"], rel=1e-5)
    sel = np.asarray(gold["sample_points"])
    X1s, X2s = np.asarray(gold["X1_sample"]), np.asarray(gold["X2_sample"])
    scale = max(np.abs(X1s).max(), np.abs(X2s).max())
    assert np.abs(out["X1d"][sel] - X1s).max() <= 1e-5 * scale
    assert np.abs(out["X2d"][sel] - X2s).max() <= 1e-5 * scale
    assert float(out["X1d"].sum()) == pytest.approx(gold["X1_sum"], rel=1e-6)
    assert out["scales"][0] == pytest.approx(gold["scales"][0], rel=1e-5) and out["scales"][1] == pytest.approx(gold["scales"][1], rel=1e-5)
    np.testing.assert_allclose(out["Tg"], gold["Tg"], atol=1e-5)
    # the bench's own settings (PCG 1e-10, early rejection of bad trials) must walk the same trajectory
    ctx.reset_state()
    ctx.set_pcg(rtol=1e-10, max_iters=6000, check_every=64)
    ctx.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
    recs2, st2 = ctx.optimize(gw, iters)
    assert [r.trials for r in recs2] == gold["trials"]
    for r, c in zip(recs2, gold["chi2"]):
        assert r.chi2_before == pytest.approx(c, rel=1e-5)
    assert st2.final_chi2 == pytest.approx(gold["final_chi2"], rel=1e-5)
    return st, st2


def test_config3_drunkard_1m_against_the_committed_oracle_trace(pkg, ctx):
    """BASELINE.json configs[2] at its full size (the bench workload): 1M correspondences, k = 8, 30 LM iterations."""
    _compare_with_golden(pkg, ctx, "fullsize_drunkard_1000000_k8.json")


def test_config4_realcolon_1m_k16_against_the_committed_oracle_trace(pkg, ctx):
    """BASELINE.json configs[3] at its full size: border mask, 1M surviving correspondences, k = 16, distortion."""
    _compare_with_golden(pkg, ctx, "fullsize_realcolon_1000000_k16.json")
