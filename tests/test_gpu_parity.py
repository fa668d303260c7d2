"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): float32 triangulation bit-exact for the closed-form
methods; per-iteration cost and final points within 1e-5 relative in fp64; index bookkeeping exact.
"""
import numpy as np
import pytest

from oracle import scenes, edges, lm, camera, graph as ograph
from oracle.triangulate import triangulate_pairs, init_depth_scale_sim, GATE_SIM, GATE_REAL, GATE_NONE

pytestmark = pytest.mark.gpu

REF = "/root/reference/Data/"


def _pair(pkg, sc):
    cam = (camera.KB8, sc["cam"])
    return pkg.make_pair(cam, cam, sc["T1"], sc["T2"]), cam


def _upload(pkg, ctx, p, reorder=1):
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, p.inv_sigma2_1, p.inv_sigma2_2,
                       scale1=p.s1, scale2=p.s2, Tg7=p.Tg.as7())
    g = p.graph
    ctx.set_graph(g.rowptr, g.col, g.w, g.area, g.n_triangles, reorder)
    ctx.compute_rotations()


def _w(pkg, w):
    return pkg.make_weights(w.rep, w.arap, w.depth_sigma, w.glob, w.alpha, w.beta)


@pytest.mark.parametrize("method", ["NRSLAM", "DepthMeasurement"])
@pytest.mark.parametrize("location", ["InRays", "TwoPoints", "FarPoints"])
def test_triangulation_closed_form_bit_exact(pkg, ctx, method, location):
    sc = scenes.sheet_scene(20000, seed=1)
    pair, cam = _pair(pkg, sc)
    prm = ctx.tri_params(method, location, GATE_SIM, 0.9998)
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"], sc["d1"], sc["d2"])
    oX1, oX2, ovalid, ocos = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, sc["T1"], sc["T2"], method, location,
                                                GATE_SIM, 0.9998, d1=sc["d1"], d2=sc["d2"])
    assert np.array_equal(valid, ovalid)
    assert nv == int(ovalid.sum())
    assert np.array_equal(X1, oX1) and np.array_equal(X2, oX2)
    assert np.array_equal(cosp, ocos)


@pytest.mark.parametrize("method", ["Classic", "ORBSLAM"])
def test_triangulation_svd_methods(pkg, ctx, method):
    sc = scenes.sheet_scene(5000, seed=2)
    pair, cam = _pair(pkg, sc)
    for location in ("InRays", "TwoPoints"):
        prm = ctx.tri_params(method, location, GATE_NONE)
        X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
        oX1, oX2, ovalid, ocos = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, sc["T1"], sc["T2"], method, location, GATE_NONE)
        # singular vectors: fp64 Jacobi on the device vs LAPACK in the oracle, rounded to float32
        np.testing.assert_allclose(X1, oX1, rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(X2, oX2, rtol=2e-4, atol=2e-6)


def test_triangulation_real_gates_and_kb8_distortion(pkg, ctx):
    sc = scenes.tube_scene(8000, seed=4, cam=scenes.REALCOLON_CAM)
    pair, cam = _pair(pkg, sc)
    prm = ctx.tri_params("NRSLAM", "TwoPoints", GATE_REAL, 0.9998, depth_limit=0.25, check_reproj=True)
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
    oX1, oX2, ovalid, ocos = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, sc["T1"], sc["T2"], "NRSLAM", "TwoPoints",
                                                GATE_REAL, 0.9998, depth_limit=0.25, check_reproj=True)
    assert 0 < ovalid.sum() < len(ovalid)
    assert np.array_equal(valid, ovalid)
    assert np.array_equal(X1, oX1) and np.array_equal(X2, oX2)


def test_triangulation_empty_and_single(pkg, ctx):
    sc = scenes.sheet_scene(1, seed=0)
    pair, cam = _pair(pkg, sc)
    prm = ctx.tri_params()
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert X1.shape == (0, 3) and nv == 0
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
    oX1, oX2, ovalid, _ = triangulate_pairs(sc["uv1"], sc["uv2"], cam, cam, sc["T1"], sc["T2"])
    assert np.array_equal(X1, oX1) and np.array_equal(valid, ovalid)
    # principal point: unproject is defined as the optical axis
    uv = np.array([[sc["cam"][2], sc["cam"][3]]], np.float32)
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, uv, uv)
    oX1, oX2, ovalid, _ = triangulate_pairs(uv, uv, cam, cam, sc["T1"], sc["T2"])
    assert np.array_equal(X1, oX1, equal_nan=True)


def test_depth_scale_init(pkg, ctx):
    sc = scenes.sheet_scene(4000, seed=5)
    pair, cam = _pair(pkg, sc)
    ctx.tri_upload(pair, sc["uv1"], sc["uv2"], sc["d1"], sc["d2"])
    ctx.tri_run(ctx.tri_params())
    X1, X2, valid, _, _ = ctx.tri_download()
    for which, X, d, T in ((1, X1, sc["d1"], sc["T1"]), (2, X2, sc["d2"], sc["T2"])):
        s = ctx.depth_scale_init(which)
        assert s == pytest.approx(init_depth_scale_sim(d, X, T, valid), rel=1e-12)


@pytest.mark.parametrize("kind,n", [("knn", 3000), ("delaunay", 1500)])
def test_rotations_cost_gradient_and_operator(pkg, ctx, kind, n):
    sc = scenes.sheet_scene(n, seed=7)
    p, keep = scenes.problem_from_scene(sc, kind, 8)
    w = edges.Weights(rep=1.0, arap=3.0e3, depth_sigma=0.003)
    _upload(pkg, ctx, p)
    # K7 rotations
    q = ctx.get_rotations()
    from oracle.se3 import quat_to_rot
    R = np.stack([quat_to_rot(qq) for qq in q])
    np.testing.assert_allclose(R, p.R, atol=1e-9)
    ctx.set_rotations(np.stack([__import__("oracle.se3", fromlist=["x"]).rot_to_quat(r) for r in p.R]))
    # K6 cost
    st = edges.state_of(p)
    chi, parts = edges.total_cost(p, w, st, parts=True)
    gchi, gparts = ctx.cost(_w(pkg, w))
    assert gchi == pytest.approx(chi, rel=1e-11)
    for a, b in zip(gparts, parts):
        assert a == pytest.approx(b, rel=1e-10)
    # K2/K3 gradient, Hessian diagonal; K4 operator
    import scipy.sparse as sp
    J, wt, e, chi_l = edges.linearize(p, w, st)
    JW = J.T @ sp.diags(wt)
    H = (JW @ J).tocsr()
    b = -(JW @ e)
    gb, ghd, gchi2 = ctx.debug_linearize(_w(pkg, w))
    assert gchi2 == pytest.approx(chi, rel=1e-11)
    np.testing.assert_allclose(gb, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())
    np.testing.assert_allclose(ghd, H.diagonal(), rtol=1e-9, atol=1e-12 * np.abs(H.diagonal()).max())
    rng = np.random.default_rng(0)
    x = rng.standard_normal(H.shape[0])
    lam = 1e-5 * np.abs(H.diagonal()).max()
    y = ctx.debug_matvec(_w(pkg, w), lam, x)
    yo = H @ x + lam * x
    np.testing.assert_allclose(y, yo, rtol=1e-9, atol=1e-10 * np.abs(yo).max())


def _compare_lm(pkg, ctx, p, w, iters, rtol_cost=1e-5, rtol_pts=1e-5, pcg_rtol=1e-12, reorder=1, solver=1):
    _upload(pkg, ctx, p, reorder)
    ctx.set_solver(solver)                 # 1 = PCG (the 1M path), 2 = dense Cholesky (small problems), 0 = auto
    ctx.set_pcg(rtol=pcg_rtol, max_iters=20000, check_every=64)
    recs, st = ctx.optimize(_w(pkg, w), iters)
    out = ctx.download()
    ost, otr = lm.optimize(p, w, iters)
    assert st.iterations == len(otr.chi2)
    for r, c, lam_o, q in zip(recs, otr.chi2, otr.lam, otr.trials):
        assert r.chi2_before == pytest.approx(c, rel=rtol_cost)
        assert r.lam == pytest.approx(lam_o, rel=1e-4)
        assert r.trials == q
    assert st.final_chi2 == pytest.approx(otr.final_chi2, rel=rtol_cost)
    scale = np.abs(np.concatenate([ost.X1, ost.X2])).max()
    assert np.abs(out["X1d"] - ost.X1).max() <= rtol_pts * scale
    assert np.abs(out["X2d"] - ost.X2).max() <= rtol_pts * scale
    assert out["scales"][0] == pytest.approx(ost.s1, rel=1e-5) and out["scales"][1] == pytest.approx(ost.s2, rel=1e-5)
    Tg = ost.Tg.as7()
    np.testing.assert_allclose(out["Tg"], Tg, atol=1e-5 * max(1.0, np.abs(Tg).max()))
    n1, n2, upd = lm.write_back(p, ost)
    assert out["update"] == pytest.approx(upd, rel=1e-4)
    return recs, st, out


def test_lm_config1_simulation(pkg, ctx):
    """BASELINE.json configs[0]: Data/original_points.csv -> moved_points.csv, Simulation.yaml, Delaunay graph."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden", "config1_points.npz")
    z = np.load(gold)
    fe = scenes.simulation_frontend(z["original"], z["moved"], (-0.10, 0.02, 0.12), (0.14, 0.01, 0.06))
    p, keep = scenes.build_problem(fe["uv1"], fe["uv2"], fe["d1"], fe["d2"], fe["cam"], fe["T1"], fe["T2"])
    w = edges.Weights(rep=1.0, arap=200000.0, depth_sigma=0.003, glob=50.0)
    _compare_lm(pkg, ctx, p, w, 25)


@pytest.mark.parametrize("arap", [2.0e5, 1.0e-2])
def test_lm_sheet_knn(pkg, ctx, arap):
    sc = scenes.sheet_scene(2500, seed=11)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = edges.Weights(rep=1.0, arap=arap, depth_sigma=0.003)
    _compare_lm(pkg, ctx, p, w, 6)


def test_lm_no_reorder_same_result(pkg, ctx):
    sc = scenes.sheet_scene(1500, seed=12)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = edges.Weights(rep=1.0, arap=50.0, depth_sigma=0.003)
    r1, s1, o1 = _compare_lm(pkg, ctx, p, w, 4, reorder=0)
    r2, s2, o2 = _compare_lm(pkg, ctx, p, w, 4, reorder=1)
    np.testing.assert_allclose(o1["X1d"], o2["X1d"], rtol=0, atol=1e-9)


def test_reset_state_and_pixel_sigma(pkg, ctx):
    sc = scenes.sheet_scene(1200, seed=13)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = edges.Weights(rep=1.0, arap=10.0, depth_sigma=0.003)
    _upload(pkg, ctx, p)
    s0 = ctx.pixel_sigma()
    assert s0[0] == pytest.approx(scenes.pixel_sigma(p.cam1, p.T1, p.X1, p.uv1), rel=1e-9)
    assert s0[1] == pytest.approx(scenes.pixel_sigma(p.cam2, p.T2, p.X2, p.uv2), rel=1e-9)
    c0, _ = ctx.cost(_w(pkg, w))
    ctx.optimize(_w(pkg, w), 2)
    c1, _ = ctx.cost(_w(pkg, w))
    assert c1 < c0
    ctx.reset_state()
    c2, _ = ctx.cost(_w(pkg, w))
    assert c2 == c0


def test_error_behaviour(pkg, ctx):
    sc = scenes.sheet_scene(300, seed=14)
    p, keep = scenes.problem_from_scene(sc, "knn", 6)
    w = pkg.make_weights(1.0, 1.0, 0.003)
    with pytest.raises(pkg.DscError) as e:
        ctx.cost(w)                      # nothing uploaded
    assert e.value.status == -4
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, scale1=p.s1, scale2=p.s2)
    g = p.graph
    bad_col = g.col.copy()
    bad_col[0] = (bad_col[0] + 7) % p.n       # breaks symmetry
    with pytest.raises(pkg.DscError) as e:
        ctx.set_graph(g.rowptr, bad_col, g.w, g.area, g.n_triangles)
    assert e.value.status == -7 and "symmetric" in str(e.value)
    for breaker, word in ((lambda c, ww: c.__setitem__(3, -1), "range"), (lambda c, ww: c.__setitem__(g.rowptr[5], 5), "self loop"),
                          (lambda c, ww: c.__setitem__(g.rowptr[9] + 1, c[g.rowptr[9]]), "duplicate"),
                          (lambda c, ww: ww.__setitem__(2, ww[2] + 1.0), "weights")):
        c2, w2 = g.col.copy(), g.w.copy()
        breaker(c2, w2)
        with pytest.raises(pkg.DscError) as e:
            ctx.set_graph(g.rowptr, c2, w2, g.area, g.n_triangles)
        assert e.value.status == -7 and word in str(e.value), str(e.value)
    ctx.set_graph(g.rowptr, g.col, g.w, g.area, g.n_triangles)
    ctx.compute_rotations()
    with pytest.raises(pkg.DscError) as e:
        ctx.optimize(pkg.make_weights(1.0, 1.0, 0.0), 1)    # sigma_depth = 0: 1/0 in the reference
    assert e.value.status == -1


def test_lm_tube_distorted_camera_and_ragged_size(pkg, ctx):
    """Config-3/4 shape at oracle size: Kannala-Brandt distortion, n not a multiple of the 32-row slices."""
    sc = scenes.tube_scene(1237, seed=21, cam=scenes.REALCOLON_CAM, depth_sigma=0.0003, scales=(1.3, 0.8))
    p, keep = scenes.problem_from_scene(sc, "knn", 8, min_cos=0.99999)
    assert p.n % 32 != 0
    w = edges.Weights(rep=1.0, arap=1.0e7, depth_sigma=0.0003)
    _compare_lm(pkg, ctx, p, w, 5)


def test_lm_k16_graph(pkg, ctx):
    sc = scenes.sheet_scene(900, seed=22)
    p, keep = scenes.problem_from_scene(sc, "knn", 16)
    w = edges.Weights(rep=1.0, arap=0.1, depth_sigma=0.003)
    _compare_lm(pkg, ctx, p, w, 4)


def test_isolated_vertices_and_tiny_problems(pkg, ctx):
    """Rows without neighbours (sliced-ELL width 0 / padding), a 3-correspondence problem, and an empty one."""
    sc = scenes.sheet_scene(70, seed=23)
    p, keep = scenes.problem_from_scene(sc, "knn", 4)
    g = p.graph
    # cut vertices 5 and 40 out of the graph (symmetrically)
    row = g.rows()
    keep_e = ~np.isin(row, [5, 40]) & ~np.isin(g.col, [5, 40])
    rp = np.zeros(p.n + 1, np.int32)
    np.add.at(rp, row[keep_e] + 1, 1)
    p.graph = ograph.Graph(np.cumsum(rp).astype(np.int32), g.col[keep_e], g.w[keep_e], g.area, g.n_triangles)
    p.R = ograph.compute_rotations(p.graph, p.X1, p.X2)
    w = edges.Weights(rep=1.0, arap=3.0, depth_sigma=0.003)
    _compare_lm(pkg, ctx, p, w, 3)
    # three correspondences, one triangle
    sc = scenes.sheet_scene(3, seed=24)
    p3, _ = scenes.problem_from_scene(sc, "knn", 2, gate=GATE_NONE)
    _compare_lm(pkg, ctx, p3, w, 2)
    # empty problem: every call is a no-op, nothing crashes
    pair = pkg.make_pair(p3.cam1, p3.cam2, p3.T1, p3.T2)
    z3, z2 = np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32)
    ctx.problem_upload(pair, z3, z3, z2, z2, np.zeros(0), np.zeros(0))
    ctx.set_graph(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0), 1.0, 0)
    ctx.compute_rotations()
    recs, st = ctx.optimize(pkg.make_weights(1.0, 1.0, 0.003), 3)
    assert st.iterations == 0
    out = ctx.download()
    assert out["X1"].shape == (0, 3) and out["update"] == 0.0


def test_full_size_properties_1m(pkg, ctx):
    """BASELINE config 3 size (1M correspondences): size-independent properties -- the operator is symmetric
    (x.Ay == y.Ax), z.(A z) > 0, the cost of the triangulated state is reproduced by two independent kernels
    (cost_kernel vs linearize_kernel), one LM iteration does not increase the cost, reset restores it exactly."""
    import importlib
    wl = importlib.import_module(pkg.__name__ + ".workloads")
    n = 1_000_000
    sc = wl.tube_scene(int(n * 1.08), seed=0, cam=wl.DRUNKARD_CAM, arap=1.0e7, depth_sigma=0.0003)
    cam = (0, sc["cam"])
    pair = pkg.make_pair(cam, cam, sc["T1"], sc["T2"])
    prm = ctx.tri_params("NRSLAM", "FarPoints", 1, sc["min_cos"])
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
    idx = np.nonzero(valid)[0][:n]
    assert len(idx) == n
    rowptr, col, wts = wl.knn_graph(X1[idx][:, :2].astype(np.float64), 8)
    ctx.problem_upload(pair, X1[idx], X2[idx], sc["uv1"][idx], sc["uv2"][idx], sc["d1"][idx].astype(np.float64),
                       sc["d2"][idx].astype(np.float64), scale1=1.3, scale2=0.8)
    ctx.set_graph(rowptr, col, wts, sc["area"], 2 * n, 1)
    ctx.compute_rotations()
    w = pkg.make_weights(**sc["weights"])
    chi, parts = ctx.cost(w)
    b, hd, chi_lin = ctx.debug_linearize(w)
    assert chi_lin == pytest.approx(chi, rel=1e-12) and np.isfinite(chi)
    assert np.all(hd >= 0)
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(8 + 6 * n), rng.standard_normal(8 + 6 * n)
    lam = 1e-5 * hd.max()
    Ax, Ay = ctx.debug_matvec(w, lam, x), ctx.debug_matvec(w, lam, y)
    assert float(y @ Ax) == pytest.approx(float(x @ Ay), rel=1e-9)
    assert float(x @ Ax) > 0
    ctx.set_pcg(rtol=1e-10, max_iters=6000, check_every=64)
    recs, st = ctx.optimize(w, 1)
    assert st.final_chi2 <= chi and recs[0].chi2_before == pytest.approx(chi, rel=1e-12)
    ctx.reset_state()
    chi2, _ = ctx.cost(w)
    assert chi2 == chi


def test_early_reject_keeps_the_trace(pkg, ctx):
    """dsc_set_early_reject pauses the solve at a loose tolerance and rejects clearly bad steps there; accepted
    steps are still solved to the tight tolerance, so the LM trace and the result must not change."""
    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "config1_points.npz"))
    fe = scenes.simulation_frontend(gold["original"], gold["moved"], (-0.10, 0.02, 0.12), (0.14, 0.01, 0.06))
    p, keep = scenes.build_problem(fe["uv1"], fe["uv2"], fe["d1"], fe["d2"], fe["cam"], fe["T1"], fe["T2"])
    w = _w(pkg, edges.Weights(rep=1.0, arap=200000.0, depth_sigma=0.003))
    _upload(pkg, ctx, p)
    ctx.set_solver(1)                       # the early rejection belongs to the PCG path
    ctx.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
    r0, s0 = ctx.optimize(w, 25)
    o0 = ctx.download()
    ctx.reset_state()
    ctx.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
    r1, s1 = ctx.optimize(w, 25)
    o1 = ctx.download()
    assert s1.early_rejects > 0 and s1.total_pcg_iters < s0.total_pcg_iters
    assert [r.trials for r in r1] == [r.trials for r in r0]
    assert [r.chi2_before for r in r1] == [r.chi2_before for r in r0]          # bit-identical accepted steps
    assert np.array_equal(o0["X1d"], o1["X1d"]) and np.array_equal(o0["X2d"], o1["X2d"])
    ctx.set_early_reject((), ())
    ctx.reset_state()
    r2, s2 = ctx.optimize(w, 25)
    assert s2.early_rejects == 0 and s2.total_pcg_iters == s0.total_pcg_iters


@pytest.mark.parametrize("n,k,seed", [(1, 4, 0), (5, 8, 1), (3000, 8, 2), (20000, 16, 3)])
def test_gpu_knn_graph_matches_kdtree(pkg, ctx, n, k, seed):
    """Neighbour-graph indexing must be bit-exact: the GPU grid search against the oracle's k-d tree graph."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.normal(0, 0.03, n), rng.normal(0, 0.01, n), rng.normal(0.2, 0.01, n)], 1).astype(np.float32)
    if n >= 3000:
        X[7] = X[11]                                   # a duplicate position (zero distance) must not break anything
    rowptr, col, w = ctx.knn_graph(X, k)
    if n <= k:
        # fewer points than neighbours: everybody is everybody's neighbour
        assert rowptr[-1] == n * (n - 1)
        return
    g = ograph.knn_graph(X.astype(np.float64), k, 1.0)
    assert np.array_equal(rowptr, g.rowptr)
    assert np.array_equal(col, g.col)


def test_gpu_knn_graph_on_a_clustered_tube(pkg, ctx):
    import importlib
    wl = importlib.import_module(pkg.__name__ + ".workloads")
    sc = wl.tube_scene(60000, seed=5)
    cam = (0, sc["cam"])
    pair = pkg.make_pair(cam, cam, sc["T1"], sc["T2"])
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, ctx.tri_params("NRSLAM", "FarPoints", 1, sc["min_cos"]), sc["uv1"], sc["uv2"])
    X = X1[valid]
    rowptr, col, w = ctx.knn_graph(X, 8)
    r2, c2, _ = wl.knn_graph(X[:, :2].astype(np.float64), 8)
    assert np.array_equal(rowptr, r2) and np.array_equal(col, c2)


@pytest.mark.gpu
def test_lm_matches_the_c_oracle_at_30k(pkg, ctx):
    """A size the numpy direct solve does not reach comfortably: the CUDA path against the plain-C oracle
    (oracle/c/dsc_oracle.c, OpenMP PCG), same rotations, PCG rtol 1e-12 on both sides."""
    from oracle import cport, se3
    sc = scenes.tube_scene(30011, seed=31, depth_sigma=0.0003)
    p, keep = scenes.problem_from_scene(sc, "knn", 8, min_cos=0.99999)
    w = edges.Weights(rep=1.0, arap=1.0e7, depth_sigma=0.0003)
    _upload(pkg, ctx, p)
    q = ctx.get_rotations()
    R = np.stack([se3.quat_to_rot(qi) for qi in q])
    cp = cport.CProblem(p, rotations=R)
    cq = cport.CProblem(p)
    assert cport.compute_rotations(cq) == 0
    assert np.abs(cq.R - R).max() < 1e-4                      # the C computeR squares the condition number
    c0, parts = cport.cost(cp, w)
    g0, gparts = ctx.cost(_w(pkg, w))
    assert g0 == pytest.approx(c0, rel=1e-11)
    ctx.set_pcg(rtol=1e-12, max_iters=40000, check_every=64)
    recs, st = ctx.optimize(_w(pkg, w), 3)
    out = ctx.download()
    tr = cport.optimize(cp, w, 3, pcg_rtol=1e-12, pcg_max=40000)
    assert st.iterations == len(tr["chi2"])
    for r, c, lam_o, tq in zip(recs, tr["chi2"], tr["lam"], tr["trials"]):
        assert r.chi2_before == pytest.approx(c, rel=1e-5)
        assert r.lam == pytest.approx(lam_o, rel=1e-4)
        assert r.trials == tq
    assert st.final_chi2 == pytest.approx(tr["final_chi2"], rel=1e-5)
    scale = np.abs(np.concatenate([cp.X1, cp.X2])).max()
    assert np.abs(out["X1d"] - cp.X1).max() <= 1e-5 * scale
    assert np.abs(out["X2d"] - cp.X2).max() <= 1e-5 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [2, 0])
def test_dense_solver_on_the_reference_sized_problem(pkg, ctx, solver):
    """Config 1 (120 points, Delaunay mesh): the dense Cholesky path (what DSC_SOLVER_AUTO picks at this size) against
    the direct-solve oracle -- both factorise the same matrix, as g2o's LinearSolverEigen does."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_points.npz"))
    fe = scenes.simulation_frontend(z["original"], z["moved"], (-0.10, 0.02, 0.12), (0.14, 0.01, 0.06))
    p, keep = scenes.build_problem(fe["uv1"], fe["uv2"], fe["d1"], fe["d2"], fe["cam"], fe["T1"], fe["T2"])
    w = edges.Weights(rep=1.0, arap=200000.0, depth_sigma=0.003, glob=50.0)
    recs, st, out = _compare_lm(pkg, ctx, p, w, 25, solver=solver)
    assert st.total_pcg_iters == 0
    ctx.set_solver(0)


@pytest.mark.gpu
def test_dense_and_pcg_agree_and_limits(pkg, ctx):
    sc = scenes.sheet_scene(700, seed=41)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = edges.Weights(rep=1.0, arap=50.0, depth_sigma=0.003)
    r1, s1, o1 = _compare_lm(pkg, ctx, p, w, 4, solver=1)
    r2, s2, o2 = _compare_lm(pkg, ctx, p, w, 4, solver=2)
    assert [r.trials for r in r1] == [r.trials for r in r2]
    scale = np.abs(o1["X1d"]).max()
    assert np.abs(o1["X1d"] - o2["X1d"]).max() <= 1e-6 * scale
    assert s2.total_pcg_iters == 0 and s1.total_pcg_iters > 0
    # the largest size the dense path accepts (DSC_DENSE_MAX = 1000 correspondences)
    sc = scenes.sheet_scene(1000, seed=43)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    assert 900 < p.n <= 1000
    _upload(pkg, ctx, p)
    ctx.set_solver(2)
    recs_d, st_d = ctx.optimize(_w(pkg, w), 2)
    od = ctx.download()
    ctx.reset_state()
    ctx.set_solver(1)
    ctx.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
    recs_p, st_p = ctx.optimize(_w(pkg, w), 2)
    op = ctx.download()
    assert [r.trials for r in recs_d] == [r.trials for r in recs_p]
    assert st_d.final_chi2 == pytest.approx(st_p.final_chi2, rel=1e-8)
    assert np.abs(od["X1d"] - op["X1d"]).max() <= 1e-6 * np.abs(op["X1d"]).max()
    # above DSC_DENSE_MAX the dense solver refuses; auto falls back to the PCG
    sc = scenes.sheet_scene(1300, seed=42)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    assert p.n > 1000
    _upload(pkg, ctx, p)
    ctx.set_solver(2)
    with pytest.raises(pkg.DscError) as e:
        ctx.optimize(_w(pkg, w), 1)
    assert e.value.status == -1
    ctx.set_solver(0)
    recs, st = ctx.optimize(_w(pkg, w), 1)
    assert st.total_pcg_iters > 0


def _check_linearisation(pkg, ctx, p, w):
    import scipy.sparse as sp
    st = edges.state_of(p)
    chi, parts = edges.total_cost(p, w, st, parts=True)
    gchi, gparts = ctx.cost(_w(pkg, w))
    assert gchi == pytest.approx(chi, rel=1e-11)
    for a, b in zip(gparts, parts):
        assert a == pytest.approx(b, rel=1e-10)
    J, wt, e, chi_l = edges.linearize(p, w, st)
    JW = J.T @ sp.diags(wt)
    H = (JW @ J).tocsr()
    b = -(JW @ e)
    gb, ghd, gchi2 = ctx.debug_linearize(_w(pkg, w))
    assert gchi2 == pytest.approx(chi, rel=1e-11)
    np.testing.assert_allclose(gb, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())
    np.testing.assert_allclose(ghd, H.diagonal(), rtol=1e-9, atol=1e-12 * np.abs(H.diagonal()).max())
    x = np.random.default_rng(0).standard_normal(H.shape[0])
    lam = 1e-5 * np.abs(H.diagonal()).max()
    y = ctx.debug_matvec(_w(pkg, w), lam, x)
    yo = H @ x + lam * x
    np.testing.assert_allclose(y, yo, rtol=1e-9, atol=1e-10 * np.abs(yo).max())
    return wt


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [1, 2])
def test_outliers_activate_the_huber_kernel(pkg, ctx, solver):
    """Gross reprojection outliers (g2o RobustKernelHuber, delta = sqrt(100.991), g2oBundleAdjustment.cc:631,783-785):
    rho'(chi2) < 1 on those edges must enter the cost, the gradient, the operator and the LM trace."""
    sc = scenes.sheet_scene(500, seed=51)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    rng = np.random.default_rng(5)
    bad = rng.choice(p.n, p.n // 12, replace=False)
    p.uv1 = p.uv1.copy(); p.uv2 = p.uv2.copy()
    p.uv1[bad] += rng.uniform(25, 80, (len(bad), 2)).astype(np.float32) * rng.choice([-1, 1], (len(bad), 2)).astype(np.float32)
    p.uv2[bad[::2]] += np.float32(40.0)
    w = edges.Weights(rep=1.0, arap=20.0, depth_sigma=0.003)
    _upload(pkg, ctx, p)
    wt = _check_linearisation(pkg, ctx, p, w)
    assert np.count_nonzero(wt[:4 * p.n] < 1.0) >= len(bad)          # the robust weight is active on the outliers
    _compare_lm(pkg, ctx, p, w, 5, solver=solver)


@pytest.mark.gpu
def test_non_positive_depth_scale_branch(pkg, ctx):
    """EdgeDepthCorrection multiplies the error by 500 when the scale vertex is <= 0 (g2oTypes.h:400-416)."""
    sc = scenes.sheet_scene(300, seed=52)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    p.s1 = -0.7
    w = edges.Weights(rep=1.0, arap=20.0, depth_sigma=0.05)
    _upload(pkg, ctx, p)
    _check_linearisation(pkg, ctx, p, w)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2500, 13000])
def test_cluster_pcg_matches_the_per_iteration_kernels(pkg, n):
    """Small problems run the whole PCG in one launch (dsc_small.cuh): a thread-block cluster up to 3 k correspondences, a
    cooperative launch over all SMs with grid barriers up to 120 k; DSC_NO_CLUSTER_PCG=1 keeps the per-iteration kernels of
    the 1M path.  Same algorithm, different partial-sum partition: the traces agree to rounding, and both agree with the
    oracle (the other tests)."""
    import os
    sc = scenes.tube_scene(n, seed=61, depth_sigma=0.0003)
    p, keep = scenes.problem_from_scene(sc, "knn", 8, min_cos=0.99999)
    w = edges.Weights(rep=1.0, arap=1.0e7, depth_sigma=0.0003)
    res = []
    for env in (None, "1"):
        if env is None:
            os.environ.pop("DSC_NO_CLUSTER_PCG", None)
        else:
            os.environ["DSC_NO_CLUSTER_PCG"] = env
        try:
            with pkg.Context(0) as c:
                _upload(pkg, c, p)
                c.set_solver(1)
                c.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
                c.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
                recs, st = c.optimize(_w(pkg, w), 4)
                res.append((recs, st, c.download()))
        finally:
            os.environ.pop("DSC_NO_CLUSTER_PCG", None)
    (r0, s0, o0), (r1, s1, o1) = res
    assert s0.kernel_launches < s1.kernel_launches / 5            # one launch per solve instead of two per iteration
    assert [r.trials for r in r0] == [r.trials for r in r1]
    for a, b in zip(r0, r1):
        assert a.chi2_after == pytest.approx(b.chi2_after, rel=1e-9)
    scale = np.abs(o0["X1d"]).max()
    assert np.abs(o0["X1d"] - o1["X1d"]).max() <= 1e-8 * scale


# ---------------------------------------------------------------------------------------------- reference-held numbers
def _reference_pins():
    import json
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    return json.load(open(os.path.join(gold, "reference_pins.json"))), np.load(os.path.join(gold, "reference_pins.npz"))


@pytest.mark.gpu
def test_reference_logs_reproduced_through_the_c_abi(pkg, ctx):
    """REFERENCE-HELD NUMBERS on the GPU.  The INITIAL MEASUREMENTS of Data/Experiments/**/Experiment.txt (written by
    the reference right after triangulation; tests/golden/make_reference_pins.py): mean / RMSE 3-D error and pixel
    sigma of the triangulated map points for the three seed locations, 12 database pairs, 6 printed digits each.
    The logs were written with the PinHole model, so this is also the PinHole parity test of K1 (unproject),
    pixel_sigma_kernel (project) and the gates; the CUDA points are compared bit for bit with the oracle's."""
    from oracle import metrics
    meta, arr = _reference_pins()
    checked = 0
    for m in meta:
        o, mv = arr[f"o{m['key']}"], arr[f"m{m['key']}"]
        fe = scenes.simulation_frontend(o, mv, m["C1"], m["C2"], model=camera.PINHOLE)
        cam = (camera.PINHOLE, fe["cam"])
        pair = pkg.make_pair(cam, cam, fe["T1"], fe["T2"])
        for loc, lg in m["logs"].items():
            X1, X2, valid, cosp, nv = ctx.triangulate(pair, ctx.tri_params("NRSLAM", loc, GATE_SIM, 0.9998), fe["uv1"], fe["uv2"])
            oX1, oX2, ovalid, ocos = triangulate_pairs(fe["uv1"], fe["uv2"], cam, cam, fe["T1"], fe["T2"], "NRSLAM", loc, GATE_SIM, 0.9998)
            assert np.array_equal(valid, ovalid) and np.array_equal(X1, oX1) and np.array_equal(X2, oX2) and np.array_equal(cosp, ocos)
            assert nv == lg["n_mapped"]
            _, av, rmse = metrics.sim_absolute_map_errors(X1[valid], X2[valid], o[valid], mv[valid])
            assert av == pytest.approx(lg["av_error"], rel=2e-5), (m["case"], loc)
            assert rmse == pytest.approx(lg["rmse"], rel=2e-5), (m["case"], loc)
            # calculatePixelsStandDev on the device (pixel_sigma_kernel) from the float map points
            k = int(valid.sum())
            ctx.problem_upload(pair, X1[valid], X2[valid], fe["uv1"][valid], fe["uv2"][valid], np.zeros(k), np.zeros(k))
            s = ctx.pixel_sigma()
            if loc != "InRays":
                assert s[0] == pytest.approx(lg["sigma_c1"], rel=4e-5) and s[1] == pytest.approx(lg["sigma_c2"], rel=4e-5)
            else:
                assert s[0] < 1e-4 and s[1] < 1e-4
            checked += 1
        # show-solution mode (Data/SinteticDataBase/**/Experiment.txt): ground-truth points as map points
        ctx.problem_upload(pair, o, mv, fe["uv1"], fe["uv2"], np.zeros(len(o)), np.zeros(len(o)))
        s = ctx.pixel_sigma()
        assert s[0] == pytest.approx(m["database"]["sigma_c1"], rel=1e-5) and s[1] == pytest.approx(m["database"]["sigma_c2"], rel=1e-5)
    assert checked >= 30


@pytest.mark.gpu
@pytest.mark.parametrize("method,location", [("NRSLAM", "FarPoints"), ("Classic", "InRays"), ("DepthMeasurement", "TwoPoints")])
def test_pinhole_camera_triangulation(pkg, ctx, method, location):
    """PinHole::unproject / project (PinHole.cc:25-40) on K1, including the real-image gates with the reprojection check."""
    sc = scenes.tube_scene(6000, seed=72)
    cam = (camera.PINHOLE, sc["cam"])
    c1, c2 = sc["T1"].apply(sc["original"]), sc["T2"].apply(sc["moved"])
    rng = np.random.default_rng(72)
    uv1 = (camera.pinhole_project(sc["cam"], c1) + rng.normal(0, 1, (len(c1), 2)).astype(np.float32)).astype(np.float32)
    uv2 = (camera.pinhole_project(sc["cam"], c2) + rng.normal(0, 1, (len(c1), 2)).astype(np.float32)).astype(np.float32)
    pair = pkg.make_pair(cam, cam, sc["T1"], sc["T2"])
    prm = ctx.tri_params(method, location, GATE_REAL, 0.9998, depth_limit=0.3, check_reproj=True)
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, uv1, uv2, sc["d1"], sc["d2"])
    oX1, oX2, ovalid, ocos = triangulate_pairs(uv1, uv2, cam, cam, sc["T1"], sc["T2"], method, location, GATE_REAL, 0.9998,
                                                depth_limit=0.3, check_reproj=True, d1=sc["d1"], d2=sc["d2"])
    assert 0 < ovalid.sum()
    if method == "Classic":
        np.testing.assert_allclose(X1, oX1, rtol=2e-4, atol=2e-6)
        assert np.mean(valid == ovalid) > 0.999
    else:
        assert np.array_equal(valid, ovalid)
        assert np.array_equal(X1, oX1) and np.array_equal(X2, oX2) and np.array_equal(cosp, ocos)


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [1, 2])
def test_lm_pinhole_camera(pkg, ctx, solver):
    """PinHole::project / projectJac (PinHole.cc:25-33,49-62) inside the refinement: cost, gradient, operator and the
    LM trace against the direct-solve oracle."""
    sc = scenes.sheet_scene(700, seed=73)
    n = len(sc["original"])
    c1, c2 = sc["T1"].apply(sc["original"]), sc["T2"].apply(sc["moved"])
    rng = np.random.default_rng(73)
    uv1 = (camera.pinhole_project(sc["cam"], c1) + rng.normal(0, 1, (n, 2)).astype(np.float32)).astype(np.float32)
    uv2 = (camera.pinhole_project(sc["cam"], c2) + rng.normal(0, 1, (n, 2)).astype(np.float32)).astype(np.float32)
    p, keep = scenes.build_problem(uv1, uv2, sc["d1"], sc["d2"], sc["cam"], sc["T1"], sc["T2"], graph_kind="knn", k=8,
                                   area=sc["area"], model=camera.PINHOLE)
    assert p.cam1[0] == camera.PINHOLE and p.n > 500
    w = edges.Weights(rep=1.0, arap=50.0, depth_sigma=0.003)
    _upload(pkg, ctx, p)
    _check_linearisation(pkg, ctx, p, w)
    _compare_lm(pkg, ctx, p, w, 4, solver=solver)


@pytest.mark.gpu
def test_triangulate_rays_entry_matches_the_oracle(pkg, ctx):
    """dsc_triangulate_rays = useTriangulationMethod with the reference's argument list (Geometry.h:66-69): rays, no
    camera, no gates -- including rays that point backwards (z <= 0), which the pixel entry cannot express."""
    from oracle import triangulate as otri
    from oracle.f32 import normalize3
    sc = scenes.sheet_scene(3000, seed=81)
    xn1 = normalize3(camera.kb8_unproject(sc["cam"], sc["uv1"]))
    xn2 = normalize3(camera.kb8_unproject(sc["cam"], sc["uv2"]))
    xn1[::7] *= np.float32(-1.0)                                   # backwards rays
    xn2[::11, 2] = np.float32(0.0)
    for loc_name, loc in otri.LOCATIONS.items():
        X1, X2 = ctx.triangulate_rays(sc["T1"], sc["T2"], "NRSLAM", loc_name, xn1, xn2)
        with np.errstate(all="ignore"):
            oX1, oX2 = otri.triangulate_nrslam(xn1, xn2, sc["T1"], sc["T2"], loc)
        assert np.array_equal(X1, oX1.astype(np.float32), equal_nan=True) and np.array_equal(X2, oX2.astype(np.float32), equal_nan=True)
    # DepthMeasurement: the "rays" are camera-frame points at the measured depth (Geometry.cc:189-214)
    c1, c2 = sc["T1"].apply(sc["original"]), sc["T2"].apply(sc["moved"])
    X1, X2 = ctx.triangulate_rays(sc["T1"], sc["T2"], "DepthMeasurement", "FarPoints", c1, c2)
    oX1, oX2 = otri.triangulate_depth(c1, c2, sc["T1"], sc["T2"], otri.LOCATIONS["FarPoints"])
    assert np.array_equal(X1, oX1.astype(np.float32)) and np.array_equal(X2, oX2.astype(np.float32))
    # one match, as the host shim calls it
    X1, X2 = ctx.triangulate_rays(sc["T1"], sc["T2"], "NRSLAM", "InRays", xn1[:1], xn2[:1])
    assert X1.shape == (1, 3)


@pytest.mark.gpu
def test_dense_solver_is_race_free_under_concurrent_load(pkg):
    """Round-1 advisor finding: dense_panel_kernel's block 0 overwrote the diagonal block A11 in the launch in which the
    other blocks still read it.  L11 now goes through a side buffer; with other streams keeping every SM busy (blocks of
    the panel launch start late and far apart) the dense LM run must stay bit-identical to the quiet one."""
    import threading
    sc = scenes.sheet_scene(560, seed=91)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = _w(pkg, edges.Weights(rep=1.0, arap=50.0, depth_sigma=0.003))
    big = scenes.tube_scene(120000, seed=92, depth_sigma=0.0003)
    pb, _ = scenes.problem_from_scene(big, "knn", 8, min_cos=0.99999)
    wb = _w(pkg, edges.Weights(rep=1.0, arap=1.0e7, depth_sigma=0.0003))
    with pkg.Context(0) as c, pkg.Context(0) as noisy1, pkg.Context(0) as noisy2:
        _upload(pkg, c, p)
        c.set_solver(2)
        recs0, st0 = c.optimize(w, 6)
        out0 = c.download()
        stop = threading.Event()

        def hammer(ctx_n):
            _upload(pkg, ctx_n, pb)
            ctx_n.set_solver(1)
            ctx_n.set_pcg(rtol=1e-10, max_iters=300, check_every=64)
            while not stop.is_set():
                ctx_n.reset_state()
                ctx_n.optimize(wb, 1)
        th = [threading.Thread(target=hammer, args=(x,)) for x in (noisy1, noisy2)]
        for t in th:
            t.start()
        try:
            for _ in range(8):
                c.reset_state()
                recs, st = c.optimize(w, 6)
                out = c.download()
                assert [r.chi2_after for r in recs] == [r.chi2_after for r in recs0]
                assert np.array_equal(out["X1d"], out0["X1d"]) and np.array_equal(out["X2d"], out0["X2d"])
        finally:
            stop.set()
            for t in th:
                t.join()
