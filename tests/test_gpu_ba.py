"""The classic bundle-adjustment paths on the device (include/dsc.h dsc_ba_*) against the oracle's restatement of
bundleAdjustment / localBundleAdjustment / poseOnlyOptimization (g2oBundleAdjustment.cc:38-444, oracle/ba.py).  The oracle solves
the full system (H + lambda I) directly; the device forms the Schur complement over the points (BlockSolver_6_3) -- the same
step.  Bar: LM decisions identical, costs to 1e-5 relative (north_star's fp64 bar), poses / points to 2e-6 absolute.  (Both sides project in
float32 from a double camera-frame point: a 1e-16 difference in the pose arithmetic flips the float rounding of a few
projections, ~3e-5 px each, so the costs agree to ~1e-8 at the start and ~1e-6 after several iterations.)"""
import numpy as np
import pytest

from oracle import ba as oba

pytestmark = pytest.mark.gpu


def _upload(pkg, b, p, points_fixed=False):
    b.upload(np.array([T.as7() for T in p.poses]), p.pose_fixed, p.cams, p.X, p.obs_pose, p.obs_point, p.obs_uv, p.obs_isg,
             points_fixed=points_fixed)


def _compare_traces(recs, otr):
    """LM decisions and costs, iteration by iteration, while the oracle still makes progress; once its cost moves by less than
    1e-6 relative per iteration the accept / reject decisions hang on the float32 rounding of single projections"""
    chi = otr["chi2"] + [otr["final_chi2"]]
    live = len(otr["trials"])
    for i in range(len(otr["trials"])):
        if abs(chi[i] - chi[i + 1]) <= 1e-6 * chi[i]:
            live = i
            break
    assert live >= 2 and len(recs) >= live
    assert [r.trials for r in recs[:live]] == otr["trials"][:live] and [bool(r.accepted) for r in recs[:live]] == otr["accepted"][:live]
    np.testing.assert_allclose([r.chi2_before for r in recs[:live]], otr["chi2"][:live], rtol=1e-5)
    np.testing.assert_allclose([r.lam for r in recs[:live]], otr["lam"][:live], rtol=1e-4)


def _check_state(p7, X, oposes, oX, b=None, p=None):
    """the estimates (only key frame 0 is fixed: the monocular scale stays a gauge freedom, along which two runs that differ in
    the last bits drift apart by ~1e-5 of the scene size) and, gauge-free, the chi2 of every edge"""
    for k, T in enumerate(oposes):
        np.testing.assert_allclose(p7[k], T.as7(), rtol=0, atol=5e-5)
    np.testing.assert_allclose(X, oX, rtol=0, atol=5e-5)
    if b is not None:
        chi, pos = b.edge_chi2()
        ochi, opos = oba.edge_chi2(p, oposes, oX)
        np.testing.assert_allclose(chi, ochi, rtol=2e-3, atol=2e-3)
        assert np.array_equal(pos, opos)


@pytest.mark.parametrize("n_points,n_poses,seed,outliers", [(300, 3, 1, 10), (800, 2, 2, 0), (500, 5, 3, 25)])
def test_bundle_adjustment_matches_the_oracle(pkg, n_points, n_poses, seed, outliers):
    p = oba.make_scene(n_points, n_poses, seed=seed, outliers=outliers)
    oposes, oX, otr = oba.optimize(p, 8, robust=True)
    with pkg.BundleAdjuster(0) as b:
        _upload(pkg, b, p)
        recs, st = b.optimize(8, huber_delta=oba.HUBER_2D)
        p7, X = b.download()
        assert st.kernel_launches > 0 and b.launch_count() >= st.kernel_launches
        _check_state(p7, X, oposes, oX, b, p)
    _compare_traces(recs, otr)
    assert st.final_chi2 == pytest.approx(otr["final_chi2"], rel=1e-5)
    assert otr["final_chi2"] < 0.2 * otr["chi2"][0]


def test_local_bundle_adjustment_flow(pkg):
    """5 robust iterations, chi2 > 5.991 or negative depth -> level 1, kernels off, 10 plain iterations, the same test again"""
    p = oba.make_scene(400, 4, seed=5, outliers=30)
    p.pose_fixed[3] = True                                         # a fixed key frame of the local map
    oposes, oX, oremoved, (tr1, tr2) = oba.local_bundle_adjustment(p)
    with pkg.BundleAdjuster(0) as b:
        _upload(pkg, b, p)
        r1, s1 = b.optimize(5, huber_delta=oba.HUBER_2D)
        chi_a, pos_a = b.edge_chi2()
        active = ~((chi_a > oba.CHI2_OUTLIER) | ~pos_a)
        b.set_levels(active)
        r2, s2 = b.optimize(10, huber_delta=0.0)
        chi_b, pos_b = b.edge_chi2()
        removed = (np.where(active, chi_b, chi_a) > oba.CHI2_OUTLIER) | ~pos_b
        p7, X = b.download()
    _compare_traces(r1, tr1)
    _compare_traces(r2, tr2)
    assert s1.final_chi2 == pytest.approx(tr1["final_chi2"], rel=1e-5) and s2.final_chi2 == pytest.approx(tr2["final_chi2"], rel=1e-5)
    assert np.array_equal(removed, oremoved) and 30 <= removed.sum() < len(removed)      # (30 planted outliers + edges not yet converged)
    _check_state(p7, X, oposes, oX)


def test_pose_only_optimization_flow(pkg):
    """one frame pose, points fixed, four rounds of ten iterations restarting from the frame's pose, inliers re-classified"""
    s = oba.make_scene(300, 2, seed=7, outliers=0, point_noise=0.0002)
    m = s.obs_pose == 1
    uv = s.obs_uv[m].copy()
    rng = np.random.default_rng(0)
    bad = rng.choice(len(uv), 40, replace=False)
    uv[bad] += rng.normal(0, 30.0, (40, 2)).astype(np.float32)
    p = oba.BaProblem(poses=[s.poses[1]], pose_fixed=np.array([False]), cams=[s.cams[1]], X=s.X, obs_pose=np.zeros(m.sum(), int),
                      obs_point=s.obs_point[m], obs_uv=uv, obs_isg=s.obs_isg[m], points_fixed=True)
    opose, oinl, ongood = oba.pose_only_optimization(p)
    O = len(p.obs_pose)
    with pkg.BundleAdjuster(0) as b:
        _upload(pkg, b, p, points_fixed=True)
        start = np.array([p.poses[0].as7()])
        inlier, level0 = np.ones(O, bool), np.ones(O, bool)
        stored, _ = b.edge_chi2()
        delta = oba.HUBER_2D
        for rnd in range(4):
            b.set_poses(start)
            b.set_levels(level0)
            b.optimize(10, huber_delta=delta)
            fresh, _ = b.edge_chi2()
            stored = np.where(level0, fresh, stored)
            if not inlier[rnd]:
                stored = fresh.copy()
            inlier = ~(stored > oba.CHI2_OUTLIER)
            level0 = inlier.copy()
            if rnd == 2:
                delta = 0.0
        p7, X = b.download()
    assert np.array_equal(inlier, oinl) and int(inlier.sum()) == ongood and 240 <= ongood <= 265
    np.testing.assert_allclose(p7[0], opose.as7(), rtol=0, atol=2e-6)
    np.testing.assert_array_equal(X, p.X)                          # the points are constants of the edges


def test_bundle_adjustment_edge_cases(pkg):
    p = oba.make_scene(50, 2, seed=9)
    with pkg.BundleAdjuster(0) as b:
        with pytest.raises(pkg.DscError):
            b.optimize(1)                                          # nothing uploaded
        # every pose fixed: only the points move (structure-only), no reduced system
        q = oba.BaProblem(poses=p.poses, pose_fixed=np.array([True, True]), cams=p.cams, X=p.X, obs_pose=p.obs_pose, obs_point=p.obs_point,
                          obs_uv=p.obs_uv, obs_isg=p.obs_isg)
        _upload(pkg, b, q)
        recs, st = b.optimize(4, huber_delta=oba.HUBER_2D)
        oposes, oX, otr = oba.optimize(q, 4, robust=True)
        _compare_traces(recs, otr)
        assert st.final_chi2 == pytest.approx(otr["final_chi2"], rel=1e-5)
        p7, X = b.download()
        _check_state(p7, X, oposes, oX)
        # a point nobody observes and an observation set to level 1 stay where they are / are ignored
        q2 = oba.BaProblem(poses=p.poses, pose_fixed=p.pose_fixed, cams=p.cams, X=np.concatenate([p.X, [[1.0, 2.0, 3.0]]]), obs_pose=p.obs_pose,
                           obs_point=p.obs_point, obs_uv=p.obs_uv, obs_isg=p.obs_isg)
        _upload(pkg, b, q2)
        act = np.ones(len(p.obs_pose), bool)
        act[::7] = False
        b.set_levels(act)
        recs, st = b.optimize(3, huber_delta=0.0)
        oposes, oX, otr = oba.optimize(q2, 3, active=act, robust=False)
        assert st.final_chi2 == pytest.approx(otr["final_chi2"], rel=1e-5)
        p7, X = b.download()
        assert np.array_equal(X[-1], [1.0, 2.0, 3.0])
        _check_state(p7, X, oposes, oX)
        with pytest.raises(pkg.DscError):                          # the same point twice from one pose
            b.upload(np.array([T.as7() for T in p.poses]), p.pose_fixed, p.cams, p.X, [0, 0], [3, 3], np.zeros((2, 2), np.float32))


def test_bundle_adjustment_two_views_200k_points(pkg):
    """the size the accelerated path exists for: two key frames, 200 k points -- cost falls monotonically, timing printed"""
    import time
    p = oba.make_scene(200_000, 2, seed=11, pose_noise=(0.003, 0.003), point_noise=0.001)
    with pkg.BundleAdjuster(0) as b:
        _upload(pkg, b, p)
        t0 = time.perf_counter()
        recs, st = b.optimize(10, huber_delta=oba.HUBER_2D)
        ms = (time.perf_counter() - t0) * 1e3
        chi = [r.chi2_before for r in recs] + [st.final_chi2]
        assert all(b2 <= a * (1 + 1e-12) for a, b2 in zip(chi, chi[1:])) and chi[-1] < 0.5 * chi[0]
        print(f"[ba 200k points, 2 views, {len(p.obs_pose)} observations] 10 LM iterations {ms:.1f} ms ({st.device_ms:.1f} ms on the device), "
              f"{st.total_trials} trials, {st.kernel_launches} launches, chi2 {chi[0]:.4e} -> {chi[-1]:.4e}")


# ---------------------------------------------------------------- the reference's entry points through the C++ shim
def _shim_flow(mode, p, octave, curr=0, min_common=0.0):
    import ctypes
    import os
    import __graft_entry__ as g
    lib_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "triangulation-in-deformable-scenes_b200", "lib")
    g.build()
    ctypes.CDLL(os.path.join(lib_dir, "libdsc_b200.so"), mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(os.path.join(lib_dir, "libdsc_host.so"))
    K, M, O = len(p.poses), len(p.X), len(p.obs_pose)
    pose34 = np.zeros((K, 12), np.float32)
    for k, T in enumerate(p.poses):
        pose34[k] = np.concatenate([T.R(), T.t[:, None]], 1).astype(np.float32).reshape(-1)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cam8 = np.ascontiguousarray(p.cams[0][1], np.float32)
    X = np.ascontiguousarray(p.X, np.float32)
    op, oj = np.ascontiguousarray(p.obs_pose, np.int32), np.ascontiguousarray(p.obs_point, np.int32)
    uv, oc = np.ascontiguousarray(p.obs_uv, np.float32), np.ascontiguousarray(octave, np.int32)
    p7, Xo, rem, ng = np.zeros((K, 7)), np.zeros((M, 3), np.float32), np.zeros(O, np.uint8), ctypes.c_int()
    rc = lib.dsch_ba_flow(mode, K, ptr(pose34), ptr(cam8), M, ptr(X), O, ptr(op), ptr(oj), ptr(uv), ptr(oc), curr, ctypes.c_float(min_common),
                          ptr(p7), ptr(Xo), ptr(rem), ctypes.byref(ng))
    assert rc == 0
    return p7, Xo, rem.astype(bool), ng.value


def _as_the_map_holds_it(p):
    """the Map stores float poses and float points, g2o starts from their casts: the oracle gets the same starting values"""
    from oracle.f32 import Pose
    from oracle.se3 import SE3
    octave = np.where(np.isclose(p.obs_isg, 1.0), 0, 1)
    isg = np.where(octave == 0, np.float32(1.0), np.float32(1.0) / (np.float32(1.2) * np.float32(1.2))).astype(np.float64)
    poses = [SE3.from_pose32(Pose(T.R().astype(np.float32), T.t.astype(np.float32))) for T in p.poses]
    q = oba.BaProblem(poses=poses, pose_fixed=p.pose_fixed, cams=p.cams, X=p.X.astype(np.float32).astype(np.float64), obs_pose=p.obs_pose,
                      obs_point=p.obs_point, obs_uv=p.obs_uv, obs_isg=isg, points_fixed=p.points_fixed)
    return q, octave


def test_shim_bundle_adjustment_and_local_bundle_adjustment(pkg):
    """bundleAdjustment(Map*) and localBundleAdjustment(Map*, id) of host/Optimization.cc on a Map of 4 key frames: poses (written
    back as float), points (float) and the observations taken out of the map against the oracle's flows"""
    p, octave = _as_the_map_holds_it(oba.make_scene(350, 4, seed=13, outliers=20))
    def same_up_to_scale(p7, X, oposes, oX):
        # only key frame 0 is fixed: the scale of a monocular reconstruction is a gauge freedom along which 20 LM iterations
        # drift by a factor that depends on the last bits -- rotations, and translations / points after normalising the scale
        s, so = np.linalg.norm(X - X.mean(0)), np.linalg.norm(oX - oX.mean(0))
        np.testing.assert_allclose(X / s, oX / so, rtol=0, atol=2e-4)
        for k, T in enumerate(oposes):
            np.testing.assert_allclose(p7[k][:4], T.as7()[:4], rtol=0, atol=2e-4)
            np.testing.assert_allclose(p7[k][4:] / s, T.as7()[4:] / so, rtol=0, atol=2e-4)
    oposes, oX, otr = oba.bundle_adjustment(p)
    p7, X, removed, _ = _shim_flow(0, p, octave)
    assert not removed.any()
    same_up_to_scale(p7, X, oposes, oX)
    # local map of key frame 1: every key frame shares points with it, so all four are local (none fixed but key frame 0)
    oposes, oX, oremoved, _ = oba.local_bundle_adjustment(p)
    p7, X, removed, _ = _shim_flow(1, p, octave, curr=1, min_common=0.0)
    assert np.array_equal(removed, oremoved) and removed.sum() >= 20
    same_up_to_scale(p7, X, oposes, oX)


def test_shim_pose_only_optimization(pkg):
    s = oba.make_scene(300, 2, seed=17, outliers=0, point_noise=0.0002)
    m = s.obs_pose == 1
    uv = s.obs_uv[m].copy()
    rng = np.random.default_rng(1)
    bad = rng.choice(len(uv), 35, replace=False)
    uv[bad] += rng.normal(0, 30.0, (35, 2)).astype(np.float32)
    p = oba.BaProblem(poses=[s.poses[1]], pose_fixed=np.array([False]), cams=[s.cams[1]], X=s.X[s.obs_point[m]], obs_pose=np.zeros(m.sum(), int),
                      obs_point=np.arange(m.sum()), obs_uv=uv, obs_isg=s.obs_isg[m], points_fixed=True)
    p, octave = _as_the_map_holds_it(p)
    opose, oinl, ongood = oba.pose_only_optimization(p)
    p7, _, removed, ngood = _shim_flow(2, p, octave)
    assert ngood == ongood and np.array_equal(~removed, oinl) and 250 <= ngood <= 270
    np.testing.assert_allclose(p7[0], opose.as7(), rtol=0, atol=1e-5)
