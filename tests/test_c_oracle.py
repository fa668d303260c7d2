"""The plain-C oracle (oracle/c/dsc_oracle.c) against the numpy oracle: two independently written restatements of
the same reference code must agree (CPU only)."""
import numpy as np
import pytest

from oracle import cport, edges, graph, lm, scenes


@pytest.fixture(scope="module")
def small():
    sc = scenes.sheet_scene(400, seed=21)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    w = edges.Weights(rep=1.0, arap=30.0, depth_sigma=0.003)
    return p, w


def test_c_oracle_builds_and_loads():
    cport.build()
    assert cport.threads() >= 1


def test_cost_and_rotations_match_numpy(small):
    p, w = small
    cp = cport.CProblem(p, rotations=p.R)
    v, parts = cport.cost(cp, w)
    ref, rparts = edges.total_cost(p, w, edges.state_of(p), parts=True)
    assert v == pytest.approx(ref, rel=1e-12)
    np.testing.assert_allclose(parts, rparts, rtol=1e-11)
    cq = cport.CProblem(p)
    bad = cport.compute_rotations(cq)
    R = graph.compute_rotations(p.graph, p.X1, p.X2)
    assert bad == 0
    np.testing.assert_allclose(cq.R, R, atol=1e-9)
    det = np.linalg.det(cq.R)
    np.testing.assert_allclose(det, 1.0, atol=1e-9)


def test_linearisation_matches_numpy(small):
    import scipy.sparse as sp
    p, w = small
    st = edges.state_of(p)
    J, wt, e, chi = edges.linearize(p, w, st)
    JW = J.T @ sp.diags(wt)
    H = (JW @ J).tocsr()
    b = -(JW @ e)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(H.shape[0])
    lam = 1e-5 * np.abs(H.diagonal()).max()
    cp = cport.CProblem(p, rotations=p.R)
    cb, chd, cy, cchi = cport.debug_linearize(cp, w, lam, x)
    assert cchi == pytest.approx(chi, rel=1e-12)
    np.testing.assert_allclose(cb, b, rtol=1e-9, atol=1e-11 * np.abs(b).max())
    np.testing.assert_allclose(chd, H.diagonal(), rtol=1e-9, atol=1e-13 * np.abs(H.diagonal()).max())
    yo = H @ x + lam * x
    np.testing.assert_allclose(cy, yo, rtol=1e-9, atol=1e-11 * np.abs(yo).max())


def test_numeric_jacobians_are_the_reference_mode(small):
    """fd = 1 differentiates depth and ARAP edges as g2o does (central differences, 1e-9): close to analytic."""
    p, w = small
    cp = cport.CProblem(p, rotations=p.R)
    x = np.zeros(8 + 6 * p.n)
    ba, hda, _, _ = cport.debug_linearize(cp, w, 0.0, x, fd=False)
    bn, hdn, _, _ = cport.debug_linearize(cp, w, 0.0, x, fd=True)
    assert np.abs(bn - ba).max() <= 2e-3 * np.abs(ba).max()
    assert np.abs(hdn - hda).max() <= 2e-3 * np.abs(hda).max()


def test_lm_trace_matches_the_direct_solve_oracle(small):
    p, w = small
    ost, otr = lm.optimize(p, w, 5)
    cp = cport.CProblem(p, rotations=p.R)
    tr = cport.optimize(cp, w, 5, pcg_rtol=1e-13)
    assert len(tr["chi2"]) == len(otr.chi2)
    for a, b2 in zip(tr["chi2"], otr.chi2):
        assert a == pytest.approx(b2, rel=1e-6)
    assert tr["trials"] == otr.trials
    for a, b2 in zip(tr["lam"], otr.lam):
        assert a == pytest.approx(b2, rel=1e-4)
    assert tr["final_chi2"] == pytest.approx(otr.final_chi2, rel=1e-6)
    scale = np.abs(np.concatenate([ost.X1, ost.X2])).max()
    assert np.abs(cp.X1 - ost.X1).max() <= 1e-6 * scale
    assert np.abs(cp.X2 - ost.X2).max() <= 1e-6 * scale
    s1, s2 = cp.scales()
    assert s1 == pytest.approx(ost.s1, rel=1e-6) and s2 == pytest.approx(ost.s2, rel=1e-6)
    np.testing.assert_allclose(cp.Tg7(), ost.Tg.as7(), atol=1e-8)


def test_threads_do_not_change_the_trace(small):
    p, w = small
    a = cport.CProblem(p, rotations=p.R)
    b = cport.CProblem(p, rotations=p.R)
    ta = cport.optimize(a, w, 3, threads=1)
    tb = cport.optimize(b, w, 3, threads=4)
    assert ta["trials"] == tb["trials"]
    for x, y in zip(ta["chi2"], tb["chi2"]):
        assert x == pytest.approx(y, rel=1e-9)


def test_huber_branch_matches_numpy():
    """Gross outliers switch the reprojection edges to the linear branch of the Huber kernel in both restatements."""
    sc = scenes.sheet_scene(300, seed=23)
    p, keep = scenes.problem_from_scene(sc, "knn", 8)
    rng = np.random.default_rng(1)
    bad = rng.choice(p.n, p.n // 10, replace=False)
    p.uv1 = p.uv1.copy()
    p.uv1[bad] += np.float32(45.0)
    w = edges.Weights(rep=1.0, arap=20.0, depth_sigma=0.003)
    cp = cport.CProblem(p, rotations=p.R)
    v, parts = cport.cost(cp, w)
    ref, rparts = edges.total_cost(p, w, edges.state_of(p), parts=True)
    assert v == pytest.approx(ref, rel=1e-12)
    ost, otr = lm.optimize(p, w, 4)
    tr = cport.optimize(cp, w, 4, pcg_rtol=1e-13)
    assert tr["trials"] == otr.trials
    for a, b2 in zip(tr["chi2"], otr.chi2):
        assert a == pytest.approx(b2, rel=1e-6)


def test_numeric_jacobian_lm_trace_is_within_the_parity_bar(small):
    """The reference differentiates the depth and ARAP edges numerically (g2o central differences, 1e-9); the CUDA path
    and both oracles use the analytic gradients.  Running the SAME LM with the numeric Jacobians (fd = 1, the
    reference-faithful mode of the C oracle) changes the per-iteration cost by ~1e-7 relative: far inside the 1e-5 bar."""
    p, w = small
    a = cport.CProblem(p, rotations=p.R)
    b = cport.CProblem(p, rotations=p.R)
    ta = cport.optimize(a, w, 6, fd=False, pcg_rtol=1e-13)
    tb = cport.optimize(b, w, 6, fd=True, pcg_rtol=1e-13)
    assert ta["trials"] == tb["trials"]
    for x, y in zip(ta["chi2"], tb["chi2"]):
        assert x == pytest.approx(y, rel=1e-6)
    assert ta["final_chi2"] == pytest.approx(tb["final_chi2"], rel=1e-6)
    assert np.abs(a.X1 - b.X1).max() <= 1e-6 * np.abs(a.X1).max()
