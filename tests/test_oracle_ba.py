"""oracle/ba.py (classic bundle adjustment): self-checks of the restatement -- analytic Jacobians against central
differences of the residual (the float32 projection limits the step), Schur elimination == full solve, the flows behave."""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

from oracle import ba as oba
from oracle.se3 import SE3


def test_jacobians_against_central_differences():
    p = oba.make_scene(40, 3, seed=1, px_sigma=0.0)
    active = np.ones(len(p.obs_pose), bool)
    H, b, free, npv = oba.linearize(p, p.poses, p.X, active, robust=False)
    e0, _ = oba.residuals(p, p.poses, p.X)
    g = np.zeros(H.shape[0])
    h = 2e-3                                                      # float32 projection: ~1e-5 px noise on a residual
    for c in range(H.shape[0]):
        d = np.zeros(H.shape[0]); d[c] = h
        pp, Xp = oba.apply_update(p, p.poses, p.X, d, free, npv)
        pm, Xm = oba.apply_update(p, p.poses, p.X, -d, free, npv)
        ep, _ = oba.residuals(p, pp, Xp)
        em, _ = oba.residuals(p, pm, Xm)
        J_c = ((ep - em) / (2 * h)).reshape(-1)
        g[c] = -(J_c * (np.repeat(p.obs_isg, 2) * e0.reshape(-1))).sum()
    assert np.abs(g - b).max() <= 2e-3 * np.abs(b).max()


def test_schur_complement_equals_the_full_solve():
    p = oba.make_scene(120, 3, seed=2, outliers=5)
    active = np.ones(len(p.obs_pose), bool)
    H, b, free, npv = oba.linearize(p, p.poses, p.X, active, robust=True)
    lam = 1e-5 * np.abs(H.diagonal()).max()
    A = (H + lam * sp.identity(H.shape[0])).tocsc()
    dx = spsolve(A, b)
    Ad = A.toarray()
    App, Apl, All = Ad[:npv, :npv], Ad[:npv, npv:], Ad[npv:, npv:]
    S = App - Apl @ np.linalg.solve(All, Apl.T)
    dp = np.linalg.solve(S, b[:npv] - Apl @ np.linalg.solve(All, b[npv:]))
    np.testing.assert_allclose(dp, dx[:npv], rtol=1e-8, atol=1e-12)


def test_flows_remove_the_planted_outliers():
    p = oba.make_scene(250, 3, seed=3, outliers=20)
    poses, X, removed, (tr1, tr2) = oba.local_bundle_adjustment(p)
    assert 15 <= removed.sum() <= 45 and tr2["final_chi2"] < tr1["chi2"][0]
    chi2, pos = oba.edge_chi2(p, poses, X)
    assert pos.all() and np.median(chi2[~removed]) < 2.0           # the kept edges fit at the noise level (1 px)
