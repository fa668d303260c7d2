"""Classic bundle adjustment paths of the reference (oracle = TEST INFRASTRUCTURE ONLY).

  bundleAdjustment          Modules/Optimization/g2oBundleAdjustment.cc:38-141   poses (key frame id 0 fixed) + points, 20 its
  poseOnlyOptimization      :143-244   one pose, points fixed, 4 rounds of 10 its with inlier re-classification
  localBundleAdjustment     :246-444   local poses + fixed poses + points, 5 its robust, outliers to level 1, 10 its plain
  EdgeSE3ProjectXYZ         g2oTypes.h:150-189, Jacobians g2oTypes.cc:120-140
  EdgeSE3ProjectXYZOnlyPose g2oTypes.h:191-228, Jacobian g2oTypes.cc:165-180

g2o (third party, absent, unpinned) is restated from its published sources as in oracle/lm.py: Levenberg-Marquardt of
OptimizationAlgorithmLevenberg (tau 1e-5 on the largest Hessian diagonal, rho = (chi - chi_trial) / (dx.(lambda dx + b) + 1e-3),
lambda <- lambda max(1/3, 1 - (2 rho - 1)^3) on success, lambda <- lambda nu, nu <- 2 nu on failure, at most 10 trials),
RobustKernelHuber on chi2 = e^T Omega e (weight rho' = delta / sqrt(chi2) beyond delta^2, second-order term dropped),
VertexSE3Expmap::oplus = exp([omega, upsilon]) * T.  BlockSolver_6_3 marginalises the points (Schur complement); the oracle
solves the FULL system (H + lambda I) dx = b directly, which is the same dx -- the CUDA path forms the Schur complement.
PARITY UNPINNED: no executable of the reference calls these functions and no reference-held number reaches them.
"""
from dataclasses import dataclass

import numpy as np

from . import camera
from .se3 import SE3
from .lm import TAU, GOOD_LOWER, GOOD_UPPER, MAX_TRIALS

HUBER_2D = float(np.float32(np.sqrt(5.99)))        # const float thHuber2D = sqrt(5.99)
CHI2_OUTLIER = 5.991


@dataclass
class BaProblem:
    poses: list                 # SE3 (Tcw) per pose
    pose_fixed: np.ndarray      # (K,) bool
    cams: list                  # (model, params float32[8]) per pose
    X: np.ndarray               # (M,3) float64
    obs_pose: np.ndarray        # (O,) int
    obs_point: np.ndarray       # (O,) int
    obs_uv: np.ndarray          # (O,2) float32
    obs_isg: np.ndarray         # (O,) float64   KeyFrame::getInvSigma2(octave)
    points_fixed: bool = False  # poseOnlyOptimization: the points are constants of the edges


def residuals(p, poses, X):
    """computeError of every edge: e = obs - float(project(float(Tcw.map(X)))) -> (e (O,2), Xc (O,3))"""
    e = np.zeros((len(p.obs_pose), 2))
    Xc = np.zeros((len(p.obs_pose), 3))
    for k, T in enumerate(poses):
        m = p.obs_pose == k
        if not m.any():
            continue
        xc = T.map(X[p.obs_point[m]])
        Xc[m] = xc
        proj = camera.project(p.cams[k][0], p.cams[k][1], xc.astype(np.float32))
        e[m] = p.obs_uv[m].astype(np.float64) - proj.astype(np.float64)
    return e, Xc


def edge_chi2(p, poses, X):
    e, Xc = residuals(p, poses, X)
    return p.obs_isg * (e * e).sum(1), Xc[:, 2] > 0.0


def _rho(chi2, robust, delta):
    if not robust:
        return chi2, np.ones_like(chi2)
    d2 = delta * delta
    with np.errstate(all="ignore"):
        s = np.sqrt(chi2)
        return np.where(chi2 <= d2, chi2, 2 * s * delta - d2), np.where(chi2 <= d2, 1.0, delta / s)


def cost(p, poses, X, active, robust, delta=HUBER_2D):
    chi2, _ = edge_chi2(p, poses, X)
    r0, _ = _rho(chi2, robust, delta)
    return float(r0[active].sum())


def linearize(p, poses, X, active, robust, delta=HUBER_2D):
    """dense J-free assembly of H and b over [free poses (6 each) | points (3 each, unless fixed)]"""
    import scipy.sparse as sp
    K, M = len(poses), X.shape[0]
    free = np.nonzero(~p.pose_fixed)[0]
    col_of_pose = -np.ones(K, int)
    col_of_pose[free] = 6 * np.arange(len(free))
    npv = 6 * len(free)
    nv = npv + (0 if p.points_fixed else 3 * M)
    e, Xc = residuals(p, poses, X)
    chi2 = p.obs_isg * (e * e).sum(1)
    _, r1 = _rho(chi2, robust, delta)
    rows, cols, vals = [], [], []
    wts = np.zeros(2 * len(e))
    for o in np.nonzero(active)[0]:
        k, j = p.obs_pose[o], p.obs_point[o]
        Jp = -camera.project_jac(p.cams[k][0], p.cams[k][1], Xc[o].astype(np.float32)[None])[0].astype(np.float64)   # 2x3
        x, y, z = Xc[o]
        dse3 = np.array([[0, z, -y, 1, 0, 0], [-z, 0, x, 0, 1, 0], [y, -x, 0, 0, 0, 1]], np.float64)
        if col_of_pose[k] >= 0:
            Jpose = Jp @ dse3
            for r in range(2):
                for c in range(6):
                    rows.append(2 * o + r); cols.append(col_of_pose[k] + c); vals.append(Jpose[r, c])
        if not p.points_fixed:
            Jx = Jp @ poses[k].R()
            for r in range(2):
                for c in range(3):
                    rows.append(2 * o + r); cols.append(npv + 3 * j + c); vals.append(Jx[r, c])
        wts[2 * o:2 * o + 2] = r1[o] * p.obs_isg[o]
    J = sp.csr_matrix((vals, (rows, cols)), shape=(2 * len(e), nv))
    JW = J.T @ sp.diags(wts)
    H = (JW @ J).tocsc()
    b = -(JW @ e.reshape(-1))
    return H, b, free, npv


def apply_update(p, poses, X, dx, free, npv):
    new = list(poses)
    for q, k in enumerate(free):
        new[k] = poses[k].oplus(dx[6 * q:6 * q + 6])
    Xn = X if p.points_fixed else X + dx[npv:].reshape(-1, 3)
    return new, Xn


def optimize(p, n_iters, active=None, robust=True, delta=HUBER_2D, poses=None, X=None):
    """SparseOptimizer::optimize(n_iters) on the level-0 edges -> (poses, X, trace dict)"""
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve
    poses = list(p.poses if poses is None else poses)
    X = (p.X if X is None else X).astype(np.float64).copy()
    active = np.ones(len(p.obs_pose), bool) if active is None else np.asarray(active, bool)
    tr = dict(chi2=[], lam=[], trials=[], accepted=[])
    lam, ni = 0.0, 2.0
    for it in range(n_iters):
        H, b, free, npv = linearize(p, poses, X, active, robust, delta)
        current = cost(p, poses, X, active, robust, delta)
        if it == 0:
            lam = TAU * float(np.abs(H.diagonal()).max()) if H.shape[0] else 0.0
            ni = 2.0
        tr["chi2"].append(current); tr["lam"].append(lam)
        q, rho, acc = 0, 0.0, False
        while True:
            A = (H + lam * sp.identity(H.shape[0], format="csc")).tocsc()
            try:
                # unobserved points have empty rows: lambda makes them regular; a failed factorisation = g2o's solve() false
                dx = spsolve(A, b) if H.shape[0] else np.zeros(0)
                if not np.all(np.isfinite(dx)):
                    dx = None
            except Exception:
                dx = None
            if dx is None:
                temp, scale = np.finfo(np.float64).max, 1e-3
            else:
                tp, tX = apply_update(p, poses, X, dx, free, npv)
                temp = cost(p, tp, tX, active, robust, delta)
                scale = float(np.dot(dx, lam * dx + b)) + 1e-3
            rho = (current - temp) / scale
            if rho > 0 and np.isfinite(temp):
                alpha = min(1.0 - (2 * rho - 1) ** 3, GOOD_UPPER)
                lam *= max(GOOD_LOWER, alpha)
                ni = 2.0
                current, poses, X, acc = temp, tp, tX, True
            else:
                lam *= ni
                ni *= 2
            q += 1
            if not (rho < 0 and q < MAX_TRIALS):
                break
        tr["trials"].append(q); tr["accepted"].append(acc)
        if q == MAX_TRIALS or rho == 0:
            break
    tr["final_chi2"] = cost(p, poses, X, active, robust, delta)
    return poses, X, tr


def bundle_adjustment(p):
    """g2oBundleAdjustment.cc:38-141: optimize(20) with the Huber kernel on every edge"""
    return optimize(p, 20, robust=True)


def local_bundle_adjustment(p):
    """:246-444 -> (poses, X, removed, traces): `removed` = the observations the reference takes out of the map.
    Edges hold the error of their last evaluation: a level-1 edge is not evaluated by the second run, so its chi2() in
    the final test is the one of the first classification (its isDepthPositive() reads the final estimates)."""
    poses, X, tr1 = optimize(p, 5, robust=True)
    chi2_a, pos_a = edge_chi2(p, poses, X)
    active = ~((chi2_a > CHI2_OUTLIER) | ~pos_a)                    # setLevel(1) on the others; every robust kernel removed
    poses, X, tr2 = optimize(p, 10, active=active, robust=False, poses=poses, X=X)
    chi2_b, pos_b = edge_chi2(p, poses, X)
    stored = np.where(active, chi2_b, chi2_a)
    removed = (stored > CHI2_OUTLIER) | ~pos_b
    return poses, X, removed, (tr1, tr2)


def pose_only_optimization(p):
    """:143-244 with ONE pose (index 0) and fixed points -> (pose, inlier mask, number of inliers).
    Restated literally, including `if(!vInlier[mpIndex]) e->computeError()` indexed by the ROUND number (:209-210)."""
    assert p.points_fixed and len(p.poses) == 1
    O = len(p.obs_pose)
    inlier = np.ones(O, bool)
    level0 = np.ones(O, bool)
    robust = True
    stored_chi2, _ = edge_chi2(p, p.poses, p.X)                     # (_error of an edge that was never evaluated: defined here as the initial one)
    pose = p.poses[0]
    for rnd in range(4):
        poses, _, _ = optimize(p, 10, active=level0, robust=robust, poses=[p.poses[0]])      # setEstimate(fPose): every round restarts
        pose = poses[0]
        fresh, _ = edge_chi2(p, poses, p.X)
        stored_chi2 = np.where(level0, fresh, stored_chi2)          # active edges hold the error of the last iteration's evaluation
        if rnd < O and not inlier[rnd]:
            stored_chi2 = fresh.copy()                              # computeError() on every edge
        inlier = ~(stored_chi2 > CHI2_OUTLIER)
        level0 = inlier.copy()
        if rnd == 2:
            robust = False
    return pose, inlier, int(inlier.sum())


def make_scene(n_points=400, n_poses=3, seed=0, outliers=0, cam=None, px_sigma=1.0, pose_noise=(0.01, 0.01), point_noise=0.003):
    """synthetic key frames looking at a sheet of points: ground truth, noisy key points, perturbed initial estimate"""
    from .scenes import SIM_CAM
    rng = np.random.default_rng(seed)
    cam = (camera.KB8, np.asarray(SIM_CAM if cam is None else cam, np.float32))
    Xw = np.stack([rng.normal(0, 0.05, n_points), rng.normal(0, 0.05, n_points), rng.normal(0.35, 0.02, n_points)], 1)
    true_poses = []
    for k in range(n_poses):
        w = rng.normal(0, 0.05, 3) * (k > 0)
        t = np.array([0.04 * k, 0.01 * k, 0.005 * k]) * (k > 0)
        true_poses.append(SE3().oplus(np.concatenate([w, t])))
    op, oj, uv = [], [], []
    for k, T in enumerate(true_poses):
        xc = T.map(Xw)
        pr = camera.project(cam[0], cam[1], xc.astype(np.float32)).astype(np.float64)
        keep = rng.random(n_points) < (1.0 if k < 2 else 0.7)
        for j in np.nonzero(keep)[0]:
            op.append(k); oj.append(j); uv.append(np.round(pr[j] + rng.normal(0, px_sigma, 2), 1))
    op, oj, uv = np.array(op), np.array(oj), np.array(uv, np.float32)
    if outliers:
        bad = rng.choice(len(op), outliers, replace=False)
        uv[bad] += rng.normal(0, 40.0, (outliers, 2)).astype(np.float32)
    poses = [true_poses[0]] + [T.oplus(np.concatenate([rng.normal(0, pose_noise[0], 3), rng.normal(0, pose_noise[1], 3)])) for T in true_poses[1:]]
    fixed = np.zeros(n_poses, bool)
    fixed[0] = True
    isg = np.where(rng.random(len(op)) < 0.3, 1.0 / 1.44, 1.0)        # octave 0 / 1 with scale factor 1.2
    return BaProblem(poses=poses, pose_fixed=fixed, cams=[cam] * n_poses, X=Xw + rng.normal(0, point_noise, Xw.shape), obs_pose=op, obs_point=oj,
                     obs_uv=uv, obs_isg=isg)
