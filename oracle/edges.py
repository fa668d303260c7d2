"""Residuals, costs and Jacobians of the three edge types (oracle = test infrastructure).

Follows /root/reference/Modules/Optimization/g2oTypes.h:
  EdgeSE3ProjectXYZPerKeyFrameOnlyPoints  :267-298, Jacobian g2oTypes.cc:270-283
  EdgeARAP                                :300-349 (no analytic Jacobian -> g2o central differences)
  EdgeDepthCorrection                     :390-421 (idem)
and their set-up in g2oBundleAdjustment.cc:777-953 (information matrices, Huber
delta sqrt(100.991) on the reprojection edges only).

Unknown vector layout used by the oracle (the reference's first-touch vertex-id
order only changes the elimination order of its sparse Cholesky):
  x = [ T_g: omega(3) upsilon(3) | s1 | s2 | X1_0 X2_0 | X1_1 X2_1 | ... ]   size 8 + 6N
"""
import numpy as np
from dataclasses import dataclass, field, replace
from . import camera
from .se3 import SE3, skew

HUBER_DELTA = float(np.float32(np.sqrt(100.991)))   # const float deltaMono = sqrt(100.991)  (:631)
FD_DELTA = 1e-9                                     # g2o numeric Jacobian step (base_*_edge.hpp)


@dataclass
class Weights:
    rep: float = 1.0
    arap: float = 1.0
    depth_sigma: float = 1.0        # metres; information = 1/sigma^2   (:822-825)
    glob: float = 0.0               # accepted, never used by the reference (arapOptimization body)
    alpha: float = 1.0              # stored on the edge, unused (g2oTypes.h:346-347)
    beta: float = 1.0


@dataclass
class Problem:
    cam1: tuple
    cam2: tuple
    T1: object                      # f32.Pose (Tcw of KF1)
    T2: object
    uv1: np.ndarray                 # (N,2) float32
    uv2: np.ndarray
    inv_sigma2_1: np.ndarray        # (N,) float64   KeyFrame::getInvSigma2(octave)
    inv_sigma2_2: np.ndarray
    d1: np.ndarray                  # (N,) float64   depth measurements (metres, unscaled)
    d2: np.ndarray
    graph: object                   # graph.Graph over the N correspondences
    X1: np.ndarray                  # (N,3) float64
    X2: np.ndarray
    Tg: SE3 = field(default_factory=SE3)
    s1: float = 1.0
    s2: float = 1.0
    R: np.ndarray = None            # (N,3,3) per-vertex rotations (computeR)

    @property
    def n(self):
        return self.X1.shape[0]


@dataclass
class State:
    X1: np.ndarray
    X2: np.ndarray
    Tg: SE3
    s1: float
    s2: float

    def copy(self):
        return State(self.X1.copy(), self.X2.copy(), SE3(self.Tg.q, self.Tg.t), self.s1, self.s2)


def state_of(p):
    return State(p.X1.astype(np.float64).copy(), p.X2.astype(np.float64).copy(),
                 SE3(p.Tg.q, p.Tg.t), float(p.s1), float(p.s2))


def apply_update(st, dx):
    """g2o oplus of every vertex: SE3 exp-left-multiply, scalars and points additive."""
    n = st.X1.shape[0]
    d = dx[8:].reshape(n, 6)
    return State(st.X1 + d[:, :3], st.X2 + d[:, 3:], st.Tg.oplus(dx[:6]), st.s1 + dx[6], st.s2 + dx[7])


# ---------------------------------------------------------------- reprojection
def reproj_residual(cam, T, X, uv):
    """computeError (g2oTypes.h:277-291): e = obs - float(project(float(Tcw.map(X))))."""
    Tq = SE3.from_pose32(T)
    Xc = Tq.map(X)
    proj = camera.project(cam[0], cam[1], Xc.astype(np.float32))
    return uv.astype(np.float64) - proj.astype(np.float64), Xc, Tq


def reproj_jacobian(cam, Tq, Xc):
    """linearizeOplus (g2oTypes.cc:270-283): J = -projectJac(float(Xc)) * R_cw."""
    Jp = camera.project_jac(cam[0], cam[1], Xc.astype(np.float32)).astype(np.float64)
    return -Jp @ Tq.R()


def huber(chi2, delta=HUBER_DELTA):
    """g2o::RobustKernelHuber::robustify -> rho[0], rho[1]."""
    d2 = delta * delta
    with np.errstate(all="ignore"):
        s = np.sqrt(chi2)
        rho0 = np.where(chi2 <= d2, chi2, 2 * s * delta - d2)
        rho1 = np.where(chi2 <= d2, 1.0, delta / s)
    return rho0, rho1


# ---------------------------------------------------------------- depth
def depth_residual(T, X, d, s):
    """computeError (g2oTypes.h:400-416): e = (d/s - z_c)^2, x500 if s <= 0."""
    Tq = SE3.from_pose32(T)
    zc = Tq.map(X)[:, 2]
    r = d / s - zc
    e = r * r
    if s <= 0.0:
        e = e * 500
    return e, r, Tq


def depth_jacobian(Tq, r, d, s):
    k = 500.0 if s <= 0.0 else 1.0
    JX = (2.0 * k * r)[:, None] * (-Tq.R()[2, :])[None, :]
    Js = 2.0 * k * r * (-d / (s * s))
    return JX, Js


# ---------------------------------------------------------------- ARAP
def arap_terms(p, st):
    """Per directed edge (i -> j) of the CSR graph.  computeError (g2oTypes.h:310-339):
    e = w * (|(d2 - Ri d1)/area|^2 + |(-d2 + Rj d1)/area|^2) + |Rg(X2i+X2j) - 2t - (X1i+X1j)|^2."""
    g = p.graph
    i = g.rows()
    j = g.col
    Rg = st.Tg.R()
    t = st.Tg.t
    d1 = st.X1[i] - st.X1[j]
    d2 = st.X2[i] - st.X2[j]
    a = (d2 - np.einsum("eij,ej->ei", p.R[i], d1)) / g.area
    b = (d2 - np.einsum("eij,ej->ei", p.R[j], d1)) / g.area
    S2 = st.X2[i] + st.X2[j]
    S1 = st.X1[i] + st.X1[j]
    qt = S2 @ Rg.T - 2.0 * t
    gg = qt - S1
    e = g.w * (np.einsum("ei,ei->e", a, a) + np.einsum("ei,ei->e", b, b)) + np.einsum("ei,ei->e", gg, gg)
    return e, (i, j, a, b, gg, qt, Rg)


def arap_jacobian(p, terms):
    """Analytic gradient of the scalar ARAP energy w.r.t. (T_g[6], X1i, X2i, X1j, X2j)."""
    g = p.graph
    i, j, a, b, gg, qt, Rg = terms
    c2 = (2.0 * g.w / g.area)[:, None]
    u = c2 * (a + b)                                                     # d e / d X2i (ARAP part)
    m = c2 * (np.einsum("eji,ej->ei", p.R[i], a) + np.einsum("eji,ej->ei", p.R[j], b))
    v2 = 2.0 * gg @ Rg                                                   # 2 Rg^T g
    J_i1 = -m - 2.0 * gg
    J_i2 = u + v2
    J_j1 = m - 2.0 * gg
    J_j2 = -u + v2
    J_T = np.concatenate([2.0 * np.cross(qt, gg), -4.0 * gg], axis=1)    # [d/d omega, d/d upsilon]
    return J_T, J_i1, J_i2, J_j1, J_j2


# ---------------------------------------------------------------- total cost
def total_cost(p, w, st, parts=False):
    """activeRobustChi2: Huber rho[0] on the reprojection edges, plain chi2 elsewhere."""
    e1, _, _ = reproj_residual(p.cam1, p.T1, st.X1, p.uv1)
    e2, _, _ = reproj_residual(p.cam2, p.T2, st.X2, p.uv2)
    c1 = (p.inv_sigma2_1 * w.rep) * np.einsum("ni,ni->n", e1, e1)
    c2 = (p.inv_sigma2_2 * w.rep) * np.einsum("ni,ni->n", e2, e2)
    rep = huber(c1)[0].sum() + huber(c2)[0].sum()
    od = 1.0 / (float(np.float32(w.depth_sigma)) ** 2)
    ed1, _, _ = depth_residual(p.T1, st.X1, p.d1, st.s1)
    ed2, _, _ = depth_residual(p.T2, st.X2, p.d2, st.s2)
    dep = od * (np.dot(ed1, ed1) + np.dot(ed2, ed2))
    ea, _ = arap_terms(p, st)
    oa = w.arap * float(p.graph.n_triangles) ** 2
    ar = oa * np.dot(ea, ea)
    if parts:
        return rep + dep + ar, (rep, dep, ar)
    return rep + dep + ar


# ---------------------------------------------------------------- linearisation
def linearize(p, w, st, fd=False):
    """Returns sparse Jacobian J (rows = scalar residuals), per-row weight (information x
    robust rho'), residual vector e, and the current robust chi2.
    H = J^T diag(wt) J, b = -J^T diag(wt) e  == g2o constructQuadraticForm for all edges."""
    import scipy.sparse as sp
    n = p.n
    nu = 8 + 6 * n
    idx = np.arange(n)
    c1 = 8 + 6 * idx
    c2 = c1 + 3
    rows, cols, vals, res, wts = [], [], [], [], []
    r0 = 0
    chi = 0.0
    od = 1.0 / (float(np.float32(w.depth_sigma)) ** 2)
    oa = w.arap * float(p.graph.n_triangles) ** 2

    def add(rr, cc, vv):
        rows.append(np.asarray(rr).ravel())
        cols.append(np.asarray(cc).ravel())
        vals.append(np.asarray(vv).ravel())

    # reprojection (analytic in the reference too)
    for cam, T, X, uv, isg, cbase in ((p.cam1, p.T1, st.X1, p.uv1, p.inv_sigma2_1, c1),
                                      (p.cam2, p.T2, st.X2, p.uv2, p.inv_sigma2_2, c2)):
        e, Xc, Tq = reproj_residual(cam, T, X, uv)
        J = reproj_jacobian(cam, Tq, Xc)                                  # (n,2,3)
        om = isg * w.rep
        c = om * np.einsum("ni,ni->n", e, e)
        rho0, rho1 = huber(c)
        chi += rho0.sum()
        for k in range(2):
            rr = r0 + 2 * idx + k
            add(np.repeat(rr, 3), (cbase[:, None] + np.arange(3)[None, :]), J[:, k, :])
        res.append(e.reshape(-1))
        wts.append(np.repeat(om * rho1, 2))
        r0 += 2 * n

    # depth
    for T, X, d, s, scol, cbase, which in ((p.T1, st.X1, p.d1, st.s1, 6, c1, 1),
                                           (p.T2, st.X2, p.d2, st.s2, 7, c2, 2)):
        e, r, Tq = depth_residual(T, X, d, s)
        if fd:
            JX = np.zeros((n, 3))
            for k in range(3):
                dX = np.zeros(3)
                dX[k] = FD_DELTA
                ep = depth_residual(T, X + dX, d, s)[0]
                em = depth_residual(T, X - dX, d, s)[0]
                JX[:, k] = (ep - em) / (2 * FD_DELTA)
            Js = (depth_residual(T, X, d, s + FD_DELTA)[0] - depth_residual(T, X, d, s - FD_DELTA)[0]) / (2 * FD_DELTA)
        else:
            JX, Js = depth_jacobian(Tq, r, d, s)
        chi += od * np.dot(e, e)
        rr = r0 + idx
        add(np.repeat(rr, 3), (cbase[:, None] + np.arange(3)[None, :]), JX)
        add(rr, np.full(n, scol), Js)
        res.append(e)
        wts.append(np.full(n, od))
        r0 += n

    # ARAP
    e, terms = arap_terms(p, st)
    ne = len(e)
    if ne:
        i, j = terms[0], terms[1]
        if fd:
            JT, Ji1, Ji2, Jj1, Jj2 = _arap_fd(p, st)
        else:
            JT, Ji1, Ji2, Jj1, Jj2 = arap_jacobian(p, terms)
        chi += oa * np.dot(e, e)
        rr = r0 + np.arange(ne)
        add(np.repeat(rr, 6), np.tile(np.arange(6), ne), JT)
        for Jb, cb in ((Ji1, c1[i]), (Ji2, c2[i]), (Jj1, c1[j]), (Jj2, c2[j])):
            add(np.repeat(rr, 3), cb[:, None] + np.arange(3)[None, :], Jb)
        res.append(e)
        wts.append(np.full(ne, oa))
        r0 += ne

    J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, nu))
    return J, np.concatenate(wts), np.concatenate(res), chi


def _arap_fd(p, st):
    """g2o BaseMultiEdge::linearizeOplus: central differences, delta 1e-9, 18 columns/edge."""
    ne = p.graph.n_edges
    i = p.graph.rows()
    j = p.graph.col
    JT = np.zeros((ne, 6))
    for k in range(6):
        d = np.zeros(6)
        d[k] = FD_DELTA
        sp_ = replace(st, Tg=st.Tg.oplus(d))
        sm_ = replace(st, Tg=st.Tg.oplus(-d))
        JT[:, k] = (arap_terms(p, sp_)[0] - arap_terms(p, sm_)[0]) / (2 * FD_DELTA)
    out = []
    # perturbing vertex role (i1, i2, j1, j2) of every edge at once is NOT possible globally
    # (a point is `i` of some edges and `j` of others), so evaluate per role with edge-local copies.
    g = p.graph
    Rg = st.Tg.R()
    t = st.Tg.t

    def energy(X1i, X2i, X1j, X2j):
        d1 = X1i - X1j
        d2 = X2i - X2j
        a = (d2 - np.einsum("eij,ej->ei", p.R[i], d1)) / g.area
        b = (d2 - np.einsum("eij,ej->ei", p.R[j], d1)) / g.area
        gg = (X2i + X2j) @ Rg.T - 2.0 * t - (X1i + X1j)
        return g.w * (np.einsum("ei,ei->e", a, a) + np.einsum("ei,ei->e", b, b)) + np.einsum("ei,ei->e", gg, gg)

    base = [st.X1[i], st.X2[i], st.X1[j], st.X2[j]]
    for role in range(4):
        Jr = np.zeros((ne, 3))
        for k in range(3):
            dX = np.zeros(3)
            dX[k] = FD_DELTA
            ap = list(base)
            am = list(base)
            ap[role] = base[role] + dX
            am[role] = base[role] - dX
            Jr[:, k] = (energy(*ap) - energy(*am)) / (2 * FD_DELTA)
        out.append(Jr)
    return (JT, *out)
