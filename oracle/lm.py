"""Levenberg-Marquardt driver restating g2o (oracle = test infrastructure).

The reference configures g2o::OptimizationAlgorithmLevenberg over BlockSolverX +
LinearSolverEigen (sparse Cholesky of the full system, no Schur, no marginalised
vertices) at /root/reference/Modules/Optimization/g2oBundleAdjustment.cc:619-628
and runs optimizer.optimize(nOptIterations) at :959-962.  g2o itself is absent
from /root/reference (un-vendored, unpinned).  Restated from upstream
g2o/core/optimization_algorithm_levenberg.cpp:

  per iteration:  currentChi = activeRobustChi2 ; build H, b
                  it 0: lambda = tau * max|diag H| (tau 1e-5), ni = 2
                  do { solve (H + lambda I) x = b ; x_new = x (+) dx ; tempChi
                       rho = (currentChi - tempChi) / (sum_j dx_j (lambda dx_j + b_j) + 1e-3)
                       rho > 0 and finite: lambda *= max(1/3, min(2/3, 1-(2 rho-1)^3)), ni = 2, accept
                       else              : lambda *= ni, ni *= 2, reject }
                  while (rho < 0 and trials < 10)
                  terminate if trials == 10 or rho == 0

The linear solve here is a DIRECT sparse factorisation with the 8 global unknowns
eliminated exactly (bordered / Schur solve) -- the oracle's stand-in for Eigen's
SimplicialLDLT.  `solver="pcg"` runs the same block-Jacobi PCG the CUDA path uses.
"""
import numpy as np
from dataclasses import dataclass, field
from .edges import linearize, total_cost, apply_update, state_of

TAU = 1e-5
GOOD_LOWER = 1.0 / 3.0
GOOD_UPPER = 2.0 / 3.0
MAX_TRIALS = 10


@dataclass
class Trace:
    chi2: list = field(default_factory=list)        # cost at the start of every iteration (+ final appended)
    lam: list = field(default_factory=list)         # lambda at the start of every iteration
    trials: list = field(default_factory=list)      # number of lambda trials in the iteration
    accepted: list = field(default_factory=list)    # whether the iteration ended with an accepted step
    trial_log: list = field(default_factory=list)   # (iter, lambda, tempChi, rho, info)
    final_chi2: float = None
    stop: str = "max_iterations"


def solve_direct(H, b, lam):
    """(H + lam I) x = b with the 8-wide border eliminated exactly."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = H.shape[0]
    H = H.tocsc()
    App = (H[8:, 8:] + lam * sp.identity(n - 8, format="csc")).tocsc()
    B = H[8:, :8].toarray()
    C = H[:8, :8].toarray() + lam * np.eye(8)
    try:
        lu = spla.splu(App, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                       options=dict(SymmetricMode=True))
    except RuntimeError:
        return None
    Y = lu.solve(np.concatenate([b[8:, None], B], axis=1))
    yb, YB = Y[:, 0], Y[:, 1:]
    S = C - B.T @ YB
    rhs = b[:8] - B.T @ yb
    try:
        xg = np.linalg.solve(S, rhs)
    except np.linalg.LinAlgError:
        return None
    xp = yb - YB @ xg
    x = np.concatenate([xg, xp])
    return x if np.all(np.isfinite(x)) else None


def block_jacobi_inverse(H, lam, block=6):
    """Inverse of the block diagonal of H + lam I: one 8x8 global block and one
    `block`x`block` (6: a correspondence's X1|X2, 3: per point) block per correspondence."""
    import scipy.sparse as sp
    n = H.shape[0]
    Hc = H.tocsr()
    G = np.linalg.inv(Hc[:8, :8].toarray() + lam * np.eye(8))
    nb = (n - 8) // block
    D = np.zeros((nb, block, block))
    Hp = Hc[8:, 8:].tocoo()
    same = (Hp.row // block) == (Hp.col // block)
    np.add.at(D, (Hp.row[same] // block, Hp.row[same] % block, Hp.col[same] % block), Hp.data[same])
    D += lam * np.eye(block)[None]
    Dinv = np.linalg.inv(D)

    def apply(r):
        z = np.empty_like(r)
        z[:8] = G @ r[:8]
        z[8:] = np.einsum("nij,nj->ni", Dinv, r[8:].reshape(nb, block)).reshape(-1)
        return z
    return apply


def solve_pcg(H, b, lam, rtol=1e-10, max_iter=2000, block=6, info=None):
    """Preconditioned CG on (H + lam I) x = b, x0 = 0, block-Jacobi preconditioner,
    stop when sqrt(r.z / r0.z0) <= rtol.  Same recurrence as the CUDA solver."""
    M = block_jacobi_inverse(H, lam, block)
    Hc = H.tocsr()
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = float(r @ z)
    rz0 = rz
    it = 0
    if rz0 > 0:
        for it in range(1, max_iter + 1):
            Ap = Hc @ p + lam * p
            alpha = rz / float(p @ Ap)
            x += alpha * p
            r -= alpha * Ap
            z = M(r)
            rz_new = float(r @ z)
            if not np.isfinite(rz_new):
                return None
            if np.sqrt(abs(rz_new) / rz0) <= rtol:
                rz = rz_new
                break
            p = z + (rz_new / rz) * p
            rz = rz_new
    if info is not None:
        info["iters"] = it
        info["rel"] = float(np.sqrt(abs(rz) / rz0)) if rz0 > 0 else 0.0
    return x if np.all(np.isfinite(x)) else None


def optimize(p, w, n_iters, fd=False, solver="direct", pcg_rtol=1e-10, pcg_max_iter=2000,
             pcg_block=6, verbose=False):
    """g2o SparseOptimizer::optimize(n_iters) on problem p with weights w.
    Returns (final State, Trace)."""
    import scipy.sparse as sp
    st = state_of(p)
    tr = Trace()
    lam = 0.0
    ni = 2.0
    for it in range(n_iters):
        J, wt, e, chi_lin = linearize(p, w, st, fd=fd)
        current = total_cost(p, w, st)
        JW = J.T.multiply(wt[None, :]).tocsr() if False else (J.T @ sp.diags(wt))
        H = (JW @ J).tocsr()
        b = -(JW @ e)
        if it == 0:
            lam = TAU * float(np.abs(H.diagonal()).max())
            ni = 2.0
        tr.chi2.append(current)
        tr.lam.append(lam)
        rho = 0.0
        q = 0
        acc = False
        while True:
            info = {}
            if solver == "direct":
                dx = solve_direct(H, b, lam)
            else:
                dx = solve_pcg(H, b, lam, pcg_rtol, pcg_max_iter, pcg_block, info)
            if dx is None:
                temp = np.finfo(np.float64).max
                scale = 1e-3
            else:
                trial = apply_update(st, dx)
                temp = total_cost(p, w, trial)
                scale = float(np.dot(dx, lam * dx + b)) + 1e-3
            rho = (current - temp) / scale
            tr.trial_log.append((it, lam, temp, rho, info))
            if verbose:
                print(f"  it {it} trial {q} lam {lam:.6e} chi {current:.9e} -> {temp:.9e} rho {rho:.4f} {info}")
            if rho > 0 and np.isfinite(temp):
                alpha = 1.0 - (2 * rho - 1) ** 3
                alpha = min(alpha, GOOD_UPPER)
                lam *= max(GOOD_LOWER, alpha)
                ni = 2.0
                current = temp
                st = trial
                acc = True
            else:
                lam *= ni
                ni *= 2
            q += 1
            if not (rho < 0 and q < MAX_TRIALS):
                break
        tr.trials.append(q)
        tr.accepted.append(acc)
        if q == MAX_TRIALS or rho == 0:
            tr.stop = "terminate"
            break
    tr.final_chi2 = total_cost(p, w, st)
    return st, tr


def write_back(p, st):
    """g2oBundleAdjustment.cc:967-1007: points cast to float, update = sum |p_old - p_new| (float norm)."""
    new1 = st.X1.astype(np.float32)
    new2 = st.X2.astype(np.float32)
    old1 = p.X1.astype(np.float32)
    old2 = p.X2.astype(np.float32)
    upd = np.sqrt(((old1 - new1) ** 2).sum(1, dtype=np.float32)).astype(np.float64).sum() + \
        np.sqrt(((old2 - new2) ** 2).sum(1, dtype=np.float32)).astype(np.float64).sum()
    return new1, new2, float(upd)
