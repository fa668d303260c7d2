"""The outer weight search of deformationOptimization (oracle = TEST INFRASTRUCTURE ONLY).

  g2oBundleAdjustment.cc:487-530   nlopt::opt(nlopt::LN_NELDERMEAD, 3) over (rep, global, arap), box bounds,
                                   xtol_rel / xtol_abs / maxeval, then the final refinement with the weights found
  nloptOptimization.cc:5-37        outerObjective: clone the Map, arapOptimization(copy, weights), calculatePixelsStandDev,
                                   (ln sigma_C1)^2 + (ln sigma_C2)^2

NLopt is a third-party dependency that is absent from /root/reference and unpinned (Modules/CMakeLists.txt:86-117), so
its Nelder-Mead is restated from the published algorithm (Nelder & Mead 1965 with NLopt's conventions: dimensions with
lb == ub are eliminated, default initial step = a quarter of the box clipped to the distance to the bounds, coefficients
1 / 2 / 0.5 / 0.5, trial points clamped to the box, stop when the simplex's extent in every coordinate is below xtol_abs
or xtol_rel * |best|, or after maxeval evaluations).  PARITY UNPINNED for this function (no reference-held number reaches
it: the historic logs' optimised weights depend on the NLopt and g2o builds); it is the checker of the host shim's
nelder_mead (triangulation-in-deformable-scenes_b200/host/Optimization.cc), written separately from it.
"""
import math

import numpy as np


def nelder_mead(f, x0, lb, ub, xtol_rel, xtol_abs, maxeval):
    """-> (x_best, f_best, log) with log = [(x, f)] of every evaluation in order"""
    x0 = [min(u, max(l, v)) for v, l, u in zip(x0, lb, ub)]
    free = [i for i in range(len(x0)) if ub[i] > lb[i]]
    d = len(free)
    log = []

    def full(y):
        out = list(x0)
        for k, i in enumerate(free):
            out[i] = y[k]
        return out

    def ev(y):
        v = f(full(y))
        log.append((full(y), v))
        return v

    if d == 0:
        v = ev([])
        return list(x0), v, log
    simplex = [[x0[i] for i in free]]
    for k, i in enumerate(free):
        step = (ub[i] - lb[i]) * 0.25
        if ub[i] - x0[i] < step and ub[i] > x0[i]:
            step = (ub[i] - x0[i]) * 0.75
        if x0[i] - lb[i] < step and x0[i] > lb[i]:
            step = (x0[i] - lb[i]) * 0.75
        if not (step > 0) or not math.isfinite(step):
            step = abs(x0[i]) if x0[i] != 0 else 1.0
        v = list(simplex[0])
        v[k] = v[k] + step if v[k] + step <= ub[i] else v[k] - step
        simplex.append(v)
    clamp = lambda y: [min(ub[i], max(lb[i], y[k])) for k, i in enumerate(free)]
    fv = [math.inf] * (d + 1)
    for k in range(d + 1):
        if len(log) >= maxeval:
            break
        fv[k] = ev(simplex[k])
    while len(log) < maxeval:
        order = sorted(range(d + 1), key=lambda k: fv[k])           # stable, like std::sort on distinct values
        lo, hi, nhi = order[0], order[d], order[d - 1]
        conv = True
        for k in range(d):
            col = [v[k] for v in simplex]
            ext = max(col) - min(col)
            if ext > xtol_abs and ext > xtol_rel * abs(simplex[lo][k]):
                conv = False
        if conv:
            break
        c = [0.0] * d
        for j in range(d + 1):
            if j != hi:
                for k in range(d):
                    c[k] += simplex[j][k] / d
        along = lambda t: clamp([c[k] + t * (simplex[hi][k] - c[k]) for k in range(d)])
        xr = along(-1.0)
        fr = ev(xr)
        if fr < fv[lo]:
            xe = along(-2.0)
            fe = ev(xe) if len(log) < maxeval else math.inf
            if fe < fr:
                simplex[hi], fv[hi] = xe, fe
            else:
                simplex[hi], fv[hi] = xr, fr
        elif fr < fv[nhi] or (d == 1 and fr < fv[hi]):
            simplex[hi], fv[hi] = xr, fr
        else:
            xc = along(-0.5 if fr < fv[hi] else 0.5)
            fc = ev(xc) if len(log) < maxeval else math.inf
            if fc < min(fr, fv[hi]):
                simplex[hi], fv[hi] = xc, fc
            else:
                for j in range(d + 1):
                    if j == lo or len(log) >= maxeval:
                        continue
                    simplex[j] = [simplex[lo][k] + 0.5 * (simplex[j][k] - simplex[lo][k]) for k in range(d)]
                    fv[j] = ev(simplex[j])
    best = min(range(d + 1), key=lambda k: fv[k])
    return full(simplex[best]), fv[best], log


def outer_objective(problem, rep, glob, arap, depth_sigma, n_iters):
    """nloptOptimization.cc:5-37 on a copy of the problem's state: refine, then the two pixel sigmas of the float MapPoints"""
    from . import edges, lm, scenes
    w = edges.Weights(rep=rep, arap=arap, depth_sigma=depth_sigma, glob=glob)
    st, _ = lm.optimize(problem, w, n_iters)
    s1 = scenes.pixel_sigma(problem.cam1, problem.T1, st.X1.astype(np.float32), problem.uv1)
    s2 = scenes.pixel_sigma(problem.cam2, problem.T2, st.X2.astype(np.float32), problem.uv2)
    return math.log(s1) ** 2 + math.log(s2) ** 2
