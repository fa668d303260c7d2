"""Two-view triangulation of matched keypoints, float32 (oracle = test infrastructure).

Follows /root/reference/Modules/Utils/Geometry.cc:30-230 (triangulators and
dispatcher), Modules/Mapping/Mapping.cc:280-364 (simulation call site + parallax
gate), Modules/Mapping/MonocularMapInitializer.cc:303-368 (real-image call site
+ gates), Modules/Map/KeyFrame.cc:131-153 (initial depth scale, simulation).

Divergences from the reference, all where the reference's result is undefined:
  * ORBSLAM/DLT (Geometry.cc:155-186) writes its result to locals and never to
    x3D_1/x3D_2.  Here both outputs are the de-homogenised null vector.
  * DepthMeasurement uses CameraModel::unproject(pt, z) which returns
    uninitialised memory (CameraModel.h:128-135).  Here: ray scaled so its
    camera-z equals the measured depth.
"""
import numpy as np
from .f32 import f32, F, dot3, cross3, norm3, normalize3, matvec, Pose, emu
from . import camera

METHODS = {"Classic": 0, "NRSLAM": 1, "ORBSLAM": 2, "DepthMeasurement": 3}
LOCATIONS = {"InRays": 0, "TwoPoints": 1, "FarPoints": 2}
GATE_NONE, GATE_SIM, GATE_REAL = 0, 1, 2


def method_id(name):
    # Geometry.cc:220-228: anything that is not Classic/ORBSLAM/DepthMeasurement is NRSLAM
    return METHODS.get(name, 1)


def location_id(name):
    return LOCATIONS.get(name, 0)


def cos_ray_parallax(a, b):
    """Geometry.cc:30-32."""
    return dot3(a, b) / (norm3(a) * norm3(b))


def _second_right_singular_vector_2x3(A):
    """V.col(1) of the 2x3 JacobiSVD (Geometry.cc:75-77); evaluated in double, rounded."""
    A64 = A.astype(np.float64)
    _, _, Vt = np.linalg.svd(A64, full_matrices=True)
    return Vt[..., 1, :].astype(np.float32)


def triangulate_classic(xn1, xn2, T1w, T2w, location):
    """Geometry.cc:62-101."""
    T21 = T2w.compose(T1w.inverse())
    m0 = matvec(T21.R, xn1)
    m1 = xn2
    t = normalize3(T21.t)
    M0 = normalize3(m0)
    M1 = normalize3(m1)
    # A = M^T (I - t t^T): row k = M_k - (M_k . t) t
    a0 = M0 - dot3(M0, t)[..., None] * t
    a1 = M1 - dot3(M1, t)[..., None] * t
    A = np.stack([a0, a1], axis=-2)
    n = _second_right_singular_vector_2x3(A)
    m0_ = m0 - dot3(m0, n)[..., None] * n
    m1_ = m1 - dot3(m1, n)[..., None] * n
    z = cross3(m1_, m0_)
    tt = np.broadcast_to(T21.t, m0.shape)
    with np.errstate(all="ignore"):
        lambda0 = dot3(z, cross3(tt, m1_)) / dot3(z, z)
        lambda1 = dot3(z, cross3(tt, m0_)) / dot3(z, z)
    if location == LOCATIONS["TwoPoints"]:
        p1 = tt + lambda0[..., None] * m0_
        p2 = p1.copy()
    else:
        p1 = tt + lambda0[..., None] * m0
        p2 = lambda1[..., None] * m1
    Tw2 = T2w.inverse()
    return Tw2.apply(p1), Tw2.apply(p2)


def triangulate_nrslam(xn1, xn2, T1w, T2w, location):
    """Geometry.cc:103-153 (inverse-depth weighted midpoint, Lee & Civera 2019)."""
    f0_hat = normalize3(xn1)
    f1_hat = normalize3(xn2)
    T21 = T2w.compose(T1w.inverse())
    t = np.broadcast_to(T21.t, f0_hat.shape)
    Rf0 = matvec(T21.R, f0_hat)
    p = cross3(Rf0, f1_hat)
    q = cross3(Rf0, t)
    r = cross3(f1_hat, t)
    np_, nq, nr = norm3(p), norm3(q), norm3(r)
    with np.errstate(all="ignore"):
        lambda0 = nr / np_
        lambda1 = nq / np_
        point0 = lambda0[..., None] * Rf0
        point1 = lambda1[..., None] * f1_hat
        x1 = (nq / (nq + nr))[..., None] * (t + lambda0[..., None] * (Rf0 + f1_hat))
    if location == LOCATIONS["TwoPoints"]:
        p1, p2 = x1, x1.copy()
    elif location == LOCATIONS["FarPoints"]:
        point0 = t + point0
        p1 = point0 + (point0 - x1)
        p2 = point1 + (point1 - x1)
    else:
        p1 = t + point0
        p2 = point1
    Tw2 = T2w.inverse()
    return Tw2.apply(p1), Tw2.apply(p2)


def triangulate_dlt(xn1, xn2, T1w, T2w, location):
    """Geometry.cc:155-186 (intended behaviour, see module docstring)."""
    def rows(T):
        M = T.as34()
        return M[0], M[1], M[2]
    r10, r11, r12 = rows(T1w)
    r20, r21, r22 = rows(T2w)
    A = np.stack([xn1[..., 0:1] * r12 - r10,
                  xn1[..., 1:2] * r12 - r11,
                  xn2[..., 0:1] * r22 - r20,
                  xn2[..., 1:2] * r22 - r21], axis=-2).astype(np.float32)
    _, _, Vt = np.linalg.svd(A.astype(np.float64), full_matrices=True)
    x = Vt[..., 3, :].astype(np.float32)
    with np.errstate(all="ignore"):
        X = np.where((x[..., 3:4] != 0), x[..., :3] / x[..., 3:4], F(0)).astype(np.float32)
    return X, X.copy()


def triangulate_depth(x1, x2, T1w, T2w, location):
    """Geometry.cc:189-214; x1/x2 are camera-frame points at the measured depth."""
    T21 = T2w.compose(T1w.inverse())
    point0 = T21.apply(x1)
    point1 = x2
    x1m = (point0 + point1) / F(2.0)
    if location == LOCATIONS["TwoPoints"]:
        p1, p2 = x1m, x1m.copy()
    elif location == LOCATIONS["FarPoints"]:
        p1 = point0 + (point0 - x1m)
        p2 = point1 + (point1 - x1m)
    else:
        p1, p2 = point0, point1
    Tw2 = T2w.inverse()
    return Tw2.apply(p1), Tw2.apply(p2)


def triangulate_pairs(uv1, uv2, cam1, cam2, T1w, T2w, method="NRSLAM", location="FarPoints",
                      gate=GATE_SIM, min_cos=0.9998, depth_limit=np.inf, check_reproj=False,
                      d1=None, d2=None):
    """Batched restatement of Mapping::triangulateSimulatedMapPoints (Mapping.cc:294-343)
    and MonocularMapInitializer::reconstructPoints (:303-368).

    cam = (model, params[8]).  Returns X1, X2 (N,3) float32, valid (N,) bool, cosp (N,) float32.
    """
    m = method_id(method) if isinstance(method, str) else int(method)
    loc = location_id(location) if isinstance(location, str) else int(location)
    uv1, uv2 = f32(uv1), f32(uv2)
    ray1 = camera.unproject(cam1[0], cam1[1], uv1)
    ray2 = camera.unproject(cam2[0], cam2[1], uv2)
    xn1 = normalize3(ray1)
    xn2 = normalize3(ray2)
    if m == METHODS["DepthMeasurement"]:
        with np.errstate(all="ignore"):
            a1 = ray1 * (f32(d1) / ray1[..., 2])[..., None]
            a2 = ray2 * (f32(d2) / ray2[..., 2])[..., None]
        X1, X2 = triangulate_depth(a1, a2, T1w, T2w, loc)
    elif m == METHODS["Classic"]:
        X1, X2 = triangulate_classic(xn1, xn2, T1w, T2w, loc)
    elif m == METHODS["ORBSLAM"]:
        X1, X2 = triangulate_dlt(xn1, xn2, T1w, T2w, loc)
    else:
        X1, X2 = triangulate_nrslam(xn1, xn2, T1w, T2w, loc)
    X1 = X1.astype(np.float32)
    X2 = X2.astype(np.float32)
    c1 = T1w.apply(X1)
    c2 = T2w.apply(X2)
    w1 = normalize3(matvec(T1w.inverse().R, xn1))
    w2 = normalize3(matvec(T2w.inverse().R, xn2))
    with np.errstate(all="ignore"):
        cosp = cos_ray_parallax(w1, w2).astype(np.float32)
    n = uv1.shape[0]
    valid = np.ones(n, bool)
    if gate == GATE_SIM:            # Mapping.cc:351-364
        valid &= ~((c1[:, 2] < 0) | (c2[:, 2] < 0))
        valid &= cosp <= F(min_cos)
    elif gate == GATE_REAL:         # MonocularMapInitializer.cc:315-360
        fin = np.isfinite(X1).all(1) & np.isfinite(X2).all(1)
        nz = ~((X1 == 0).all(1) | (X2 == 0).all(1))
        valid &= fin & nz
        dl = F(depth_limit)
        valid &= ~((c1[:, 2] < 0) | (c1[:, 2] > dl))
        valid &= ~((c2[:, 2] < 0) | (c2[:, 2] > dl))
        if check_reproj:
            p1 = camera.project(cam1[0], cam1[1], c1)
            p2 = camera.project(cam2[0], cam2[1], c2)
            e1 = (uv1[:, 0] - p1[:, 0]) * (uv1[:, 0] - p1[:, 0]) + (uv1[:, 1] - p1[:, 1]) * (uv1[:, 1] - p1[:, 1])
            e2 = (uv2[:, 0] - p2[:, 0]) * (uv2[:, 0] - p2[:, 0]) + (uv2[:, 1] - p2[:, 1]) * (uv2[:, 1] - p2[:, 1])
            with np.errstate(all="ignore"):
                valid &= ~(e1.astype(np.float64) > 5.991) & ~(e2.astype(np.float64) > 5.991)
    return X1, X2, valid, cosp


def init_depth_scale_sim(d, X, Tcw, valid):
    """KeyFrame::setInitialDepthScaleInSimulationImages (KeyFrame.cc:131-153):
    mean over mapped points with non-zero measurement of d / z_c (float division,
    double accumulation)."""
    d = f32(d)
    zc = Tcw.apply(f32(X))[:, 2]
    use = valid & (d != 0)
    with np.errstate(all="ignore"):
        ratio = (d / zc).astype(np.float64)
    n = F(np.count_nonzero(use))
    return float(np.sum(ratio[use]) / np.float64(n))
