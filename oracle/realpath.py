"""Real-image map initialisation restated (oracle = test infrastructure).

Follows /root/reference/Modules/Mapping/Mapping.cc:183-254 (MapPoint creation gates, slot / observation index
convention, initial depth scales), MonocularMapInitializer.cc:303-368 (triangulation + gates, through
oracle.triangulate), Modules/Map/KeyFrame.cc:181-202 and Modules/Utils/Geometry.cc:607-619 (depth-image sampling).
"""
import numpy as np

from . import camera
from .f32 import f32, F, Pose
from .triangulate import triangulate_pairs, GATE_REAL


def interpolate(x, y, im):
    """Geometry.cc:607-619 bilinear sample in float32 (weights w00, w01, w10, w11 = 1 - the rest)."""
    x, y = f32(x), f32(y)
    xi, yi = np.floor(x).astype(np.float32), np.floor(y).astype(np.float32)     # modf for non-negative pixels
    fx, fy = (x - xi).astype(np.float32), (y - yi).astype(np.float32)
    one = F(1.0)
    w00 = (one - fx) * (one - fy)
    w01 = (one - fx) * fy
    w10 = fx * (one - fy)
    w11 = one - w00 - w01 - w10
    X, Y = xi.astype(np.int64), yi.astype(np.int64)
    flat = f32(im).reshape(-1)
    cols = im.shape[1]

    def at(yy, xx):
        k = yy * cols + xx
        return np.where(k < flat.size, flat[np.minimum(k, flat.size - 1)], F(0))
    return (at(Y, X) * w00 + at(Y, X + 1) * w10 + at(Y + 1, X) * w01 + at(Y + 1, X + 1) * w11).astype(np.float32)


def depth_measure(im, x, y, scaled, image_depth_scale):
    """KeyFrame::getDepthMeasure(x, y, scaled) (KeyFrame.cc:181-202): sample / 100 [* imageDepthScale]."""
    d = interpolate(x, y, im).astype(np.float64) / 100
    return d if scaled else d * image_depth_scale


def init_from_matches(kp1, kp2, matches, cam, T1, T2, im1, im2, ids1, ids2, method="NRSLAM", location="TwoPoints",
                      depth_limit=np.inf, check_reproj=False, min_cos=0.9998):
    """Returns dict(slots, X1, X2, uv1, uv2, d1, d2, s1, s2): the correspondences Mapping.cc:183-254 inserts, in
    REF-slot order, the observation of the curr MapPoint being the MATCHED key point."""
    kp1, kp2 = f32(kp1), f32(kp2)
    ref = np.nonzero(matches >= 0)[0]
    uv1, uv2 = kp1[ref], kp2[matches[ref]]
    camt = (camera.KB8, f32(cam))
    X1, X2, valid, cosp = triangulate_pairs(uv1, uv2, camt, camt, T1, T2, method, location, GATE_REAL, 1.0,
                                            depth_limit=depth_limit, check_reproj=check_reproj)
    d1s = depth_measure(im1, uv1[:, 0], uv1[:, 1], True, ids1)
    d2s = depth_measure(im2, uv2[:, 0], uv2[:, 1], True, ids2)
    ok = valid & ~((d1s <= 0.0) | (d2s <= 0.0))
    for uv in (uv1, uv2):
        ok &= ~((uv[:, 0] <= F(0.1)) | (uv[:, 0] >= F(1500)) | (uv[:, 1] <= F(0.1)) | (uv[:, 1] >= F(1500)))
    with np.errstate(all="ignore"):
        deg = (np.arccos(cosp.astype(np.float64)).astype(np.float32) * F(180.0 / np.pi)).astype(np.float32)
    use = ok & (deg > F(min_cos))
    d1u = depth_measure(im1, uv1[:, 0], uv1[:, 1], False, ids1)
    d2u = depth_measure(im2, uv2[:, 0], uv2[:, 1], False, ids2)
    z1 = T1.apply(X1)[:, 2].astype(np.float64)
    z2 = T2.apply(X2)[:, 2].astype(np.float64)
    n_pts = np.float32(np.count_nonzero(use))
    s1 = float(np.sum(d1u[use] / z1[use]) / np.float64(n_pts))
    s2 = float(np.sum(d2u[use] / z2[use]) / np.float64(n_pts))
    return dict(slots=ref[ok], X1=X1[ok], X2=X2[ok], uv1=uv1[ok], uv2=uv2[ok], d1=d1u[ok], d2=d2u[ok], s1=s1, s2=s2)
