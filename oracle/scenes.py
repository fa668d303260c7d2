"""Synthetic inputs for the hot path (oracle = test infrastructure).

Config 1 restates the reference's simulation front end:
  SLAM::loadPoints / setCameraPoses / getSimulatedDepthMeasurements / createKeyPoints / lookAt
  (/root/reference/Modules/System/SLAM.cc:172-351), roundToDecimals (Utils/Conversions.cc:64-67),
  including libstdc++'s std::default_random_engine (= minstd_rand0) and
  std::normal_distribution<float> (Marsaglia polar) so the key points are the
  ones the reference binary would draw.
Configs 2-5 are the synthetic shapes SURVEY.md 8d / BASELINE.md 4 define
(Data/Scripts/synthetic/create_data.py:27-126 distribution, kNN graphs).
"""
import numpy as np
from .f32 import f32, F, Pose, cross3, normalize3, matvec
from . import camera
from .graph import delaunay_graph, knn_graph, compute_rotations
from .triangulate import triangulate_pairs, init_depth_scale_sim, GATE_SIM, GATE_NONE
from .edges import Problem, Weights
from .se3 import SE3

SIM_CAM = np.array([458.654, 457.296, 367.215, 248.375, 0, 0, 0, 0], np.float32)      # Data/Simulation.yaml
DRUNKARD_CAM = np.array([190.68059285, 190.68059285, 160.0, 160.0, 0, 0, 0, 0], np.float32)  # Data/Drunkard.yaml:5-12
REALCOLON_CAM = np.array([727.1851, 728.5954, 738.1817, 537.4003, -0.1311029, -0.005149247,
                          0.001512357, -6.998448e-05], np.float32)                        # Data/Realcolon.yaml:15-38


# ------------------------------------------------------------------ libstdc++ RNG restatement
class MinStdRand0:
    """std::default_random_engine in libstdc++ = minstd_rand0 (a=16807, m=2^31-1), default seed 1."""

    def __init__(self, seed=1):
        self.x = seed % 2147483647 or 1

    def __call__(self):
        self.x = (self.x * 16807) % 2147483647
        return self.x

    def canonical_float(self):
        """std::generate_canonical<float,24> (one draw: log2(range) > 24)."""
        s = F(self() - 1) * F(1.0)
        r = s / F(2147483646.0)
        if r >= F(1.0):
            r = np.nextafter(F(1.0), F(0.0))
        return r


class NormalFloat:
    """std::normal_distribution<float> of libstdc++ (polar method with a saved value)."""

    def __init__(self, mean, stddev):
        self.mean, self.std = F(mean), F(stddev)
        self.saved = None

    def __call__(self, g):
        if self.saved is not None:
            ret = self.saved
            self.saved = None
        else:
            while True:
                x = F(2.0) * g.canonical_float() - F(1.0)
                y = F(2.0) * g.canonical_float() - F(1.0)
                r2 = x * x + y * y
                if not (r2 > F(1.0) or r2 == F(0.0)):
                    break
            lg = F(np.log(np.float64(r2)))                 # logf, emulated
            mult = np.sqrt(F(-2.0) * lg / r2)              # sqrtf(-2 * logf(r2) / r2)
            self.saved = x * mult
            ret = y * mult
        return ret * self.std + self.mean


def round_to_decimals(value, decimals):
    """Conversions.cc:64-67 (double arithmetic, std::round = half away from zero)."""
    factor = 10.0 ** decimals
    v = np.float64(value) * factor
    return np.float64(np.sign(v) * np.floor(np.abs(v) + 0.5)) / factor


def look_at(camera_pos, target_pos, up=(0, 1, 0)):
    """SLAM::lookAt (SLAM.cc:340-351), float32."""
    c, t, up = f32(camera_pos), f32(target_pos), f32(up)
    fwd = normalize3(t - c)
    right = normalize3(cross3(up, fwd))
    upv = normalize3(cross3(fwd, right))
    R = np.stack([right, upv, fwd], axis=1).astype(np.float32)
    return R


def simulation_poses(C1, C2, moved0):
    """SLAM::setCameraPoses (SLAM.cc:223-235): the camera centre is stored as the
    translation of Tcw [sic]."""
    T1 = Pose(np.eye(3, dtype=np.float32), f32(C1))
    T2 = Pose(look_at(C2, moved0), f32(C2))
    return T1, T2


def simulation_frontend(original, moved, C1, C2, cam=SIM_CAM, rep_sigma=1.0, decimals=1,
                        depth_sigma_mm=3.0, scale_c1=0.4, scale_c2=1.7, model=camera.KB8):
    """Key points and depth measurements exactly as the reference draws them.  model = camera.KB8 (what HEAD's
    Settings.cc:43-51 builds) or camera.PINHOLE (what the revision that wrote Data/SinteticDataBase and
    Data/Experiments used: tests/golden/reference_pins.json reproduces those logs to 6 digits with it)."""
    original, moved = f32(original), f32(moved)
    T1, T2 = simulation_poses(C1, C2, moved[0])
    n = original.shape[0]
    # getSimulatedDepthMeasurements (own generator, default seed)
    g = MinStdRand0()
    dist = NormalFloat(0.0, F(depth_sigma_mm) / F(1000))
    c1 = T1.apply(original)
    c2 = T2.apply(moved)
    d1 = np.zeros(n, np.float32)
    d2 = np.zeros(n, np.float32)
    for i in range(n):
        d1[i] = c1[i, 2] * F(scale_c1) + dist(g)
        d2[i] = c2[i, 2] * F(scale_c2) + dist(g)
    # createKeyPoints (own generator, default seed)
    g = MinStdRand0()
    dist = NormalFloat(0.0, F(rep_sigma))
    p1 = camera.project(model, cam, c1)
    p2 = camera.project(model, cam, c2)
    uv1 = np.zeros((n, 2), np.float32)
    uv2 = np.zeros((n, 2), np.float32)
    for i in range(n):
        uv1[i, 0] = round_to_decimals(p1[i, 0] + dist(g), decimals)
        uv1[i, 1] = round_to_decimals(p1[i, 1] + dist(g), decimals)
        uv2[i, 0] = round_to_decimals(p2[i, 0] + dist(g), decimals)
        uv2[i, 1] = round_to_decimals(p2[i, 1] + dist(g), decimals)
    return dict(T1=T1, T2=T2, uv1=uv1, uv2=uv2, d1=d1, d2=d2, cam=f32(cam))


def pixel_sigma(cam, T, X, uv):
    """calculatePixelsStandDev (Utils/Geometry.cc:370-498) for one camera: mean over (u,v) of
    sqrt(mean |obs - project(T * X)|^2); projection in float32 from the float MapPoint position."""
    proj = camera.project(cam[0], cam[1], T.apply(f32(X)))
    err = np.abs(f32(uv).astype(np.float64) - proj.astype(np.float64))
    return float(np.sqrt((err ** 2).mean(0)).mean())


def load_points_csv(path):
    return np.loadtxt(path, dtype=np.float64).astype(np.float32)


# ------------------------------------------------------------------ problem assembly
def build_problem(uv1, uv2, d1, d2, cam, T1, T2, graph_kind="delaunay", k=8, area=None,
                  method="NRSLAM", location="FarPoints", min_cos=0.9998, gate=GATE_SIM,
                  scale_init=(0.0, 0.0), model=camera.KB8):
    """Triangulate (K1), compact away rejected correspondences, build the neighbour graph on
    KF1's points, initial depth scales and per-vertex rotations: everything arapOptimization
    sets up before optimizer.optimize() (g2oBundleAdjustment.cc:640-957)."""
    camt = (model, f32(cam))
    X1, X2, valid, cosp = triangulate_pairs(uv1, uv2, camt, camt, T1, T2, method, location, gate, min_cos,
                                            d1=d1, d2=d2)
    keep = np.nonzero(valid)[0]
    X1, X2 = X1[keep], X2[keep]
    uv1, uv2, d1, d2 = uv1[keep], uv2[keep], f32(d1)[keep], f32(d2)[keep]
    n = len(keep)
    s1 = scale_init[0] or init_depth_scale_sim(d1, X1, T1, np.ones(n, bool))
    s2 = scale_init[1] or init_depth_scale_sim(d2, X2, T2, np.ones(n, bool))
    X1d, X2d = X1.astype(np.float64), X2.astype(np.float64)
    if graph_kind == "delaunay":
        g = delaunay_graph(X1d)
    else:
        g = knn_graph(X1d, k, area)
    R = compute_rotations(g, X1d, X2d)
    p = Problem(cam1=camt, cam2=camt, T1=T1, T2=T2, uv1=uv1, uv2=uv2,
                inv_sigma2_1=np.ones(n), inv_sigma2_2=np.ones(n),
                d1=d1.astype(np.float64), d2=d2.astype(np.float64), graph=g, X1=X1d, X2=X2d,
                Tg=SE3(), s1=float(s1), s2=float(s2), R=R)
    return p, keep


def config1(original_csv, moved_csv, C1=(-0.10, 0.02, 0.12), C2=(0.14, 0.01, 0.06), **kw):
    """BASELINE.json configs[0]: Execution/simulation on Data/Simulation.yaml."""
    o, m = load_points_csv(original_csv), load_points_csv(moved_csv)
    fe = simulation_frontend(o, m, C1, C2, **kw)
    p, keep = build_problem(fe["uv1"], fe["uv2"], fe["d1"], fe["d2"], fe["cam"], fe["T1"], fe["T2"])
    w = Weights(rep=1.0, arap=200000.0, depth_sigma=0.003, glob=50.0)
    return p, w, fe, keep


def sheet_scene(n, seed=0, cam=SIM_CAM, rigid=0.0025, gauss=0.0025, px_sigma=1.0, depth_sigma=0.003,
                scales=(0.4, 1.7), C1=(-0.10, 0.02, 0.12), C2=(0.14, 0.01, 0.06)):
    """Config-2 generator: create_data.py distribution widened to a non-degenerate sheet
    (x,y ~ N(0,0.03), z ~ N(0.2,0.01) after Rz(45)Rx(-45)), rigid + Gaussian deformation,
    cameras of Simulation.yaml, N(0,px_sigma) pixel noise rounded to 0.1 px, noisy scaled depth."""
    rng = np.random.default_rng(seed)
    o = np.stack([rng.normal(0, 0.03, n), rng.normal(0, 0.03, n), rng.normal(0, 0.01, n)], 1)
    m = o.copy()
    m[:, 1] += rigid
    m += rng.normal(0, gauss, (n, 3)) if gauss > 0 else 0.0
    ax, az = np.deg2rad(-45), np.deg2rad(45)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    Rm = Rz @ Rx
    o = o @ Rm.T + np.array([0, 0, 0.2])
    m = m @ Rm.T + np.array([0, 0, 0.2])
    o, m = o.astype(np.float32), m.astype(np.float32)
    T1, T2 = simulation_poses(C1, C2, m[0])
    c1, c2 = T1.apply(o), T2.apply(m)
    uv1 = camera.kb8_project(cam, c1) + rng.normal(0, px_sigma, (n, 2)).astype(np.float32)
    uv2 = camera.kb8_project(cam, c2) + rng.normal(0, px_sigma, (n, 2)).astype(np.float32)
    uv1 = (np.round(uv1.astype(np.float64) * 10) / 10).astype(np.float32)
    uv2 = (np.round(uv2.astype(np.float64) * 10) / 10).astype(np.float32)
    d1 = (c1[:, 2] * F(scales[0]) + rng.normal(0, depth_sigma, n).astype(np.float32)).astype(np.float32)
    d2 = (c2[:, 2] * F(scales[1]) + rng.normal(0, depth_sigma, n).astype(np.float32)).astype(np.float32)
    area = float(np.pi * (3 * 0.03) ** 2)          # 3-sigma disc of the sheet
    return dict(T1=T1, T2=T2, uv1=uv1, uv2=uv2, d1=d1, d2=d2, cam=f32(cam), original=o, moved=m, area=area)


def tube_scene(n, seed=0, cam=DRUNKARD_CAM, radius=0.03, length=0.3, motion=0.035, wave=0.005,
               noise=0.0005, px_sigma=1.0, depth_sigma=0.0003, scales=(1.3, 0.8), width=320, height=320):
    """Config-3/4 generator: colon-like tube around the optical axis, camera on the axis looking
    down the tube, second camera advanced by `motion`; deformation = radial peristaltic wave of
    amplitude `wave` + N(0, noise)."""
    rng = np.random.default_rng(seed)
    zs = 0.03 + length * rng.random(n)
    th = 2 * np.pi * rng.random(n)
    o = np.stack([radius * np.cos(th), radius * np.sin(th), zs], 1)
    rad = radius + wave * np.sin(2 * np.pi * zs / 0.1 + 1.0) * (0.5 + 0.5 * np.cos(th))
    m = np.stack([rad * np.cos(th), rad * np.sin(th), zs], 1) + rng.normal(0, noise, (n, 3))
    o, m = o.astype(np.float32), m.astype(np.float32)
    T1 = Pose(np.eye(3, dtype=np.float32), f32([0, 0, 0]))
    ang = 0.02
    R2 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    T2 = Pose(R2, f32([0.004, -0.003, -motion * 0.3]))
    c1, c2 = T1.apply(o), T2.apply(m)
    uv1 = camera.kb8_project(cam, c1) + rng.normal(0, px_sigma, (n, 2)).astype(np.float32)
    uv2 = camera.kb8_project(cam, c2) + rng.normal(0, px_sigma, (n, 2)).astype(np.float32)
    d1 = (c1[:, 2] * F(scales[0]) + rng.normal(0, depth_sigma, n).astype(np.float32)).astype(np.float32)
    d2 = (c2[:, 2] * F(scales[1]) + rng.normal(0, depth_sigma, n).astype(np.float32)).astype(np.float32)
    area = float(2 * np.pi * radius * length)
    return dict(T1=T1, T2=T2, uv1=uv1.astype(np.float32), uv2=uv2.astype(np.float32), d1=d1, d2=d2,
                cam=f32(cam), original=o, moved=m, area=area, width=width, height=height)


def problem_from_scene(sc, graph_kind="knn", k=8, **kw):
    return build_problem(sc["uv1"], sc["uv2"], sc["d1"], sc["d2"], sc["cam"], sc["T1"], sc["T2"],
                         graph_kind=graph_kind, k=k, area=sc.get("area"), **kw)
