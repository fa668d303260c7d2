"""Synthetic inputs of the shapes BASELINE.json names (host side, numpy; no part of the timed path).

The generators only *synthesise* measurements (true surface points -> noisy key points and depths);
triangulation and refinement of those measurements is done by the CUDA library.  Distributions follow
SURVEY.md section 8d: config 2 = create_data.py sheet (Data/Scripts/synthetic/create_data.py:27-126,
widened), config 3 = Drunkard.yaml-shaped tube, config 4 = Realcolon.yaml-shaped tube with a border
mask, config 5 = many small config-2 problems.
"""
import numpy as np

SIM_CAM = np.array([458.654, 457.296, 367.215, 248.375, 0, 0, 0, 0], np.float32)            # Data/Simulation.yaml
DRUNKARD_CAM = np.array([190.68059285, 190.68059285, 160.0, 160.0, 0, 0, 0, 0], np.float32)  # Data/Drunkard.yaml:5-12
REALCOLON_CAM = np.array([727.1851, 728.5954, 738.1817, 537.4003, -0.1311029, -0.005149247,
                          0.001512357, -6.998448e-05], np.float32)                           # Data/Realcolon.yaml:15-38


def kb8_project64(P, Xc):
    """Kannala-Brandt projection in float64 -- data synthesis only (the measured path projects on the GPU)."""
    P = np.asarray(P, np.float64)
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    th = np.arctan2(np.sqrt(x * x + y * y), z)
    psi = np.arctan2(y, x)
    t2 = th * th
    r = th * (1 + t2 * (P[4] + t2 * (P[5] + t2 * (P[6] + t2 * P[7]))))
    return np.stack([P[0] * r * np.cos(psi) + P[2], P[1] * r * np.sin(psi) + P[3]], 1)


def _pose(R, t):
    return np.concatenate([np.asarray(R, np.float32).reshape(3, 3), np.asarray(t, np.float32).reshape(3, 1)], 1)


def _look_at(c, target, up=(0.0, 1.0, 0.0)):
    """SLAM::lookAt (Modules/System/SLAM.cc:340-351)."""
    c, target, up = np.asarray(c, np.float64), np.asarray(target, np.float64), np.asarray(up, np.float64)
    f = target - c
    f /= np.linalg.norm(f)
    r = np.cross(up, f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    u /= np.linalg.norm(u)
    return np.stack([r, u, f], 1)


def _measure(o, m, T1, T2, cam, rng, px_sigma, depth_sigma, scales, round_px):
    c1 = o @ T1[:, :3].T.astype(np.float64) + T1[:, 3].astype(np.float64)
    c2 = m @ T2[:, :3].T.astype(np.float64) + T2[:, 3].astype(np.float64)
    n = o.shape[0]
    uv1 = kb8_project64(cam, c1) + rng.normal(0, px_sigma, (n, 2))
    uv2 = kb8_project64(cam, c2) + rng.normal(0, px_sigma, (n, 2))
    if round_px:
        uv1, uv2 = np.round(uv1 * 10) / 10, np.round(uv2 * 10) / 10      # Keypoints.decimalsApproximation: 1
    d1 = c1[:, 2] * scales[0] + rng.normal(0, depth_sigma, n)
    d2 = c2[:, 2] * scales[1] + rng.normal(0, depth_sigma, n)
    return uv1.astype(np.float32), uv2.astype(np.float32), d1.astype(np.float32), d2.astype(np.float32)


def sheet_scene(n, seed=0, cam=SIM_CAM, rigid=0.0025, gauss=0.0025, px_sigma=1.0, depth_sigma=0.003, scales=(0.4, 1.7)):
    """Configs 2 and 5: deforming sheet in front of the Simulation.yaml cameras."""
    rng = np.random.default_rng(seed)
    o = np.stack([rng.normal(0, 0.03, n), rng.normal(0, 0.03, n), rng.normal(0, 0.01, n)], 1)
    m = o.copy()
    m[:, 1] += rigid
    m += rng.normal(0, gauss, (n, 3))
    ax, az = np.deg2rad(-45), np.deg2rad(45)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    Rm = Rz @ Rx
    o = o @ Rm.T + np.array([0, 0, 0.2])
    m = m @ Rm.T + np.array([0, 0, 0.2])
    C1, C2 = np.array([-0.10, 0.02, 0.12]), np.array([0.14, 0.01, 0.06])
    T1 = _pose(np.eye(3), C1)                               # SLAM::setCameraPoses stores the centre as t [sic]
    T2 = _pose(_look_at(C2, m[0]), C2)
    uv1, uv2, d1, d2 = _measure(o, m, T1, T2, cam, rng, px_sigma, depth_sigma, scales, True)
    return dict(cam=np.asarray(cam, np.float32), T1=T1, T2=T2, uv1=uv1, uv2=uv2, d1=d1, d2=d2,
                area=float(np.pi * 0.09 ** 2), original=o, moved=m, min_cos=0.9998,
                weights=dict(rep=1.0, arap=200000.0, depth_sigma=0.003), lm_iters=25)


def tube_scene(n, seed=0, cam=DRUNKARD_CAM, radius=0.03, length=0.3, motion=0.035, wave=0.005, noise=0.0005,
               px_sigma=1.0, depth_sigma=0.0003, scales=(1.3, 0.8), arap=1.0e7, lm_iters=30):
    """Configs 3 and 4: colon-like tube, camera near the axis, peristaltic radial wave + noise."""
    rng = np.random.default_rng(seed)
    zs = 0.03 + length * rng.random(n)
    th = 2 * np.pi * rng.random(n)
    o = np.stack([radius * np.cos(th), radius * np.sin(th), zs], 1)
    rad = radius + wave * np.sin(2 * np.pi * zs / 0.1 + 1.0) * (0.5 + 0.5 * np.cos(th))
    m = np.stack([rad * np.cos(th), rad * np.sin(th), zs], 1) + rng.normal(0, noise, (n, 3))
    T1 = _pose(np.eye(3), [0, 0, 0])
    ang = 0.02
    R2 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    T2 = _pose(R2, [0.3 * motion, -0.1 * motion, -0.3 * motion])
    uv1, uv2, d1, d2 = _measure(o, m, T1, T2, cam, rng, px_sigma, depth_sigma, scales, False)
    return dict(cam=np.asarray(cam, np.float32), T1=T1, T2=T2, uv1=uv1, uv2=uv2, d1=d1, d2=d2,
                area=float(2 * np.pi * radius * length), original=o, moved=m, min_cos=0.99999,
                weights=dict(rep=1.0, arap=arap, depth_sigma=depth_sigma), lm_iters=lm_iters)


def border_mask_keep(uv, width, height, frac=0.124, kernel=13):
    """Config 4: drop key points under an endoscope-style border mask (everything outside a centred
    disc, ~12.4 % of the image as in mask_border_endo_ori.jpg) dilated by the level-0 FAST kernel
    (side 13, Modules/Features/FAST.cc:510-519)."""
    cx, cy = 0.5 * width, 0.5 * height
    r_keep = np.sqrt((1.0 - frac) * width * height / np.pi) - 0.5 * kernel
    px = np.round(uv)
    inside = (px[:, 0] >= 0) & (px[:, 0] < width) & (px[:, 1] >= 0) & (px[:, 1] < height)
    return inside & (((px[:, 0] - cx) / (width / height)) ** 2 + (px[:, 1] - cy) ** 2 <= (r_keep * height / np.sqrt(width * height)) ** 2)


def knn_graph(xy, k):
    """Symmetrised k-nearest-neighbour graph in the plane (x,y) -> CSR (rowptr, col ascending, unit w)."""
    from scipy.spatial import cKDTree
    xy = np.asarray(xy, np.float64)
    n = xy.shape[0]
    _, idx = cKDTree(xy).query(xy, k=k + 1, workers=-1)
    i = np.repeat(np.arange(n, dtype=np.int64), k + 1)
    j = idx.reshape(-1).astype(np.int64)
    keep = i != j
    i, j = i[keep], j[keep]
    key = np.unique(np.concatenate([i * n + j, j * n + i]))
    ii = (key // n).astype(np.int32)
    jj = (key % n).astype(np.int32)
    rowptr = np.zeros(n + 1, np.int64)
    np.add.at(rowptr, ii.astype(np.int64) + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, jj, np.ones(len(jj), np.float64)


def make_scene(workload, n, seed=0):
    """The bench / full-size-parity workloads by name: 'drunkard' (config 3), 'realcolon' (config 4: distorted camera +
    border mask; build the graph with k = 16), 'sheet' (configs 2 and 5).  Generates ~8 % more matches than n so that
    n survive the triangulation gates."""
    n_gen = int(n * 1.08) + 64
    if workload == "drunkard":
        sc = tube_scene(n_gen, seed=seed, cam=DRUNKARD_CAM, arap=1.0e7, depth_sigma=0.0003, lm_iters=30)
        sc["name"] = "config3: Drunkard.yaml-shaped tube, 1 frame pair"
    elif workload == "realcolon":
        sc = tube_scene(n_gen, seed=seed, cam=REALCOLON_CAM, arap=0.1, depth_sigma=1e-6, lm_iters=30, scales=(1.0, 1.0))
        keep = border_mask_keep(sc["uv1"], 1440, 1080) & border_mask_keep(sc["uv2"], 1440, 1080)
        for key in ("uv1", "uv2", "d1", "d2", "original", "moved"):
            sc[key] = sc[key][keep]
        sc["name"] = "config4: Realcolon.yaml-shaped tube + border mask, 1 frame pair"
    elif workload == "sheet":
        sc = sheet_scene(n_gen, seed=seed)
        sc["name"] = "config2: Simulation.yaml sheet, 1 frame pair"
    else:
        raise ValueError(f"unknown workload {workload!r}")
    sc["workload"] = workload
    return sc


def _rodrigues(w):
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / (th * th) * (K @ K)


def _quat_of(R):
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2      # small rotations only: w > 0
    return np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])


def ba_scene(n_points, n_views=2, seed=0, cam=SIM_CAM, px_sigma=1.0, pose_noise=0.003, point_noise=0.001):
    """Classic bundle adjustment: n_views key frames looking at a sheet of n_points map points, every point seen
    from every view; noisy key points (1 px, one decimal), perturbed poses (key frame 0 exact and fixed) and points."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.normal(0, 0.05, n_points), rng.normal(0, 0.05, n_points), rng.normal(0.35, 0.02, n_points)], 1)
    poses7, obs_pose, obs_point, uv = [], [], [], []
    for k in range(n_views):
        R = _rodrigues(rng.normal(0, 0.05, 3) * (k > 0))
        t = np.array([0.04 * k, 0.01 * k, 0.005 * k])
        pr = kb8_project64(cam, X @ R.T + t)
        uv.append(np.round((pr + rng.normal(0, px_sigma, pr.shape)) * 10) / 10)
        obs_pose.append(np.full(n_points, k, np.int32))
        obs_point.append(np.arange(n_points, dtype=np.int32))
        if k > 0:
            R = _rodrigues(rng.normal(0, pose_noise, 3)) @ R
            t = t + rng.normal(0, pose_noise, 3)
        poses7.append(np.concatenate([_quat_of(R), t]))
    fixed = np.zeros(n_views, bool)
    fixed[0] = True
    return dict(poses7=np.array(poses7), pose_fixed=fixed, cams=[(0, np.asarray(cam, np.float32))] * n_views,
                X=X + rng.normal(0, point_noise, X.shape), obs_pose=np.concatenate(obs_pose), obs_point=np.concatenate(obs_point),
                obs_uv=np.concatenate(uv).astype(np.float32), obs_isg=np.ones(n_views * n_points, np.float32))
