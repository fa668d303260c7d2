"""fp64 SE(3) helpers restating g2o::SE3Quat (oracle = test infrastructure).

g2o is NOT in /root/reference (Thirdparty/ is git-ignored upstream; version
unpinned).  Restated from upstream g2o `types/slam3d/se3quat.h` as used by the
reference through g2o::VertexSE3Expmap (g2oBundleAdjustment.cc:701-704) and
g2o::SE3Quat camera poses (:788-789):
  * update vector = [omega(3), upsilon(3)], oplus: T <- exp(update) * T
  * exp(): Rodrigues with the theta < 1e-5 second-order branch
  * quaternion storage (x, y, z, w), normalised with w >= 0.
"""
import numpy as np


def quat_to_rot(q):
    """(x,y,z,w) unit quaternion -> 3x3 (Eigen::Quaternion::toRotationMatrix)."""
    x, y, z, w = [float(v) for v in q]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]], np.float64)


def rot_to_quat(R):
    """Eigen's quaternion-from-matrix (trace based), returns (x,y,z,w)."""
    R = np.asarray(R, np.float64)
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def quat_normalize(q):
    """SE3Quat::normalizeRotation: w >= 0, unit norm."""
    q = np.asarray(q, np.float64).copy()
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], np.float64)


def se3_exp(update):
    """g2o::SE3Quat::exp -> (q(x,y,z,w), t)."""
    update = np.asarray(update, np.float64)
    omega, upsilon = update[:3], update[3:]
    theta = np.linalg.norm(omega)
    Om = skew(omega)
    Om2 = Om @ Om
    I = np.eye(3)
    if theta < 0.00001:
        R = I + Om + 0.5 * Om2
        V = I + 0.5 * Om + (1.0 / 6.0) * Om2
    else:
        R = I + np.sin(theta) / theta * Om + (1 - np.cos(theta)) / (theta * theta) * Om2
        V = I + (1 - np.cos(theta)) / (theta * theta) * Om + (theta - np.sin(theta)) / (theta ** 3) * Om2
    return quat_normalize(rot_to_quat(R)), V @ upsilon


class SE3:
    """g2o::SE3Quat: x' = R(q) x + t."""

    def __init__(self, q=(0, 0, 0, 1), t=(0, 0, 0)):
        self.q = quat_normalize(q)
        self.t = np.asarray(t, np.float64).copy()

    def R(self):
        return quat_to_rot(self.q)

    def map(self, X):
        return np.asarray(X, np.float64) @ self.R().T + self.t

    def oplus(self, update):
        """VertexSE3Expmap::oplusImpl: T <- exp(update) * T."""
        dq, dt = se3_exp(update)
        q = quat_normalize(quat_mul(dq, self.q))
        t = quat_to_rot(dq) @ self.t + dt
        return SE3(q, t)

    def as7(self):
        return np.concatenate([self.q, self.t])

    @staticmethod
    def from7(v):
        v = np.asarray(v, np.float64)
        return SE3(v[:4], v[4:])

    @staticmethod
    def from_pose32(pose):
        """Sophus::SE3f -> g2o::SE3Quat(unit_quaternion().cast<double>(), translation().cast<double>())."""
        return SE3(rot_to_quat(pose.R.astype(np.float64)), pose.t.astype(np.float64))
