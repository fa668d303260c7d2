"""float32 building blocks with a FIXED operation order (oracle = test infrastructure).

The reference does its camera and triangulation maths in float32 through Eigen
expression templates (Modules/Utils/Geometry.cc, Modules/Calibration/*.cc).  The
exact rounding order of those templates is compiler dependent and the reference
cannot be built here, so the oracle *defines* the order once, and the CUDA
kernels restate the very same order with __fmul_rn/__fadd_rn (no FMA
contraction), which makes the float32 stages bit-comparable:

  dot3(a,b)   = (a0*b0 + a1*b1) + a2*b2
  matvec(R,v) = row-wise dot3
  libm calls  = evaluated in double and rounded once to float32 ("emu"), i.e.
                the correctly rounded float result up to double rounding
                (p ~ 2^-29); glibc's atan2f/sinf/cosf are within 1 ulp of that.
"""
import numpy as np

F = np.float32


def f32(x):
    return np.asarray(x, dtype=np.float32)


def emu(fn, *args):
    """libm call on float32 data: evaluate in double, round once to float32."""
    with np.errstate(all="ignore"):
        return fn(*[np.asarray(a, np.float64) for a in args]).astype(np.float32)


def dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def cross3(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def norm3(a):
    with np.errstate(all="ignore"):
        return np.sqrt(dot3(a, a))          # IEEE sqrt: correctly rounded in numpy and CUDA


def normalize3(a):
    with np.errstate(all="ignore"):
        return a / norm3(a)[..., None]


def matvec(R, v):
    """R: (3,3) float32, v: (...,3) float32."""
    return np.stack([(R[i, 0] * v[..., 0] + R[i, 1] * v[..., 1]) + R[i, 2] * v[..., 2]
                     for i in range(3)], axis=-1)


def matTvec(R, v):
    return np.stack([(R[0, i] * v[..., 0] + R[1, i] * v[..., 1]) + R[2, i] * v[..., 2]
                     for i in range(3)], axis=-1)


def matmul33(A, B):
    out = np.zeros((3, 3), np.float32)
    for i in range(3):
        for j in range(3):
            out[i, j] = (A[i, 0] * B[0, j] + A[i, 1] * B[1, j]) + A[i, 2] * B[2, j]
    return out


class Pose:
    """Rigid transform x_c = R x_w + t in float32 (the reference's Sophus::SE3f Tcw)."""

    def __init__(self, R, t):
        self.R = f32(R).reshape(3, 3).copy()
        self.t = f32(t).reshape(3).copy()

    def apply(self, X):
        return matvec(self.R, f32(X)) + self.t

    def inverse(self):
        Rt = self.R.T.copy()
        return Pose(Rt, -matvec(Rt, self.t))

    def compose(self, other):
        """self * other."""
        return Pose(matmul33(self.R, other.R), matvec(self.R, other.t) + self.t)

    def as34(self):
        return np.concatenate([self.R, self.t[:, None]], axis=1).astype(np.float32)

    @staticmethod
    def from34(M):
        M = f32(M).reshape(3, 4)
        return Pose(M[:, :3], M[:, 3])
