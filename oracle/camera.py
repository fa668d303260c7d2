"""Camera models in float32 (oracle = test infrastructure).

Follows /root/reference/Modules/Calibration/KannalaBrandt8.cc:32-114 and
PinHole.cc:25-59.  params = float32[8] = fx, fy, cx, cy, k0..k3
(Modules/System/Settings.cc:38-50).  libm calls are "emu" (see f32.py).
"""
import numpy as np
from .f32 import f32, emu, F

KB8, PINHOLE = 0, 1


def kb8_project(P, Xc):
    """KannalaBrandt8.cc:32-49.  Xc float32 (...,3) camera-frame point -> (...,2) pixel."""
    P = f32(P)
    Xc = f32(Xc)
    x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    x2_plus_y2 = x * x + y * y
    theta = emu(np.arctan2, emu(np.sqrt, x2_plus_y2), z)
    psi = emu(np.arctan2, y, x)
    theta2 = theta * theta
    theta3 = theta * theta2
    theta5 = theta3 * theta2
    theta7 = theta5 * theta2
    theta9 = theta7 * theta2
    r = (((theta + P[4] * theta3) + P[5] * theta5) + P[6] * theta7) + P[7] * theta9
    u = (P[0] * r) * emu(np.cos, psi) + P[2]
    v = (P[1] * r) * emu(np.sin, psi) + P[3]
    return np.stack([u, v], axis=-1)


def kb8_project_double_trig(P, Xc):
    """The OTHER reading of KannalaBrandt8.cc:47-48: `cos(psi)` / `sin(psi)` taken as ::cos(double) / ::sin(double) -- the float
    product fx * r promoted to double, multiplied, cx added in double, the sum rounded to float on assignment.  Not what
    libstdc++ resolves the unqualified call to (its <cmath> / <math.h> put the float overloads of std::cos into the global
    namespace, so `cos(float)` is cosf and kb8_project above is the reading); kept to MEASURE the difference
    (tests/test_oracle_pins.py::test_kb8_cos_overload_reading)."""
    P = f32(P)
    Xc = f32(Xc)
    x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    theta = emu(np.arctan2, emu(np.sqrt, x * x + y * y), z)
    psi = emu(np.arctan2, y, x)
    theta2 = theta * theta
    theta3 = theta * theta2
    theta5 = theta3 * theta2
    theta7 = theta5 * theta2
    theta9 = theta7 * theta2
    r = (((theta + P[4] * theta3) + P[5] * theta5) + P[6] * theta7) + P[7] * theta9
    psd = psi.astype(np.float64)
    u = ((P[0] * r).astype(np.float64) * np.cos(psd) + np.float64(P[2])).astype(np.float32)
    v = ((P[1] * r).astype(np.float64) * np.sin(psd) + np.float64(P[3])).astype(np.float32)
    return np.stack([u, v], axis=-1)


def kb8_unproject(P, uv):
    """KannalaBrandt8.cc:51-83.  <=10 Newton steps on theta, tol 1e-6 (KannalaBrandt8.h:29).

    Reference leaves `th` uninitialised when theta_d <= 1e-8 (:59-78) and then
    divides by theta_d; the oracle (and the CUDA path) DEFINE that case as the
    optical axis (0,0,1).
    """
    P = f32(P)
    uv = f32(uv)
    pwx = (uv[..., 0] - P[2]) / P[0]
    pwy = (uv[..., 1] - P[3]) / P[1]
    theta_d = emu(np.sqrt, pwx * pwx + pwy * pwy)
    theta = theta_d.copy()
    active = theta_d.astype(np.float64) > 1e-8      # float compared against a double literal
    one = F(1.0)
    for _ in range(10):
        theta2 = theta * theta
        theta4 = theta2 * theta2
        theta6 = theta4 * theta2
        theta8 = theta4 * theta4
        k0t2 = P[4] * theta2
        k1t4 = P[5] * theta4
        k2t6 = P[6] * theta6
        k3t8 = P[7] * theta8
        num = theta * ((((one + k0t2) + k1t4) + k2t6) + k3t8) - theta_d
        den = (((one + F(3) * k0t2) + F(5) * k1t4) + F(7) * k2t6) + F(9) * k3t8
        fix = num / den
        theta = np.where(active, theta - fix, theta)
        active = active & ~(np.abs(fix) < F(1e-6))
    s = emu(np.sin, theta)
    c = emu(np.cos, theta)
    ok = theta_d.astype(np.float64) > 1e-8
    with np.errstate(all="ignore"):
        rx = np.where(ok, s * pwx / theta_d, F(0))
        ry = np.where(ok, s * pwy / theta_d, F(0))
    rz = np.where(ok, c, F(1))
    return np.stack([rx, ry, rz], axis=-1).astype(np.float32)


def kb8_project_jac(P, Xc):
    """KannalaBrandt8.cc:85-114 -> (...,2,3) float32."""
    P = f32(P)
    Xc = f32(Xc)
    x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    fx, fy = P[0], P[1]
    x2 = x * x
    y2 = y * y
    z2 = z * z
    r2 = x2 + y2
    r = emu(np.sqrt, r2)
    r3 = r2 * r
    theta = emu(np.arctan2, r, z)
    theta2 = theta * theta
    theta3 = theta2 * theta
    theta4 = theta2 * theta2
    theta5 = theta4 * theta
    theta6 = theta2 * theta4
    theta7 = theta6 * theta
    theta8 = theta4 * theta4
    theta9 = theta8 * theta
    f = (((theta + theta3 * P[4]) + theta5 * P[5]) + theta7 * P[6]) + theta9 * P[7]
    fd = (((F(1) + (F(3) * P[4]) * theta2) + (F(5) * P[5]) * theta4)
          + (F(7) * P[6]) * theta6) + (F(9) * P[7]) * theta8
    den = r2 * (r2 + z2)
    with np.errstate(all="ignore"):
        J00 = fx * (((fd * z) * x2) / den + (f * y2) / r3)
        J01 = fx * ((((fd * z) * y) * x) / den - ((f * y) * x) / r3)
        J02 = (((-fx) * fd) * x) / (r2 + z2)
        J10 = fy * ((((fd * z) * y) * x) / den - ((f * y) * x) / r3)
        J11 = fy * (((fd * z) * y2) / den + (f * x2) / r3)
        J12 = (((-fy) * fd) * y) / (r2 + z2)
    J = np.stack([np.stack([J00, J01, J02], -1), np.stack([J10, J11, J12], -1)], -2)
    return J.astype(np.float32)


def pinhole_project(P, Xc):
    """PinHole.cc:25-33."""
    P = f32(P)
    Xc = f32(Xc)
    with np.errstate(all="ignore"):
        u = (P[0] * Xc[..., 0]) / Xc[..., 2] + P[2]
        v = (P[1] * Xc[..., 1]) / Xc[..., 2] + P[3]
    return np.stack([u, v], -1)


def pinhole_unproject(P, uv):
    """PinHole.cc:35-40."""
    P = f32(P)
    uv = f32(uv)
    return np.stack([(uv[..., 0] - P[2]) / P[0], (uv[..., 1] - P[3]) / P[1],
                     np.ones_like(uv[..., 0])], -1).astype(np.float32)


def pinhole_project_jac(P, Xc):
    """PinHole.cc:49-62."""
    P = f32(P)
    Xc = f32(Xc)
    x, y, z = Xc[..., 0], Xc[..., 1], Xc[..., 2]
    zero = np.zeros_like(x)
    with np.errstate(all="ignore"):
        J = np.stack([np.stack([P[0] / z, zero, ((-P[0]) * x) / (z * z)], -1),
                      np.stack([zero, P[1] / z, ((-P[1]) * y) / (z * z)], -1)], -2)
    return J.astype(np.float32)


def project(model, P, Xc):
    return kb8_project(P, Xc) if model == KB8 else pinhole_project(P, Xc)


def unproject(model, P, uv):
    return kb8_unproject(P, uv) if model == KB8 else pinhole_unproject(P, uv)


def project_jac(model, P, Xc):
    return kb8_project_jac(P, Xc) if model == KB8 else pinhole_project_jac(P, Xc)
