// dsc_math.cuh -- small device math for the deformable two-view hot path (sm_100a).
//
// Two arithmetic domains, as in the reference:
//   * float32 camera / triangulation maths (Modules/Calibration/*.cc, Modules/Utils/Geometry.cc),
//     written with __f*_rn intrinsics so that ptxas never contracts a*b+c into an FMA: the
//     operation order is the one fixed in oracle/f32.py and the results are bit-comparable.
//     libm calls are evaluated in double and rounded once (the correctly rounded float result
//     up to double rounding), matching oracle/f32.py:emu.
//   * float64 optimiser maths (g2o): plain double with FMA contraction allowed.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define DSC_HD __host__ __device__ __forceinline__
#define DSC_D __device__ __forceinline__

namespace dsc {

// ------------------------------------------------------------------ float32, fixed order
DSC_D float fm(float a, float b) { return __fmul_rn(a, b); }
DSC_D float fa(float a, float b) { return __fadd_rn(a, b); }
DSC_D float fs(float a, float b) { return __fsub_rn(a, b); }
DSC_D float fdv(float a, float b) { return __fdiv_rn(a, b); }
DSC_D float fsq(float a) { return __fsqrt_rn(a); }
DSC_D float emu_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
DSC_D float emu_sinf(float a) { return (float)sin((double)a); }
DSC_D float emu_cosf(float a) { return (float)cos((double)a); }

struct F3 { float x, y, z; };
DSC_D F3 mk3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
DSC_D float dot3(F3 a, F3 b) { return fa(fa(fm(a.x, b.x), fm(a.y, b.y)), fm(a.z, b.z)); }
DSC_D F3 cross3(F3 a, F3 b) {
    return mk3(fs(fm(a.y, b.z), fm(a.z, b.y)), fs(fm(a.z, b.x), fm(a.x, b.z)), fs(fm(a.x, b.y), fm(a.y, b.x)));
}
DSC_D float norm3(F3 a) { return fsq(dot3(a, a)); }
DSC_D F3 scale3(float s, F3 a) { return mk3(fm(s, a.x), fm(s, a.y), fm(s, a.z)); }
DSC_D F3 div3(F3 a, float s) { return mk3(fdv(a.x, s), fdv(a.y, s), fdv(a.z, s)); }
DSC_D F3 add3(F3 a, F3 b) { return mk3(fa(a.x, b.x), fa(a.y, b.y), fa(a.z, b.z)); }
DSC_D F3 sub3(F3 a, F3 b) { return mk3(fs(a.x, b.x), fs(a.y, b.y), fs(a.z, b.z)); }
DSC_D F3 normalize3(F3 a) { return div3(a, norm3(a)); }

// rigid transform x_c = R x + t, row-major R
struct PoseF { float R[9]; float t[3]; };
DSC_D F3 rot(const float* R, F3 v) {
    return mk3(fa(fa(fm(R[0], v.x), fm(R[1], v.y)), fm(R[2], v.z)),
               fa(fa(fm(R[3], v.x), fm(R[4], v.y)), fm(R[5], v.z)),
               fa(fa(fm(R[6], v.x), fm(R[7], v.y)), fm(R[8], v.z)));
}
DSC_D F3 rotT(const float* R, F3 v) {
    return mk3(fa(fa(fm(R[0], v.x), fm(R[3], v.y)), fm(R[6], v.z)),
               fa(fa(fm(R[1], v.x), fm(R[4], v.y)), fm(R[7], v.z)),
               fa(fa(fm(R[2], v.x), fm(R[5], v.y)), fm(R[8], v.z)));
}
DSC_D F3 apply(const PoseF& T, F3 v) {
    F3 r = rot(T.R, v);
    return mk3(fa(r.x, T.t[0]), fa(r.y, T.t[1]), fa(r.z, T.t[2]));
}
DSC_D PoseF inverse(const PoseF& T) {
    PoseF o;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = T.R[j * 3 + i];
    F3 t = rot(o.R, mk3(T.t[0], T.t[1], T.t[2]));
    o.t[0] = -t.x; o.t[1] = -t.y; o.t[2] = -t.z;
    return o;
}
DSC_D PoseF compose(const PoseF& A, const PoseF& B) {   // A * B
    PoseF o;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            o.R[i * 3 + j] = fa(fa(fm(A.R[i * 3 + 0], B.R[0 * 3 + j]), fm(A.R[i * 3 + 1], B.R[1 * 3 + j])),
                                fm(A.R[i * 3 + 2], B.R[2 * 3 + j]));
    F3 t = rot(A.R, mk3(B.t[0], B.t[1], B.t[2]));
    o.t[0] = fa(t.x, A.t[0]); o.t[1] = fa(t.y, A.t[1]); o.t[2] = fa(t.z, A.t[2]);
    return o;
}

// ------------------------------------------------------------------ cameras (float32)
struct CamF { int model; float p[8]; };

// KannalaBrandt8::project (KannalaBrandt8.cc:32-49)
DSC_D void kb8_project(const float* P, F3 X, float& u, float& v) {
    float x2y2 = fa(fm(X.x, X.x), fm(X.y, X.y));
    float theta = emu_atan2f(fsq(x2y2), X.z);
    float psi = emu_atan2f(X.y, X.x);
    float t2 = fm(theta, theta), t3 = fm(theta, t2), t5 = fm(t3, t2), t7 = fm(t5, t2), t9 = fm(t7, t2);
    float r = fa(fa(fa(fa(theta, fm(P[4], t3)), fm(P[5], t5)), fm(P[6], t7)), fm(P[7], t9));
    u = fa(fm(fm(P[0], r), emu_cosf(psi)), P[2]);
    v = fa(fm(fm(P[1], r), emu_sinf(psi)), P[3]);
}
// KannalaBrandt8::unproject (KannalaBrandt8.cc:51-83); theta_d <= 1e-8 -> optical axis (defined here,
// undefined in the reference)
DSC_D F3 kb8_unproject(const float* P, float u, float v) {
    float pwx = fdv(fs(u, P[2]), P[0]);
    float pwy = fdv(fs(v, P[3]), P[1]);
    float theta_d = fsq(fa(fm(pwx, pwx), fm(pwy, pwy)));
    if (!((double)theta_d > 1e-8)) return mk3(0.f, 0.f, 1.f);
    float theta = theta_d;
    for (int j = 0; j < 10; ++j) {
        float t2 = fm(theta, theta), t4 = fm(t2, t2), t6 = fm(t4, t2), t8 = fm(t4, t4);
        float k0 = fm(P[4], t2), k1 = fm(P[5], t4), k2 = fm(P[6], t6), k3 = fm(P[7], t8);
        float num = fs(fm(theta, fa(fa(fa(fa(1.f, k0), k1), k2), k3)), theta_d);
        float den = fa(fa(fa(fa(1.f, fm(3.f, k0)), fm(5.f, k1)), fm(7.f, k2)), fm(9.f, k3));
        float fix = fdv(num, den);
        theta = fs(theta, fix);
        if (fabsf(fix) < 1e-6f) break;
    }
    float s = emu_sinf(theta), c = emu_cosf(theta);
    return mk3(fdv(fm(s, pwx), theta_d), fdv(fm(s, pwy), theta_d), c);
}
// KannalaBrandt8::projectJac (KannalaBrandt8.cc:85-114), row-major 2x3
DSC_D void kb8_project_jac(const float* P, F3 X, float* J) {
    float x = X.x, y = X.y, z = X.z, fx = P[0], fy = P[1];
    float x2 = fm(x, x), y2 = fm(y, y), z2 = fm(z, z);
    float r2 = fa(x2, y2), r = fsq(r2), r3 = fm(r2, r);
    float theta = emu_atan2f(r, z);
    float t2 = fm(theta, theta), t3 = fm(t2, theta), t4 = fm(t2, t2), t5 = fm(t4, theta);
    float t6 = fm(t2, t4), t7 = fm(t6, theta), t8 = fm(t4, t4), t9 = fm(t8, theta);
    float f = fa(fa(fa(fa(theta, fm(t3, P[4])), fm(t5, P[5])), fm(t7, P[6])), fm(t9, P[7]));
    float fd = fa(fa(fa(fa(1.f, fm(fm(3.f, P[4]), t2)), fm(fm(5.f, P[5]), t4)), fm(fm(7.f, P[6]), t6)),
                  fm(fm(9.f, P[7]), t8));
    float den = fm(r2, fa(r2, z2));
    float fdz = fm(fd, z);
    float cross = fs(fdv(fm(fm(fdz, y), x), den), fdv(fm(fm(f, y), x), r3));
    J[0] = fm(fx, fa(fdv(fm(fdz, x2), den), fdv(fm(f, y2), r3)));
    J[1] = fm(fx, cross);
    J[2] = fdv(fm(fm(-fx, fd), x), fa(r2, z2));
    J[3] = fm(fy, cross);
    J[4] = fm(fy, fa(fdv(fm(fdz, y2), den), fdv(fm(f, x2), r3)));
    J[5] = fdv(fm(fm(-fy, fd), y), fa(r2, z2));
}
// PinHole (PinHole.cc:25-62)
DSC_D void pinhole_project(const float* P, F3 X, float& u, float& v) {
    u = fa(fdv(fm(P[0], X.x), X.z), P[2]);
    v = fa(fdv(fm(P[1], X.y), X.z), P[3]);
}
DSC_D F3 pinhole_unproject(const float* P, float u, float v) {
    return mk3(fdv(fs(u, P[2]), P[0]), fdv(fs(v, P[3]), P[1]), 1.f);
}
DSC_D void pinhole_project_jac(const float* P, F3 X, float* J) {
    float zz = fm(X.z, X.z);
    J[0] = fdv(P[0], X.z); J[1] = 0.f; J[2] = fdv(fm(-P[0], X.x), zz);
    J[3] = 0.f; J[4] = fdv(P[1], X.z); J[5] = fdv(fm(-P[1], X.y), zz);
}
DSC_D void cam_project(const CamF& c, F3 X, float& u, float& v) {
    if (c.model == 0) kb8_project(c.p, X, u, v); else pinhole_project(c.p, X, u, v);
}
DSC_D F3 cam_unproject(const CamF& c, float u, float v) {
    return c.model == 0 ? kb8_unproject(c.p, u, v) : pinhole_unproject(c.p, u, v);
}
DSC_D void cam_project_jac(const CamF& c, F3 X, float* J) {
    if (c.model == 0) kb8_project_jac(c.p, X, J); else pinhole_project_jac(c.p, X, J);
}

// ------------------------------------------------------------------ float64 helpers
struct D3 { double x, y, z; };
DSC_HD D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
DSC_HD D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
DSC_HD D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
DSC_HD D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
DSC_HD double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DSC_HD D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DSC_HD D3 mul(const double* R, D3 v) {
    return d3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
DSC_HD D3 mulT(const double* R, D3 v) {
    return d3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}

// Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w)
DSC_HD void quat_to_rot(const double* q, double* R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen's quaternion-from-matrix, then SE3Quat::normalizeRotation (w >= 0, unit norm)
DSC_HD void rot_to_quat(const double* R, double* q) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
    double s = (q[3] < 0 ? -1.0 : 1.0) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
DSC_HD void quat_mul(const double* a, const double* b, double* o) {
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
// T <- exp([omega, upsilon]) * T   (g2o::SE3Quat::exp + VertexSE3Expmap::oplusImpl), T = (q[4], t[3])
DSC_HD void se3_oplus(const double* T7, const double* upd, double* out7) {
    double wx = upd[0], wy = upd[1], wz = upd[2];
    double theta = sqrt(wx * wx + wy * wy + wz * wz);
    double Om[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double Om2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        Om2[i * 3 + j] = Om[i * 3 + 0] * Om[0 * 3 + j] + Om[i * 3 + 1] * Om[1 * 3 + j] + Om[i * 3 + 2] * Om[2 * 3 + j];
    double a, b, c, d;
    if (theta < 0.00001) { a = 1.0; b = 0.5; c = 0.5; d = 1.0 / 6.0; }
    else {
        a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta);
        c = b; d = (theta - sin(theta)) / (theta * theta * theta);
    }
    double R[9], V[9];
    for (int i = 0; i < 9; ++i) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * Om[i] + b * Om2[i];
        V[i] = I + c * Om[i] + d * Om2[i];
    }
    double dq[4];
    rot_to_quat(R, dq);
    double dt[3] = {V[0] * upd[3] + V[1] * upd[4] + V[2] * upd[5], V[3] * upd[3] + V[4] * upd[4] + V[5] * upd[5],
                    V[6] * upd[3] + V[7] * upd[4] + V[8] * upd[5]};
    double q[4];
    quat_mul(dq, T7, q);
    double s = (q[3] < 0 ? -1.0 : 1.0) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    out7[0] = q[0] * s; out7[1] = q[1] * s; out7[2] = q[2] * s; out7[3] = q[3] * s;
    double Rd[9];
    quat_to_rot(dq, Rd);
    D3 t = mul(Rd, d3(T7[4], T7[5], T7[6]));
    out7[4] = t.x + dt[0]; out7[5] = t.y + dt[1]; out7[6] = t.z + dt[2];
}

// ---- one-sided Jacobi SVD of a small NxN matrix held in registers (A = U S V^T).
// On exit the columns of A are sigma_k u_k, V holds the right singular vectors, sig[k] = |column k|,
// sorted descending.
template <int N>
DSC_HD void jacobi_svd(double (&A)[N][N], double (&V)[N][N], double (&sig)[N]) {
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < N; ++k) { alpha += A[k][p] * A[k][p]; beta += A[k][q] * A[k][q]; gamma += A[k][p] * A[k][q]; }
                if (gamma == 0.0 || fabs(gamma) <= 1e-300) continue;
                double lim = 1e-32 * alpha * beta;
                if (gamma * gamma <= lim) continue;
                off = fmax(off, gamma * gamma / (alpha * beta));
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < N; ++k) {
                    double ap = A[k][p], aq = A[k][q];
                    A[k][p] = c * ap - s * aq; A[k][q] = s * ap + c * aq;
                    double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
                }
            }
        if (off < 1e-30) break;
    }
    for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < N; ++k) s += A[k][j] * A[k][j]; sig[j] = sqrt(s); }
    for (int i = 0; i < N - 1; ++i)             // selection sort, descending
        for (int j = i + 1; j < N; ++j)
            if (sig[j] > sig[i]) {
                double ts = sig[i]; sig[i] = sig[j]; sig[j] = ts;
                for (int k = 0; k < N; ++k) {
                    double ta = A[k][i]; A[k][i] = A[k][j]; A[k][j] = ta;
                    double tv = V[k][i]; V[k][i] = V[k][j]; V[k][j] = tv;
                }
            }
}

// computeR (Geometry.cc:590-599): S = U Sig V^T, R = V U^T with the det<0 fix on the smallest singular
// direction == v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T.   S row-major in, R row-major out.
DSC_HD void rotation_from_covariance(const double* S, double* R) {
    double A[3][3], V[3][3], sig[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = S[i * 3 + j];
    jacobi_svd<3>(A, V, sig);
    if (!(sig[0] > 0.0)) {                              // S == 0: Rs stays at its identity initialisation
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    if (!(sig[1] > 1e-12 * sig[0])) {
        // rank 1 (a vertex whose only non-zero-weight neighbour is one edge): V U^T is not unique in the
        // reference (any completion of the SVD bases).  Defined here, and in oracle/graph.py, as the minimal
        // rotation taking u1 to v1.
        D3 u = (1.0 / sig[0]) * d3(A[0][0], A[1][0], A[2][0]);
        D3 v = d3(V[0][0], V[1][0], V[2][0]);
        double d = dot(u, v);
        if (d > -1.0 + 1e-12) {
            D3 c = cross(u, v);
            double k = 1.0 / (1.0 + d);
            double K[9] = {0, -c.z, c.y, c.z, 0, -c.x, -c.y, c.x, 0};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                double k2 = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
                R[i * 3 + j] = (i == j ? 1.0 : 0.0) + K[i * 3 + j] + k * k2;
            }
        } else {
            double ax = fabs(u.x), ay = fabs(u.y), az = fabs(u.z);
            D3 e = (ax <= ay && ax <= az) ? d3(1, 0, 0) : ((ay <= az) ? d3(0, 1, 0) : d3(0, 0, 1));
            D3 a = cross(u, e);
            a = (1.0 / sqrt(dot(a, a))) * a;
            double aa[3] = {a.x, a.y, a.z};
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i * 3 + j] = 2.0 * aa[i] * aa[j] - (i == j ? 1.0 : 0.0);
        }
        return;
    }
    D3 u1 = (1.0 / sig[0]) * d3(A[0][0], A[1][0], A[2][0]);
    D3 u2 = (1.0 / sig[1]) * d3(A[0][1], A[1][1], A[2][1]);
    u2 = u2 - dot(u1, u2) * u1; u2 = (1.0 / sqrt(dot(u2, u2))) * u2;
    D3 u3 = cross(u1, u2);
    D3 v1 = d3(V[0][0], V[1][0], V[2][0]), v2 = d3(V[0][1], V[1][1], V[2][1]);
    D3 v3 = cross(v1, v2);
    double vv[3][3] = {{v1.x, v2.x, v3.x}, {v1.y, v2.y, v3.y}, {v1.z, v2.z, v3.z}};
    double uu[3][3] = {{u1.x, u2.x, u3.x}, {u1.y, u2.y, u3.y}, {u1.z, u2.z, u3.z}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        R[i * 3 + j] = vv[i][0] * uu[j][0] + vv[i][1] * uu[j][1] + vv[i][2] * uu[j][2];
}

// ---- symmetric positive definite NxN inverse (Cholesky), packed upper storage helpers
// packed index of (i,j), i<=j, row-major upper triangle of an NxN
template <int N> DSC_HD int pk(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }

// in: full symmetric A (row-major NxN); out: Ainv full.  Returns false if not positive definite.
template <int N>
DSC_HD bool spd_inverse(const double* A, double* Ainv) {
    // every loop has compile-time bounds and is fully unrolled so that L and its inverse stay in registers
    double L[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) L[i][j] = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = A[j * N + j];
#pragma unroll
        for (int k = 0; k < N; ++k) if (k < j) d -= L[j][k] * L[j][k];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        d = sqrt(d);
        L[j][j] = d;
        const double inv = 1.0 / d;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i > j) {
                double s = A[i * N + j];
#pragma unroll
                for (int k = 0; k < N; ++k) if (k < j) s -= L[i][k] * L[j][k];
                L[i][j] = s * inv;
            }
        }
    }
    if (!ok) return false;
    double Li[N][N];                                  // inverse of L (lower)
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) Li[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        Li[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i > j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) if (k >= j && k < i) s -= L[i][k] * Li[k][j];
                Li[i][j] = s / L[i][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (j <= i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) if (k >= i) s += Li[k][i] * Li[k][j];
                Ainv[i * N + j] = s; Ainv[j * N + i] = s;
            }
        }
    return true;
}

}  // namespace dsc
