/* dsc_oracle.c -- plain-C restatement of the reference's non-rigid refinement (arapOptimization).
 *
 * TEST INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
 * PARITY UNPINNED: the reference cannot be built in this image (Eigen, Sophus, g2o, OpenCV, Open3D, Qhull, NLopt are
 * absent and it ships no golden vectors); this file follows the reference sources cited below and is cross-checked
 * against the independent numpy restatement in oracle/ (python modules).
 *
 * What it restates (paths relative to the reference repository):
 *   cameras        Modules/Calibration/KannalaBrandt8.cc:32-49,85-114, PinHole.cc:25-33,49-62     (float32)
 *   reprojection   Modules/Optimization/g2oTypes.h:267-298, g2oTypes.cc:270-283                   (Huber sqrt(100.991))
 *   depth edge     g2oTypes.h:390-421                                                             (numeric J upstream)
 *   ARAP edge      g2oTypes.h:300-349                                                             (numeric J upstream)
 *   set-up         Modules/Optimization/g2oBundleAdjustment.cc:608-1008 (information matrices, vertex layout)
 *   rotations      Modules/Utils/Geometry.cc:549-604 (computeR)
 *   LM             g2o OptimizationAlgorithmLevenberg as configured at g2oBundleAdjustment.cc:619-628,959-962
 *                  (g2o is not vendored; restated from its published algorithm, see oracle/lm.py)
 * The linear solve is a matrix-free block-Jacobi PCG run to a tight tolerance (the reference factorises the same
 * system with Eigen's sparse Cholesky).  `fd = 1` differentiates the depth and ARAP edges by central differences with
 * g2o's step 1e-9 as the reference does (37 energy evaluations per ARAP edge); `fd = 0` uses the analytic gradients.
 * Unknown layout: [T_g omega(3) upsilon(3) | s1 | s2 | X1_0 X2_0 | X1_1 X2_1 | ...].
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/c/Makefile); float32 expressions keep the order
 * fixed in oracle/f32.py, libm calls on floats are evaluated in double and rounded once.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int n;
    int cam_model[2];          /* 0 = KannalaBrandt8, 1 = PinHole */
    float cam[2][8];           /* fx fy cx cy k0..k3 */
    float T[2][12];            /* Tcw of KF1 / KF2, row-major 3x4 float32 (Sophus::SE3f) */
    const float* uv[2];        /* n x 2 observations */
    const double* isg[2];      /* n inverse sigma^2 of the key point octave */
    const double* d[2];        /* n depth measurements */
    const int* rowptr;         /* CSR of the directed neighbour graph */
    const int* col;
    const double* w;           /* per directed edge */
    double area;
    int ntri;
    double* R;                 /* n x 9 per-vertex rotations (row-major); filled by dso_compute_rotations */
    double* X[2];              /* n x 3 world points of KF1 / KF2, in/out */
    double Tg[7];              /* qx qy qz qw tx ty tz, in/out */
    double s[2];               /* depth scales, in/out */
} dso_problem;

typedef struct { double rep, arap, depth_sigma; } dso_weights;
typedef struct { int fd; int threads; double pcg_rtol; int pcg_max; double budget_s; /* > 0: stop after the LM iteration in which this much wall time has passed */ } dso_options;

#define FD_DELTA 1e-9

/* ------------------------------------------------------------------ small helpers */
static void quat_to_rot(const double* q, double* R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void rot_to_quat(const double* R, double* q) {          /* Eigen's trace-based conversion */
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
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
}
static void quat_normalize(double* q) {                          /* SE3Quat::normalizeRotation */
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    double nn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; ++k) q[k] /= nn;
}
static void quat_mul(const double* a, const double* b, double* o) {
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
static void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
static void mv3(const double* R, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}
static void mtv3(const double* R, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
static void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* g2o::SE3Quat::exp and VertexSE3Expmap::oplusImpl: T <- exp([omega, upsilon]) * T */
static void se3_oplus(const double* Tg7, const double* upd, double* out7) {
    const double* om = upd; const double* up = upd + 3;
    double theta = sqrt(dot3(om, om));
    double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, Om2[9], R[9], V[9];
    mat3_mul(Om, Om, Om2);
    double a, b, c, d;
    if (theta < 0.00001) { a = 1.0; b = 0.5; c = 0.5; d = 1.0 / 6.0; }
    else { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); c = b; d = (theta - sin(theta)) / (theta * theta * theta); }
    for (int k = 0; k < 9; ++k) { double I = (k % 4 == 0) ? 1.0 : 0.0; R[k] = I + a * Om[k] + b * Om2[k]; V[k] = I + c * Om[k] + d * Om2[k]; }
    double dq[4], dt[3], Rdq[9], q[4], rt[3];
    rot_to_quat(R, dq); quat_normalize(dq);
    mv3(V, up, dt);
    quat_mul(dq, Tg7, q); quat_normalize(q);
    quat_to_rot(dq, Rdq);
    mv3(Rdq, Tg7 + 4, rt);
    for (int k = 0; k < 4; ++k) out7[k] = q[k];
    for (int k = 0; k < 3; ++k) out7[4 + k] = rt[k] + dt[k];
}

/* Sophus::SE3f -> g2o::SE3Quat(unit_quaternion().cast<double>(), translation().cast<double>()) */
static void pose_from_float(const float* T34, double* R, double* t) {
    double Rf[9], q[4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Rf[i * 3 + j] = (double)T34[i * 4 + j]; t[i] = (double)T34[i * 4 + 3]; }
    rot_to_quat(Rf, q); quat_normalize(q); quat_to_rot(q, R);
}

/* ------------------------------------------------------------------ cameras (float32, fixed order) */
static float emu_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
static float emu_sqrt(float x) { return (float)sqrt((double)x); }
static void cam_project(int model, const float* P, const float* X, float* uv) {
    float x = X[0], y = X[1], z = X[2];
    if (model == 1) { uv[0] = (P[0] * x) / z + P[2]; uv[1] = (P[1] * y) / z + P[3]; return; }
    float r2 = x * x + y * y;
    float theta = emu_atan2(emu_sqrt(r2), z);
    float psi = emu_atan2(y, x);
    float t2 = theta * theta, t3 = theta * t2, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
    float r = (((theta + P[4] * t3) + P[5] * t5) + P[6] * t7) + P[7] * t9;
    uv[0] = (P[0] * r) * (float)cos((double)psi) + P[2];
    uv[1] = (P[1] * r) * (float)sin((double)psi) + P[3];
}
static void cam_project_jac(int model, const float* P, const float* X, float* J) {
    float x = X[0], y = X[1], z = X[2];
    if (model == 1) {
        J[0] = P[0] / z; J[1] = 0.f; J[2] = ((-P[0]) * x) / (z * z);
        J[3] = 0.f; J[4] = P[1] / z; J[5] = ((-P[1]) * y) / (z * z);
        return;
    }
    float fx = P[0], fy = P[1];
    float x2 = x * x, y2 = y * y, z2 = z * z, r2 = x2 + y2;
    float r = emu_sqrt(r2), r3 = r2 * r;
    float theta = emu_atan2(r, z);
    float t2 = theta * theta, t3 = t2 * theta, t4 = t2 * t2, t5 = t4 * theta, t6 = t2 * t4, t7 = t6 * theta, t8 = t4 * t4, t9 = t8 * theta;
    float f = (((theta + t3 * P[4]) + t5 * P[5]) + t7 * P[6]) + t9 * P[7];
    float fd = (((1.f + (3.f * P[4]) * t2) + (5.f * P[5]) * t4) + (7.f * P[6]) * t6) + (9.f * P[7]) * t8;
    float den = r2 * (r2 + z2);
    J[0] = fx * (((fd * z) * x2) / den + (f * y2) / r3);
    J[1] = fx * ((((fd * z) * y) * x) / den - ((f * y) * x) / r3);
    J[2] = (((-fx) * fd) * x) / (r2 + z2);
    J[3] = fy * ((((fd * z) * y) * x) / den - ((f * y) * x) / r3);
    J[4] = fy * (((fd * z) * y2) / den + (f * x2) / r3);
    J[5] = (((-fy) * fd) * y) / (r2 + z2);
}

/* ------------------------------------------------------------------ edges */
typedef struct { double R[2][9], t[2][3]; double Rg[9], tg[3]; double od, oa, huber; } ctx_t;

static void make_ctx(const dso_problem* p, const dso_weights* w, const double* Tg7, ctx_t* c) {
    for (int k = 0; k < 2; ++k) pose_from_float(p->T[k], c->R[k], c->t[k]);
    quat_to_rot(Tg7, c->Rg);
    for (int k = 0; k < 3; ++k) c->tg[k] = Tg7[4 + k];
    double ds = (double)(float)w->depth_sigma;
    c->od = 1.0 / (ds * ds);                                    /* g2oBundleAdjustment.cc:822-825 */
    c->oa = w->arap * (double)p->ntri * (double)p->ntri;        /* :946 */
    c->huber = (double)(float)sqrt(100.991);                    /* :631 */
}
/* e = obs - float(project(float(Tcw X)))  (g2oTypes.h:277-291) */
static void reproj(const dso_problem* p, const ctx_t* c, int cam, int i, const double* X, double* e, float* Xcf) {
    double Xc[3];
    mv3(c->R[cam], X, Xc);
    for (int k = 0; k < 3; ++k) Xcf[k] = (float)(Xc[k] + c->t[cam][k]);
    float uv[2];
    cam_project(p->cam_model[cam], p->cam[cam], Xcf, uv);
    e[0] = (double)p->uv[cam][2 * i] - (double)uv[0];
    e[1] = (double)p->uv[cam][2 * i + 1] - (double)uv[1];
}
static void huber(double chi2, double delta, double* rho0, double* rho1) {
    double d2 = delta * delta;
    if (chi2 <= d2) { *rho0 = chi2; *rho1 = 1.0; }
    else { double s = sqrt(chi2); *rho0 = 2 * s * delta - d2; *rho1 = delta / s; }
}
/* e = (d/s - z_c)^2, x500 if s <= 0 (g2oTypes.h:400-416) */
static double depth_energy(const ctx_t* c, int cam, const double* X, double dmeas, double s) {
    double zc = c->R[cam][6] * X[0] + c->R[cam][7] * X[1] + c->R[cam][8] * X[2] + c->t[cam][2];
    double r = dmeas / s - zc;
    double e = r * r;
    return s <= 0.0 ? e * 500 : e;
}
/* e = w (|(d2 - Ri d1)/A|^2 + |(d2 - Rj d1)/A|^2) + |Rg (X2i + X2j) - 2 t - (X1i + X1j)|^2   (g2oTypes.h:310-339) */
static double arap_energy(const double* X1i, const double* X2i, const double* X1j, const double* X2j, const double* Ri,
                          const double* Rj, double w, double area, const double* Rg, const double* tg,
                          double* a, double* b, double* g, double* qt) {
    double d1[3], d2[3], r1[3], r2[3], S2[3], rs[3];
    for (int k = 0; k < 3; ++k) { d1[k] = X1i[k] - X1j[k]; d2[k] = X2i[k] - X2j[k]; S2[k] = X2i[k] + X2j[k]; }
    mv3(Ri, d1, r1); mv3(Rj, d1, r2); mv3(Rg, S2, rs);
    for (int k = 0; k < 3; ++k) {
        a[k] = (d2[k] - r1[k]) / area; b[k] = (d2[k] - r2[k]) / area;
        qt[k] = rs[k] - 2.0 * tg[k];
        g[k] = qt[k] - (X1i[k] + X1j[k]);
    }
    return w * (dot3(a, a) + dot3(b, b)) + dot3(g, g);
}

/* activeRobustChi2 */
double dso_cost_state(const dso_problem* p, const dso_weights* w, const double* X1, const double* X2, const double* Tg7,
                      const double* s, double* parts) {
    ctx_t c; make_ctx(p, w, Tg7, &c);
    double rep = 0, dep = 0, ar = 0;
    const double* X[2] = {X1, X2};
#pragma omp parallel for reduction(+ : rep, dep, ar) schedule(static)
    for (int i = 0; i < p->n; ++i) {
        for (int cam = 0; cam < 2; ++cam) {
            double e[2], r0, r1; float xcf[3];
            reproj(p, &c, cam, i, X[cam] + 3 * i, e, xcf);
            huber(p->isg[cam][i] * w->rep * (e[0] * e[0] + e[1] * e[1]), c.huber, &r0, &r1);
            rep += r0;
            double ed = depth_energy(&c, cam, X[cam] + 3 * i, p->d[cam][i], s[cam]);
            dep += c.od * ed * ed;
        }
        double ea = 0;
        for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) {
            int j = p->col[k];
            double a[3], b[3], g[3], qt[3];
            double e = arap_energy(X1 + 3 * i, X2 + 3 * i, X1 + 3 * j, X2 + 3 * j, p->R + 9 * i, p->R + 9 * j, p->w[k], p->area,
                                   c.Rg, c.tg, a, b, g, qt);
            ea += e * e;
        }
        ar += c.oa * ea;
    }
    if (parts) { parts[0] = rep; parts[1] = dep; parts[2] = ar; }
    return rep + dep + ar;
}
double dso_cost(const dso_problem* p, const dso_weights* w, double* parts) {
    return dso_cost_state(p, w, p->X[0], p->X[1], p->Tg, p->s, parts);
}

/* ------------------------------------------------------------------ computeR (Geometry.cc:549-604) */
static void jacobi_eig3(double* A, double* V) {                 /* symmetric 3x3, cyclic Jacobi; A -> diag, V columns */
    for (int k = 0; k < 9; ++k) V[k] = (k % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[p * 3 + q];
                if (fabs(apq) < 1e-300) continue;
                double th = (A[q * 3 + q] - A[p * 3 + p]) / (2 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                double cs = 1 / sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < 3; ++k) {
                    double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = cs * akp - sn * akq; A[k * 3 + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = cs * apk - sn * aqk; A[q * 3 + k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = cs * vkp - sn * vkq; V[k * 3 + q] = sn * vkp + cs * vkq;
                }
            }
    }
}
/* R = V U^T of S = U Sigma V^T with the determinant fix; via the polar decomposition (R = the rotation closest to S^T).
 * Rank-deficient S (fewer than two independent edges) is outside what this routine defines; callers that need those
 * rows take them from the numpy oracle (the CUDA path and oracle/graph.py define the rank-1 case explicitly). */
static int rotation_from_cov(const double* S, double* R) {
    double StS[9], V[9], ev[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) StS[i * 3 + j] = S[i] * S[j] + S[3 + i] * S[3 + j] + S[6 + i] * S[6 + j];
    jacobi_eig3(StS, V);
    for (int k = 0; k < 3; ++k) ev[k] = StS[k * 4];
    /* sort descending */
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (ev[o[b]] > ev[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
    double sv[3], Vs[9], U[9];
    for (int k = 0; k < 3; ++k) { sv[k] = sqrt(fmax(ev[o[k]], 0.0)); for (int r = 0; r < 3; ++r) Vs[r * 3 + k] = V[r * 3 + o[k]]; }
    if (sv[1] <= 1e-12 * sv[0] || sv[0] == 0.0) return 0;       /* rank < 2: undefined here */
    for (int k = 0; k < 2; ++k) {                               /* u_k = S v_k / sigma_k */
        double v[3] = {Vs[k], Vs[3 + k], Vs[6 + k]}, u[3];
        mv3(S, v, u);
        for (int r = 0; r < 3; ++r) U[r * 3 + k] = u[r] / sv[k];
    }
    {   /* third columns: right-handed completion, then the reference's det fix flips the smallest direction if needed */
        double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]}, u2[3];
        cross3(u0, u1, u2);
        double v0[3] = {Vs[0], Vs[3], Vs[6]}, v1[3] = {Vs[1], Vs[4], Vs[7]}, v2[3];
        cross3(v0, v1, v2);
        for (int r = 0; r < 3; ++r) { U[r * 3 + 2] = u2[r]; Vs[r * 3 + 2] = v2[r]; }
    }
    /* R = V U^T with det(R) = +1 by construction (both completions right-handed) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = Vs[i * 3] * U[j * 3] + Vs[i * 3 + 1] * U[j * 3 + 1] + Vs[i * 3 + 2] * U[j * 3 + 2];
    return 1;
}
/* S_i = sum_j w_ij (p1i - p1j)(p2i - p2j)^T ; R_i = V U^T.  Returns the number of rank-deficient vertices (R = I there). */
int dso_compute_rotations(dso_problem* p) {
    int bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int i = 0; i < p->n; ++i) {
        double S[9] = {0};
        for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) {
            int j = p->col[k];
            double d1[3], d2[3];
            for (int c = 0; c < 3; ++c) { d1[c] = p->X[0][3 * i + c] - p->X[0][3 * j + c]; d2[c] = p->X[1][3 * i + c] - p->X[1][3 * j + c]; }
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) S[r * 3 + c] += p->w[k] * d1[r] * d2[c];
        }
        double* R = p->R + 9 * (size_t)i;
        if (!rotation_from_cov(S, R)) { for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0; ++bad; }
    }
    return bad;
}

/* ------------------------------------------------------------------ linearisation */
typedef struct {
    int n; long E;
    double* Jp;        /* n x 2 cams x (2 x 3): reprojection Jacobians */
    double* wr;        /* n x 2: information x rho' */
    double* er;        /* n x 2 x 2 residuals */
    double* Jd;        /* n x 2 x 4: depth Jacobian (X[3], s) */
    double* ed;        /* n x 2 */
    double* Ja;        /* E x 18: T(6) i1(3) i2(3) j1(3) j2(3) */
    double* ea;        /* E */
    int* src;          /* E: source vertex of every directed edge */
    int* tptr; int* tidx;  /* incoming edges of every vertex (transpose index) */
    double od, oa;
    double* b;         /* 8 + 6n */
    double* D;         /* n x 36 diagonal blocks of H (without lambda) */
    double C[64];      /* 8x8 global block */
    double chi2, maxdiag;
} lin_t;

static void lin_free(lin_t* L) {
    free(L->Jp); free(L->wr); free(L->er); free(L->Jd); free(L->ed); free(L->Ja); free(L->ea); free(L->src); free(L->tptr);
    free(L->tidx); free(L->b); free(L->D);
    memset(L, 0, sizeof(*L));
}
static int lin_alloc(lin_t* L, const dso_problem* p) {
    int n = p->n; long E = p->rowptr[n];
    memset(L, 0, sizeof(*L));
    L->n = n; L->E = E;
    L->Jp = malloc(sizeof(double) * 12 * (size_t)n); L->wr = malloc(sizeof(double) * 2 * (size_t)n);
    L->er = malloc(sizeof(double) * 4 * (size_t)n); L->Jd = malloc(sizeof(double) * 8 * (size_t)n);
    L->ed = malloc(sizeof(double) * 2 * (size_t)n); L->Ja = malloc(sizeof(double) * 18 * (size_t)(E > 0 ? E : 1));
    L->ea = malloc(sizeof(double) * (size_t)(E > 0 ? E : 1)); L->src = malloc(sizeof(int) * (size_t)(E > 0 ? E : 1));
    L->tptr = calloc((size_t)n + 1, sizeof(int)); L->tidx = malloc(sizeof(int) * (size_t)(E > 0 ? E : 1));
    L->b = malloc(sizeof(double) * (8 + 6 * (size_t)n)); L->D = malloc(sizeof(double) * 36 * (size_t)n);
    if (!L->Jp || !L->wr || !L->er || !L->Jd || !L->ed || !L->Ja || !L->ea || !L->src || !L->tptr || !L->tidx || !L->b || !L->D) return -1;
    for (int i = 0; i < n; ++i) for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) { L->src[k] = i; L->tptr[p->col[k] + 1]++; }
    for (int i = 0; i < n; ++i) L->tptr[i + 1] += L->tptr[i];
    int* cur = malloc(sizeof(int) * ((size_t)n + 1));
    if (!cur) return -1;
    memcpy(cur, L->tptr, sizeof(int) * ((size_t)n + 1));
    for (long k = 0; k < E; ++k) L->tidx[cur[p->col[k]]++] = (int)k;
    free(cur);
    return 0;
}

static void linearize(const dso_problem* p, const dso_weights* w, const double* X1, const double* X2, const double* Tg7,
                      const double* s, int fd, lin_t* L) {
    ctx_t c; make_ctx(p, w, Tg7, &c);
    L->od = c.od; L->oa = c.oa;
    const double* X[2] = {X1, X2};
    int n = p->n;
    /* perturbed global transforms of the numeric T_g columns (g2o: oplus(+d), oplus(-d) per column) */
    double Rgp[6][2][9], tgp[6][2][3];
    if (fd)
        for (int k = 0; k < 6; ++k)
            for (int sgn = 0; sgn < 2; ++sgn) {
                double upd[6] = {0}, T7[7];
                upd[k] = sgn == 0 ? FD_DELTA : -FD_DELTA;
                se3_oplus(Tg7, upd, T7);
                quat_to_rot(T7, Rgp[k][sgn]);
                for (int q = 0; q < 3; ++q) tgp[k][sgn][q] = T7[4 + q];
            }
    double chi = 0;
#pragma omp parallel for reduction(+ : chi) schedule(static)
    for (int i = 0; i < n; ++i) {
        for (int cam = 0; cam < 2; ++cam) {
            const double* Xi = X[cam] + 3 * i;
            double e[2], r0, r1; float xcf[3], Jf[6];
            reproj(p, &c, cam, i, Xi, e, xcf);
            double om = p->isg[cam][i] * w->rep;
            huber(om * (e[0] * e[0] + e[1] * e[1]), c.huber, &r0, &r1);
            chi += r0;
            cam_project_jac(p->cam_model[cam], p->cam[cam], xcf, Jf);
            double* J = L->Jp + 12 * (size_t)i + 6 * cam;             /* J = -Jproj * Rcw (g2oTypes.cc:270-283) */
            for (int r = 0; r < 2; ++r)
                for (int q = 0; q < 3; ++q)
                    J[r * 3 + q] = -((double)Jf[r * 3] * c.R[cam][q] + (double)Jf[r * 3 + 1] * c.R[cam][3 + q] + (double)Jf[r * 3 + 2] * c.R[cam][6 + q]);
            L->wr[2 * i + cam] = om * r1;
            L->er[4 * i + 2 * cam] = e[0]; L->er[4 * i + 2 * cam + 1] = e[1];
            /* depth */
            double dm = p->d[cam][i], sc = s[cam];
            double ed = depth_energy(&c, cam, Xi, dm, sc);
            double* Jd = L->Jd + 8 * (size_t)i + 4 * cam;
            if (fd) {
                for (int q = 0; q < 3; ++q) {
                    double Xp[3] = {Xi[0], Xi[1], Xi[2]}, Xm[3] = {Xi[0], Xi[1], Xi[2]};
                    Xp[q] += FD_DELTA; Xm[q] -= FD_DELTA;
                    Jd[q] = (depth_energy(&c, cam, Xp, dm, sc) - depth_energy(&c, cam, Xm, dm, sc)) / (2 * FD_DELTA);
                }
                Jd[3] = (depth_energy(&c, cam, Xi, dm, sc + FD_DELTA) - depth_energy(&c, cam, Xi, dm, sc - FD_DELTA)) / (2 * FD_DELTA);
            } else {
                double kf = sc <= 0.0 ? 500.0 : 1.0;
                double zc = c.R[cam][6] * Xi[0] + c.R[cam][7] * Xi[1] + c.R[cam][8] * Xi[2] + c.t[cam][2];
                double r = dm / sc - zc;
                for (int q = 0; q < 3; ++q) Jd[q] = 2.0 * kf * r * (-c.R[cam][6 + q]);
                Jd[3] = 2.0 * kf * r * (-dm / (sc * sc));
            }
            L->ed[2 * i + cam] = ed;
            chi += c.od * ed * ed;
        }
        double ea2 = 0;
        for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) {
            int j = p->col[k];
            const double *X1i = X1 + 3 * i, *X2i = X2 + 3 * i, *X1j = X1 + 3 * j, *X2j = X2 + 3 * j;
            const double *Ri = p->R + 9 * (size_t)i, *Rj = p->R + 9 * (size_t)j;
            double a[3], b[3], g[3], qt[3];
            double e = arap_energy(X1i, X2i, X1j, X2j, Ri, Rj, p->w[k], p->area, c.Rg, c.tg, a, b, g, qt);
            double* J = L->Ja + 18 * (size_t)k;
            if (fd) {                                                  /* g2o BaseMultiEdge::linearizeOplus */
                double t0[3], t1[3], t2[3], t3[3];
                for (int q = 0; q < 6; ++q) {
                    double ep = arap_energy(X1i, X2i, X1j, X2j, Ri, Rj, p->w[k], p->area, Rgp[q][0], tgp[q][0], t0, t1, t2, t3);
                    double em = arap_energy(X1i, X2i, X1j, X2j, Ri, Rj, p->w[k], p->area, Rgp[q][1], tgp[q][1], t0, t1, t2, t3);
                    J[q] = (ep - em) / (2 * FD_DELTA);
                }
                for (int role = 0; role < 4; ++role)
                    for (int q = 0; q < 3; ++q) {
                        double V[4][3];
                        for (int c3 = 0; c3 < 3; ++c3) { V[0][c3] = X1i[c3]; V[1][c3] = X2i[c3]; V[2][c3] = X1j[c3]; V[3][c3] = X2j[c3]; }
                        V[role][q] += FD_DELTA;
                        double ep = arap_energy(V[0], V[1], V[2], V[3], Ri, Rj, p->w[k], p->area, c.Rg, c.tg, t0, t1, t2, t3);
                        V[role][q] -= 2 * FD_DELTA;
                        double em = arap_energy(V[0], V[1], V[2], V[3], Ri, Rj, p->w[k], p->area, c.Rg, c.tg, t0, t1, t2, t3);
                        J[6 + 3 * role + q] = (ep - em) / (2 * FD_DELTA);
                    }
            } else {
                double c2 = 2.0 * p->w[k] / p->area, u[3], m[3], ra[3], rb[3], v2[3], cx[3];
                mtv3(Ri, a, ra); mtv3(Rj, b, rb); mtv3(c.Rg, g, v2); cross3(qt, g, cx);
                for (int q = 0; q < 3; ++q) {
                    u[q] = c2 * (a[q] + b[q]); m[q] = c2 * (ra[q] + rb[q]); v2[q] *= 2.0;
                    J[q] = 2.0 * cx[q]; J[3 + q] = -4.0 * g[q];
                }
                for (int q = 0; q < 3; ++q) {
                    J[6 + q] = -m[q] - 2.0 * g[q]; J[9 + q] = u[q] + v2[q]; J[12 + q] = m[q] - 2.0 * g[q]; J[15 + q] = -u[q] + v2[q];
                }
            }
            L->ea[k] = e;
            ea2 += e * e;
        }
        chi += c.oa * ea2;
    }
    L->chi2 = chi;
    /* b = -J^T W e, diagonal blocks, global block */
    memset(L->C, 0, sizeof(L->C));
    double bg[8] = {0}, C[64];
    memset(C, 0, sizeof(C));
#pragma omp parallel
    {
        double bgl[8] = {0}, Cl[64];
        memset(Cl, 0, sizeof(Cl));
#pragma omp for schedule(static) nowait
        for (int i = 0; i < n; ++i) {
            double* bi = L->b + 8 + 6 * (size_t)i;
            double* Di = L->D + 36 * (size_t)i;
            for (int k = 0; k < 6; ++k) bi[k] = 0;
            for (int k = 0; k < 36; ++k) Di[k] = 0;
            for (int cam = 0; cam < 2; ++cam) {
                const double* J = L->Jp + 12 * (size_t)i + 6 * cam;
                const double* e = L->er + 4 * i + 2 * cam;
                double wr = L->wr[2 * i + cam];
                const double* Jd = L->Jd + 8 * (size_t)i + 4 * cam;
                double ed = L->ed[2 * i + cam];
                for (int r = 0; r < 3; ++r) {
                    bi[3 * cam + r] -= wr * (J[r] * e[0] + J[3 + r] * e[1]) + L->od * Jd[r] * ed;
                    for (int q = 0; q < 3; ++q)
                        Di[(3 * cam + r) * 6 + 3 * cam + q] += wr * (J[r] * J[q] + J[3 + r] * J[3 + q]) + L->od * Jd[r] * Jd[q];
                }
                bgl[6 + cam] -= L->od * Jd[3] * ed;
                Cl[(6 + cam) * 8 + 6 + cam] += L->od * Jd[3] * Jd[3];
            }
            for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) {       /* source role */
                const double* J = L->Ja + 18 * (size_t)k;
                double we = L->oa * L->ea[k];
                for (int r = 0; r < 6; ++r) {
                    bi[r] -= we * J[6 + r];
                    for (int q = 0; q < 6; ++q) Di[r * 6 + q] += L->oa * J[6 + r] * J[6 + q];
                    bgl[r] -= we * J[r];
                    for (int q = 0; q < 6; ++q) Cl[r * 8 + q] += L->oa * J[r] * J[q];
                }
            }
            for (int t = L->tptr[i]; t < L->tptr[i + 1]; ++t) {           /* target role */
                const double* J = L->Ja + 18 * (size_t)L->tidx[t];
                double we = L->oa * L->ea[L->tidx[t]];
                for (int r = 0; r < 6; ++r) {
                    bi[r] -= we * J[12 + r];
                    for (int q = 0; q < 6; ++q) Di[r * 6 + q] += L->oa * J[12 + r] * J[12 + q];
                }
            }
        }
#pragma omp critical
        {
            for (int k = 0; k < 8; ++k) bg[k] += bgl[k];
            for (int k = 0; k < 64; ++k) C[k] += Cl[k];
        }
    }
    for (int k = 0; k < 8; ++k) L->b[k] = bg[k];
    memcpy(L->C, C, sizeof(C));
    double mx = 0;
    for (int k = 0; k < 8; ++k) mx = fmax(mx, fabs(C[k * 9]));
#pragma omp parallel for reduction(max : mx) schedule(static)
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 6; ++r) mx = fmax(mx, fabs(L->D[36 * (size_t)i + r * 7]));
    L->maxdiag = mx;
}

/* y = (H + lambda I) x with H = J^T W J applied edge by edge; se = scratch of E doubles */
static void apply_H(const dso_problem* p, const lin_t* L, double lambda, const double* x, double* y, double* se) {
    int n = p->n;
    double yg[8] = {0};
#pragma omp parallel
    {
        double ygl[8] = {0};
#pragma omp for schedule(static)
        for (long k = 0; k < L->E; ++k) {
            const double* J = L->Ja + 18 * (size_t)k;
            const double* xi = x + 8 + 6 * (size_t)L->src[k];
            const double* xj = x + 8 + 6 * (size_t)p->col[k];
            double sv = 0;
            for (int r = 0; r < 6; ++r) sv += J[r] * x[r] + J[6 + r] * xi[r] + J[12 + r] * xj[r];
            sv *= L->oa;
            se[k] = sv;
            for (int r = 0; r < 6; ++r) ygl[r] += sv * J[r];
        }
#pragma omp for schedule(static) nowait
        for (int i = 0; i < n; ++i) {
            const double* xi = x + 8 + 6 * (size_t)i;
            double* yi = y + 8 + 6 * (size_t)i;
            for (int r = 0; r < 6; ++r) yi[r] = lambda * xi[r];
            for (int cam = 0; cam < 2; ++cam) {
                const double* J = L->Jp + 12 * (size_t)i + 6 * cam;
                const double* Jd = L->Jd + 8 * (size_t)i + 4 * cam;
                const double* xc = xi + 3 * cam;
                double wr = L->wr[2 * i + cam];
                double r0 = wr * (J[0] * xc[0] + J[1] * xc[1] + J[2] * xc[2]);
                double r1 = wr * (J[3] * xc[0] + J[4] * xc[1] + J[5] * xc[2]);
                double rd = L->od * (Jd[0] * xc[0] + Jd[1] * xc[1] + Jd[2] * xc[2] + Jd[3] * x[6 + cam]);
                for (int q = 0; q < 3; ++q) yi[3 * cam + q] += J[q] * r0 + J[3 + q] * r1 + Jd[q] * rd;
                ygl[6 + cam] += Jd[3] * rd;
            }
            for (int k = p->rowptr[i]; k < p->rowptr[i + 1]; ++k) {
                const double* J = L->Ja + 18 * (size_t)k;
                for (int r = 0; r < 6; ++r) yi[r] += se[k] * J[6 + r];
            }
            for (int t = L->tptr[i]; t < L->tptr[i + 1]; ++t) {
                const double* J = L->Ja + 18 * (size_t)L->tidx[t];
                for (int r = 0; r < 6; ++r) yi[r] += se[L->tidx[t]] * J[12 + r];
            }
        }
#pragma omp critical
        for (int k = 0; k < 8; ++k) yg[k] += ygl[k];
    }
    for (int k = 0; k < 8; ++k) y[k] = yg[k] + lambda * x[k];
}

static int spd_inverse(int m, const double* A, double* Ai) {      /* Gauss-Jordan with partial pivoting, m <= 8 */
    double M[8][16];
    for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) { M[r][c] = A[r * m + c]; M[r][m + c] = r == c ? 1.0 : 0.0; }
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int r = c + 1; r < m; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (!(fabs(M[piv][c]) > 0.0) || !isfinite(M[piv][c])) return 0;
        if (piv != c) for (int k = 0; k < 2 * m; ++k) { double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
        double d = M[c][c];
        for (int k = 0; k < 2 * m; ++k) M[c][k] /= d;
        for (int r = 0; r < m; ++r) if (r != c) { double f = M[r][c]; if (f != 0.0) for (int k = 0; k < 2 * m; ++k) M[r][k] -= f * M[c][k]; }
    }
    for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) Ai[r * m + c] = M[r][m + c];
    return 1;
}

static double vdot(const double* a, const double* b, size_t m) {
    double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (long k = 0; k < (long)m; ++k) s += a[k] * b[k];
    return s;
}

/* block-Jacobi PCG, x0 = 0, stop when sqrt(|r.z| / r0.z0) <= rtol (same criterion as oracle/lm.py and the CUDA path) */
static int solve_pcg(const dso_problem* p, const lin_t* L, double lambda, double rtol, int max_iter, double* x, int* iters,
                     double* work /* 4 x (8 + 6n) + E */) {
    size_t m = 8 + 6 * (size_t)p->n;
    double *r = work, *z = work + m, *pp = work + 2 * m, *Ap = work + 3 * m, *se = work + 4 * m;
    double* Mi = malloc(sizeof(double) * 36 * (size_t)p->n);
    double G[64], Gi[64];
    if (!Mi) return -1;
    int bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int i = 0; i < p->n; ++i) {
        double A[36];
        for (int k = 0; k < 36; ++k) A[k] = L->D[36 * (size_t)i + k] + ((k % 7 == 0) ? lambda : 0.0);
        if (!spd_inverse(6, A, Mi + 36 * (size_t)i)) ++bad;
    }
    for (int k = 0; k < 64; ++k) G[k] = L->C[k] + ((k % 9 == 0) ? lambda : 0.0);
    if (bad || !spd_inverse(8, G, Gi)) { free(Mi); return -1; }
#define PRECOND(src, dst)                                                                                 \
    do {                                                                                                  \
        for (int a_ = 0; a_ < 8; ++a_) { double s_ = 0; for (int c_ = 0; c_ < 8; ++c_) s_ += Gi[a_ * 8 + c_] * (src)[c_]; (dst)[a_] = s_; } \
        _Pragma("omp parallel for schedule(static)")                                                      \
        for (int i_ = 0; i_ < p->n; ++i_)                                                                 \
            for (int a_ = 0; a_ < 6; ++a_) {                                                              \
                double s_ = 0;                                                                            \
                for (int c_ = 0; c_ < 6; ++c_) s_ += Mi[36 * (size_t)i_ + a_ * 6 + c_] * (src)[8 + 6 * (size_t)i_ + c_];       \
                (dst)[8 + 6 * (size_t)i_ + a_] = s_;                                                      \
            }                                                                                             \
    } while (0)
    memset(x, 0, sizeof(double) * m);
    memcpy(r, L->b, sizeof(double) * m);
    PRECOND(r, z);
    memcpy(pp, z, sizeof(double) * m);
    double rz = vdot(r, z, m), rz0 = rz;
    int it = 0, ok = 1;
    if (rz0 > 0)
        for (it = 1; it <= max_iter; ++it) {
            apply_H(p, L, lambda, pp, Ap, se);
            double alpha = rz / vdot(pp, Ap, m);
#pragma omp parallel for schedule(static)
            for (long k = 0; k < (long)m; ++k) { x[k] += alpha * pp[k]; r[k] -= alpha * Ap[k]; }
            PRECOND(r, z);
            double rzn = vdot(r, z, m);
            if (!isfinite(rzn)) { ok = 0; break; }
            if (sqrt(fabs(rzn) / rz0) <= rtol) { rz = rzn; break; }
            double beta = rzn / rz;
#pragma omp parallel for schedule(static)
            for (long k = 0; k < (long)m; ++k) pp[k] = z[k] + beta * pp[k];
            rz = rzn;
        }
#undef PRECOND
    free(Mi);
    if (iters) *iters = it > max_iter ? max_iter : it;
    if (!ok) return -1;
    for (size_t k = 0; k < m; ++k) if (!isfinite(x[k])) return -1;
    return 0;
}

/* g2o SparseOptimizer::optimize(iters) with OptimizationAlgorithmLevenberg.  chi2[iters + 1]: cost at the start of
 * every iteration and the final cost at [n_done]; lam[iters]; trials[iters]; pcg_its[iters] (sum over the trials). */
int dso_optimize(dso_problem* p, const dso_weights* w, const dso_options* opt, int iters, double* chi2, double* lam_out,
                 int* trials, int* pcg_its, int* n_done) {
#ifdef _OPENMP
    omp_set_num_threads(opt->threads > 0 ? opt->threads : omp_get_num_procs());
#endif
    int n = p->n;
    size_t m = 8 + 6 * (size_t)n;
    lin_t L;
    if (lin_alloc(&L, p)) { lin_free(&L); return -1; }
    double* work = malloc(sizeof(double) * (4 * m + (size_t)(L.E > 0 ? L.E : 1)));
    double* dx = malloc(sizeof(double) * m);
    double* T1 = malloc(sizeof(double) * 3 * (size_t)n);
    double* T2 = malloc(sizeof(double) * 3 * (size_t)n);
    if (!work || !dx || !T1 || !T2) { free(work); free(dx); free(T1); free(T2); lin_free(&L); return -1; }
    double lambda = 0, ni = 2;
    int it, done = 0;
#ifdef _OPENMP
    const double t_start = omp_get_wtime();
#endif
    for (it = 0; it < iters; ++it) {
#ifdef _OPENMP
        if (opt->budget_s > 0 && it > 0 && omp_get_wtime() - t_start > opt->budget_s) break;
#endif
        linearize(p, w, p->X[0], p->X[1], p->Tg, p->s, opt->fd, &L);
        double current = dso_cost_state(p, w, p->X[0], p->X[1], p->Tg, p->s, NULL);
        if (it == 0) { lambda = 1e-5 * L.maxdiag; ni = 2; }             /* computeLambdaInit, tau = 1e-5 */
        chi2[it] = current; lam_out[it] = lambda; pcg_its[it] = 0;
        double rho = 0;
        int q = 0;
        do {
            int its = 0;
            double temp = DBL_MAX, scale = 1e-3, Tg7[7], s2[2];
            int rc = solve_pcg(p, &L, lambda, opt->pcg_rtol, opt->pcg_max, dx, &its, work);
            pcg_its[it] += its;
            if (rc == 0) {
                se3_oplus(p->Tg, dx, Tg7);
                s2[0] = p->s[0] + dx[6]; s2[1] = p->s[1] + dx[7];
#pragma omp parallel for schedule(static)
                for (int i = 0; i < n; ++i)
                    for (int k = 0; k < 3; ++k) {
                        T1[3 * i + k] = p->X[0][3 * i + k] + dx[8 + 6 * (size_t)i + k];
                        T2[3 * i + k] = p->X[1][3 * i + k] + dx[8 + 6 * (size_t)i + 3 + k];
                    }
                temp = dso_cost_state(p, w, T1, T2, Tg7, s2, NULL);
                double sc = 0;
                for (size_t k = 0; k < m; ++k) sc += dx[k] * (lambda * dx[k] + L.b[k]);
                scale = sc + 1e-3;
            }
            rho = (current - temp) / scale;
            if (rho > 0 && isfinite(temp)) {
                double alpha = 1.0 - pow(2 * rho - 1, 3);
                if (alpha > 2.0 / 3.0) alpha = 2.0 / 3.0;
                lambda *= alpha > 1.0 / 3.0 ? alpha : 1.0 / 3.0;
                ni = 2;
                current = temp;
                memcpy(p->X[0], T1, sizeof(double) * 3 * (size_t)n);
                memcpy(p->X[1], T2, sizeof(double) * 3 * (size_t)n);
                memcpy(p->Tg, Tg7, sizeof(Tg7));
                p->s[0] = s2[0]; p->s[1] = s2[1];
            } else {
                lambda *= ni;
                ni *= 2;
            }
            ++q;
        } while (rho < 0 && q < 10);
        trials[it] = q;
        done = it + 1;
        chi2[done] = current;
        if (q == 10 || rho == 0) break;
    }
    if (done == 0) chi2[0] = dso_cost_state(p, w, p->X[0], p->X[1], p->Tg, p->s, NULL);
    *n_done = done;
    free(work); free(dx); free(T1); free(T2);
    lin_free(&L);
    return 0;
}

/* gradient b, diagonal of H and y = (H + lambda I) x at the current state: lets the tests compare the C linearisation
 * with the numpy oracle's assembled sparse matrices */
int dso_debug_linearize(dso_problem* p, const dso_weights* w, int fd, double lambda, const double* x, double* b, double* hdiag,
                        double* y, double* chi2) {
    lin_t L;
    if (lin_alloc(&L, p)) { lin_free(&L); return -1; }
    linearize(p, w, p->X[0], p->X[1], p->Tg, p->s, fd, &L);
    size_t m = 8 + 6 * (size_t)p->n;
    if (b) memcpy(b, L.b, sizeof(double) * m);
    if (hdiag) {
        for (int k = 0; k < 8; ++k) hdiag[k] = L.C[k * 9];
        for (int i = 0; i < p->n; ++i) for (int r = 0; r < 6; ++r) hdiag[8 + 6 * (size_t)i + r] = L.D[36 * (size_t)i + r * 7];
    }
    if (x && y) {
        double* se = malloc(sizeof(double) * (size_t)(L.E > 0 ? L.E : 1));
        if (!se) { lin_free(&L); return -1; }
        apply_H(p, &L, lambda, x, y, se);
        free(se);
    }
    if (chi2) *chi2 = L.chi2;
    lin_free(&L);
    return 0;
}

/* Timing aid of bench.py's CPU baseline (wall clock, omp_get_wtime): out[0] = one linearisation + cost evaluation
 * (what starts every LM iteration), out[1] = one PCG iteration (a solve at lambda = 1e-5 max diag H cut after pcg_reps
 * iterations, preconditioner set-up excluded by differencing two lengths), out[2] = one cost evaluation (one per LM
 * trial), out[3] = the block-Jacobi set-up of a solve.  The state of p is not changed. */
int dso_time_components(dso_problem* p, const dso_weights* w, int threads, int pcg_reps, double* out) {
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
    lin_t L;
    if (lin_alloc(&L, p)) { lin_free(&L); return -1; }
    size_t m = 8 + 6 * (size_t)p->n;
    double* work = malloc(sizeof(double) * (4 * m + (size_t)(L.E > 0 ? L.E : 1)));
    double* dx = malloc(sizeof(double) * m);
    if (!work || !dx) { free(work); free(dx); lin_free(&L); return -1; }
    double t0 = omp_get_wtime();
    linearize(p, w, p->X[0], p->X[1], p->Tg, p->s, 0, &L);
    double c = dso_cost_state(p, w, p->X[0], p->X[1], p->Tg, p->s, NULL);
    out[0] = omp_get_wtime() - t0;
    t0 = omp_get_wtime();
    c += dso_cost_state(p, w, p->X[0], p->X[1], p->Tg, p->s, NULL);
    out[2] = omp_get_wtime() - t0;
    double lambda = 1e-5 * L.maxdiag;
    int its = 0;
    t0 = omp_get_wtime();
    solve_pcg(p, &L, lambda, 0.0, 1, dx, &its, work);
    double t1 = omp_get_wtime() - t0;
    t0 = omp_get_wtime();
    solve_pcg(p, &L, lambda, 0.0, 1 + pcg_reps, dx, &its, work);
    double t2 = omp_get_wtime() - t0;
    out[1] = (t2 - t1) / (double)pcg_reps;
    out[3] = t1 - out[1];
    free(work); free(dx);
    lin_free(&L);
    return c == c ? 0 : 1;
#else
    (void)p; (void)w; (void)threads; (void)pcg_reps; (void)out;
    return -2;
#endif
}

int dso_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
