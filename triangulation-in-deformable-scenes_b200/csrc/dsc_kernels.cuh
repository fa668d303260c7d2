// dsc_kernels.cuh -- hand-written sm_100a kernels of the deformable two-view hot path.
//
// Layout in HBM (n correspondences in the library's internal order: Morton order of KF1's world (x, y), then a stable
// degree sort inside groups of kSortGroup rows; E directed neighbour edges as a sliced ELL):
//   P      double[2][n][4] {X1.xyz, 0} plane then {X2.xyz, 0} plane: 32 B records, one neighbour gather of a
//                        point = 1 aligned sector (LDG.E.256); the PCG operator only touches the X1 plane
//   Q      double[n][4]  per-vertex ARAP rotation as unit quaternion (computeR)   32 B = 1 sector
//   uv     float4[n]     {u1, v1, u2, v2}
//   dm     double2[n]    depth measurements (KF1, KF2)
//   isg    float2[n]     KeyFrame::getInvSigma2(octave) of the two observations
//   sliced ELL: slice = 32 consecutive rows = one warp, lane = row; block b = column k of a slice (b = sliceptr[slice] + k)
//   ecol   int[nblk][32]         neighbour of the lane's row (own tile first, halo last; padding = the row itself)
//   ewgt   double[nblk][32]      edge weight (0 on padding)
//   Je     double[nblk][9][32]   ARAP Jacobian record {u, m, g} of the directed edge, written by the linearisation
//   U      double[n/32][14][32]  unary Hessian record {U1[6], U2[6], kd1, kd2}, slice-major
//   D      double[n/32][21][32]  packed upper 6x6 diagonal block of H, slice-major (block-Jacobi preconditioner source)
//   Minv   double[n/32][21][32]  packed inverse of D + lambda I, slice-major
//   vectors b, x, r, z, w, p, s: double[n][6] {d/dX1, d/dX2}; the 8 global unknowns
//   (T_g omega/upsilon, s1, s2) live in separate 8-vectors.
//
// Work decomposition: one warp per 32-row slice, lane = row, the row's sums stay in registers (no cross-lane reduction,
// no global atomics anywhere); a block owns a tile of kSortGroup rows whose points / vectors are staged in shared memory
// (dsc_kernels_ell.cuh for the gather kernels, cg_spmv_kernel below for the PCG operator).  The neighbour graph is
// symmetric and the reference adds one EdgeARAP per *directed* pair with identical residual
// (g2oBundleAdjustment.cc:883-953), so vertex i's row of J^T W J is 2x the sum over its own row.
// Grids are persistent (multiples of the SM count); every reduction is two-stage and deterministic (per-block
// partials, summed in a fixed order by the consumer).
#pragma once
#include "dsc_math.cuh"
#include "dsc_shard.cuh"

namespace dsc {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 2048;
constexpr int kSortGroup = 512;                 // rows are degree-sorted inside groups of this many (host, dsc_set_graph)

struct PairDev {
    CamF cam1, cam2;
    PoseF T1f, T2f;
    double R1[9], t1[3], R2[9], t2[3];          // g2o::SE3Quat(unit_quaternion().cast<double>(), ...)
};

struct WeightsDev {
    double rep;          // repBalanceWeight
    double arap_info;    // arapBalanceWeight * n_triangles^2   (g2oBundleAdjustment.cc:946)
    double depth_info;   // 1 / DepthError^2                    (:822-825)
    double inv_area;     // 1 / mesh area                       (:942-943, g2oTypes.h:333-334)
    double huber;        // deltaMono = (float)sqrt(100.991)    (:631)
};

struct Globals {         // the 8 global unknowns + derived rotation
    double Tg[7];        // qx qy qz qw tx ty tz
    double s1, s2;
    double Rg[9];
};

struct LinGlobal {       // reduced by finalize_linearize
    double chi2[3];      // reprojection (robust), depth, ARAP
    double maxdiag;
    double bg[8];        // gradient of the global unknowns (b = -J^T W e)
    double C[64];        // 8x8 global block of H
};

struct CgScalars { double gamma_prev, alpha_prev; };
struct CgControl {
    CgScalars sc[2];
    double gamma0;
    double lambda;       // damping of the running solve  } read by the iteration kernels from device memory so that a
    double rtol2;        // squared stopping tolerance     } captured CUDA graph of iterations can be replayed unchanged
    int iters;
    int converged;
    int breakdown;
    double gamma_true;   // fp32 mode: r.M^-1 r of the true (double) residual at the last restart (refine_gamma_kernel)
};

// ------------------------------------------------------------------ helpers
DSC_D D3 qrot(const double* q, D3 v) {              // Eigen::Quaternion::_transformVector
    D3 qv = d3(q[0], q[1], q[2]);
    D3 t = 2.0 * cross(qv, v);
    return v + q[3] * t + cross(qv, t);
}
DSC_D D3 qrotT(const double* q, D3 v) {
    D3 qv = d3(-q[0], -q[1], -q[2]);
    D3 t = 2.0 * cross(qv, v);
    return v + q[3] * t + cross(qv, t);
}
DSC_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// deterministic block reduction of K values held per thread; result valid in thread 0
template <int K>
DSC_D void block_reduce(double (&v)[K], double* smem /* [K][kThreads/32] */) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) smem[k * (kThreads / 32) + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) s += smem[k * (kThreads / 32) + w];
            v[k] = s;
        }
    }
    __syncthreads();
}
// every thread of the block gets the fixed-order sum of part[0..nb) (stride `stride`)
DSC_D double sum_partials(const double* part, int nb, int stride, double* smem32) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s += part[(size_t)i * stride];
    s = warp_sum(s);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem32[warp] = s;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += smem32[w];
    return t;
}

// 256-bit read-only global load (sm_100: ld.global.nc.v4.f64 -> one LDG.E.256; p must be 32-byte aligned)
DSC_D double4 ldg256(const double4* p) {
    double4 r;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
// Load policy of the kernels that exist in two forms.  kRO = true: the data is constant for the whole launch (the
// per-phase kernels of the 1M path): non-coherent read-only loads.  kRO = false: other CTAs of the SAME launch write the
// data between cluster barriers (the one-launch LM of the batched path, dsc_batch.cuh): loads that are coherent at L2
// (ld.global.cg -> LDG.E.ENL2.256.STRONG.GPU), never served from a stale L1 / read-only line.
template <bool kRO>
DSC_D double4 ld256(const double4* p) {
    if (kRO) return ldg256(p);
    double4 r;
    asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
}
template <bool kRO>
DSC_D double ld64(const double* p) { return kRO ? __ldg(p) : __ldcg(p); }
struct P8 { D3 a, b; };                             // X1, X2 of one correspondence
template <bool kRO = true>
DSC_D P8 load_P(const double* __restrict__ P, int n, int i) {
    const double4* p = reinterpret_cast<const double4*>(P);
    double4 u = ld256<kRO>(p + (size_t)i), v = ld256<kRO>(p + (size_t)n + (size_t)i);
    P8 r; r.a = d3(u.x, u.y, u.z); r.b = d3(v.x, v.y, v.z);
    return r;
}
template <bool kRO>
DSC_D void load6_t(const double* V, int i, D3& a, D3& b) {
    const double2* p = reinterpret_cast<const double2*>(V) + 3 * (size_t)i;
    double2 u, v, w;
    if (kRO) { u = p[0]; v = p[1]; w = p[2]; }
    else { u = __ldcg(p); v = __ldcg(p + 1); w = __ldcg(p + 2); }
    a = d3(u.x, u.y, v.x); b = d3(v.y, w.x, w.y);
}
DSC_D void load6(const double* __restrict__ V, int i, D3& a, D3& b) {
    const double2* p = reinterpret_cast<const double2*>(V) + 3 * (size_t)i;
    double2 u = p[0], v = p[1], w = p[2];
    a = d3(u.x, u.y, v.x); b = d3(v.y, w.x, w.y);
}
DSC_D void store6(double* __restrict__ V, int i, D3 a, D3 b) {
    double2* p = reinterpret_cast<double2*>(V) + 3 * (size_t)i;
    p[0] = make_double2(a.x, a.y); p[1] = make_double2(a.z, b.x); p[2] = make_double2(b.y, b.z);
}
// float storage of the solver vectors (dsc_set_precision, DSC_PRECISION_F32): the same records in 24 bytes; every
// value is widened on load and all arithmetic (sums, dot products, the 8x8 global block) stays double
template <bool kRO>
DSC_D void load6_t(const float* V, int i, D3& a, D3& b) {
    const float2* p = reinterpret_cast<const float2*>(V) + 3 * (size_t)i;
    float2 u, v, w;
    if (kRO) { u = p[0]; v = p[1]; w = p[2]; }
    else { u = __ldcg(p); v = __ldcg(p + 1); w = __ldcg(p + 2); }
    a = d3((double)u.x, (double)u.y, (double)v.x); b = d3((double)v.y, (double)w.x, (double)w.y);
}
DSC_D void load6(const float* __restrict__ V, int i, D3& a, D3& b) { load6_t<true>(V, i, a, b); }
DSC_D void store6(float* __restrict__ V, int i, D3 a, D3 b) {
    float2* p = reinterpret_cast<float2*>(V) + 3 * (size_t)i;
    p[0] = make_float2((float)a.x, (float)a.y); p[1] = make_float2((float)a.z, (float)b.x); p[2] = make_float2((float)b.y, (float)b.z);
}
DSC_D void load_q(const double* __restrict__ Q, int i, double* q) {
    const double4* p = reinterpret_cast<const double4*>(Q) + (size_t)i;
    double4 u = ldg256(p);
    q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w;
}

// One directed ARAP edge (i -> j).  EdgeARAP::computeError, g2oTypes.h:310-339:
//   e = w (|(d2 - Ri d1)/A|^2 + |(-d2 + Rj d1)/A|^2) + |Rg (X2i + X2j) - 2 t - (X1i + X1j)|^2
// and its analytic gradient (the reference differentiates numerically, g2o central differences).
struct ArapGrad { D3 gi1, gi2, gj1, gj2, gw, gv; D3 u, m, g; double e; };
template <bool kGrad>
DSC_D void arap_edge(const P8& Pi, const P8& Pj, const double* qi, const double* qj, double w, double inv_area,
                     const Globals& G, ArapGrad& o) {
    D3 d1 = Pi.a - Pj.a, d2 = Pi.b - Pj.b;
    D3 a = inv_area * (d2 - qrot(qi, d1));
    D3 b = inv_area * (d2 - qrot(qj, d1));
    D3 S2 = Pi.b + Pj.b, S1 = Pi.a + Pj.a;
    D3 tg = d3(G.Tg[4], G.Tg[5], G.Tg[6]);
    D3 qt = mul(G.Rg, S2) - 2.0 * tg;
    D3 g = qt - S1;
    o.e = w * (dot(a, a) + dot(b, b)) + dot(g, g);
    if (kGrad) {
        double c2 = 2.0 * w * inv_area;
        D3 u = c2 * (a + b);
        D3 m = c2 * (qrotT(qi, a) + qrotT(qj, b));
        D3 v2 = 2.0 * mulT(G.Rg, g);
        D3 g2 = 2.0 * g;
        o.gi1 = d3(-m.x - g2.x, -m.y - g2.y, -m.z - g2.z);
        o.gi2 = u + v2;
        o.gj1 = m - g2;
        o.gj2 = v2 - u;
        o.gw = 2.0 * cross(qt, g);
        o.gv = d3(-4.0 * g.x, -4.0 * g.y, -4.0 * g.z);
        o.u = u; o.m = m; o.g = g;
    }
}

// Reprojection edge EdgeSE3ProjectXYZPerKeyFrameOnlyPoints (g2oTypes.h:277-291, g2oTypes.cc:270-283):
// e = obs - (double)project((float)(R X + t)),  J = -(double)projectJac((float)Xc) R.
DSC_D void reproj_residual(const CamF& cam, const double* R, const double* t, D3 X, float u, float v,
                           double& e0, double& e1, F3& Xcf) {
    D3 Xc = mul(R, X);
    Xc = d3(Xc.x + t[0], Xc.y + t[1], Xc.z + t[2]);
    Xcf = mk3((float)Xc.x, (float)Xc.y, (float)Xc.z);
    float pu, pv;
    cam_project(cam, Xcf, pu, pv);
    e0 = (double)u - (double)pu;
    e1 = (double)v - (double)pv;
}
DSC_D void huber(double chi2, double delta, double& rho0, double& rho1) {   // g2o::RobustKernelHuber
    double d2 = delta * delta;
    if (chi2 <= d2) { rho0 = chi2; rho1 = 1.0; }
    else { double s = sqrt(chi2); rho0 = 2.0 * s * delta - d2; rho1 = delta / s; }
}

// ================================================================== K1: triangulation
struct TriParams { int method, location, gate; float min_cos, depth_limit; int check_reproj; };

DSC_D void tri_nrslam(F3 xn1, F3 xn2, const PoseF& T21, int loc, F3& p1, F3& p2) {   // Geometry.cc:103-153
    F3 f0 = normalize3(xn1), f1 = normalize3(xn2);
    F3 t = mk3(T21.t[0], T21.t[1], T21.t[2]);
    F3 Rf0 = rot(T21.R, f0);
    F3 p = cross3(Rf0, f1), q = cross3(Rf0, t), r = cross3(f1, t);
    float np_ = norm3(p), nq = norm3(q), nr = norm3(r);
    float l0 = fdv(nr, np_), l1 = fdv(nq, np_);
    F3 point0 = scale3(l0, Rf0), point1 = scale3(l1, f1);
    F3 x1 = scale3(fdv(nq, fa(nq, nr)), add3(t, scale3(l0, add3(Rf0, f1))));
    if (loc == 1) { p1 = x1; p2 = x1; }
    else if (loc == 2) {
        point0 = add3(t, point0);
        p1 = add3(point0, sub3(point0, x1));
        p2 = add3(point1, sub3(point1, x1));
    } else { p1 = add3(t, point0); p2 = point1; }
}
DSC_D void tri_classic(F3 xn1, F3 xn2, const PoseF& T21, int loc, F3& p1, F3& p2) {  // Geometry.cc:62-101
    F3 m0 = rot(T21.R, xn1), m1 = xn2;
    F3 tt = mk3(T21.t[0], T21.t[1], T21.t[2]);
    F3 t = normalize3(tt), M0 = normalize3(m0), M1 = normalize3(m1);
    F3 a0 = sub3(M0, scale3(dot3(M0, t), t)), a1 = sub3(M1, scale3(dot3(M1, t), t));
    // V.col(1) of the 2x3 SVD: right singular vector of the smaller singular value (double, rounded)
    double g00 = (double)a0.x * a0.x + (double)a0.y * a0.y + (double)a0.z * a0.z;
    double g11 = (double)a1.x * a1.x + (double)a1.y * a1.y + (double)a1.z * a1.z;
    double g01 = (double)a0.x * a1.x + (double)a0.y * a1.y + (double)a0.z * a1.z;
    double th = 0.5 * atan2(2.0 * g01, g00 - g11);
    double ux = -sin(th), uy = cos(th);
    double nx = ux * a0.x + uy * a1.x, ny = ux * a0.y + uy * a1.y, nz = ux * a0.z + uy * a1.z;
    double nn = sqrt(nx * nx + ny * ny + nz * nz);
    F3 n = mk3((float)(nx / nn), (float)(ny / nn), (float)(nz / nn));
    F3 m0_ = sub3(m0, scale3(dot3(m0, n), n)), m1_ = sub3(m1, scale3(dot3(m1, n), n));
    F3 z = cross3(m1_, m0_);
    float zz = dot3(z, z);
    float l0 = fdv(dot3(z, cross3(tt, m1_)), zz), l1 = fdv(dot3(z, cross3(tt, m0_)), zz);
    if (loc == 1) { p1 = add3(tt, scale3(l0, m0_)); p2 = p1; }
    else { p1 = add3(tt, scale3(l0, m0)); p2 = scale3(l1, m1); }
}
DSC_D void tri_dlt(F3 xn1, F3 xn2, const PoseF& T1, const PoseF& T2, F3& X) {         // Geometry.cc:155-186 (intended)
    double A[4][4], V[4][4], sig[4];
    for (int k = 0; k < 4; ++k) {
        float r10 = k < 3 ? T1.R[0 + k] : T1.t[0], r11 = k < 3 ? T1.R[3 + k] : T1.t[1], r12 = k < 3 ? T1.R[6 + k] : T1.t[2];
        float r20 = k < 3 ? T2.R[0 + k] : T2.t[0], r21 = k < 3 ? T2.R[3 + k] : T2.t[1], r22 = k < 3 ? T2.R[6 + k] : T2.t[2];
        A[0][k] = (double)fs(fm(xn1.x, r12), r10);
        A[1][k] = (double)fs(fm(xn1.y, r12), r11);
        A[2][k] = (double)fs(fm(xn2.x, r22), r20);
        A[3][k] = (double)fs(fm(xn2.y, r22), r21);
    }
    jacobi_svd<4>(A, V, sig);
    float x0 = (float)V[0][3], x1 = (float)V[1][3], x2 = (float)V[2][3], x3 = (float)V[3][3];
    if (x3 != 0.f) X = mk3(fdv(x0, x3), fdv(x1, x3), fdv(x2, x3)); else X = mk3(0.f, 0.f, 0.f);
}

// useTriangulationMethod (Geometry.cc:216-230) for one match: xn1 / xn2 are the (normalised) bearing rays of the two
// key points; DepthMeasurement takes the camera-frame points at the measured depth instead (cp1 / cp2).  World points out.
DSC_D void triangulate_match(const PairDev& pr, int method, int location, F3 xn1, F3 xn2, F3 cp1, F3 cp2, F3& w1, F3& w2) {
    if (method == 2) {
        F3 X; tri_dlt(xn1, xn2, pr.T1f, pr.T2f, X);
        w1 = X; w2 = X;
        return;
    }
    const PoseF T21 = compose(pr.T2f, inverse(pr.T1f));
    const PoseF Tw2 = inverse(pr.T2f);
    F3 p1, p2;
    if (method == 3) {                               // Geometry.cc:189-214
        F3 point0 = apply(T21, cp1), point1 = cp2;
        F3 xm = div3(add3(point0, point1), 2.0f);
        if (location == 1) { p1 = xm; p2 = xm; }
        else if (location == 2) { p1 = add3(point0, sub3(point0, xm)); p2 = add3(point1, sub3(point1, xm)); }
        else { p1 = point0; p2 = point1; }
    } else if (method == 0) tri_classic(xn1, xn2, T21, location, p1, p2);
    else tri_nrslam(xn1, xn2, T21, location, p1, p2);
    w1 = apply(Tw2, p1); w2 = apply(Tw2, p2);
}

__global__ void __launch_bounds__(kThreads)
triangulate_kernel(int n, const float2* __restrict__ uv1, const float2* __restrict__ uv2,
                   const float* __restrict__ dep1, const float* __restrict__ dep2,
                   const __grid_constant__ PairDev pr, const __grid_constant__ TriParams tp,
                   float* __restrict__ X1, float* __restrict__ X2, uint8_t* __restrict__ valid,
                   float* __restrict__ cosp) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float2 a = uv1[i], b = uv2[i];
        F3 ray1 = cam_unproject(pr.cam1, a.x, a.y), ray2 = cam_unproject(pr.cam2, b.x, b.y);
        F3 xn1 = normalize3(ray1), xn2 = normalize3(ray2);
        F3 cp1 = ray1, cp2 = ray2;
        if (tp.method == 3) {
            float s1 = fdv(dep1[i], ray1.z), s2 = fdv(dep2[i], ray2.z);
            cp1 = mk3(fm(ray1.x, s1), fm(ray1.y, s1), fm(ray1.z, s1));
            cp2 = mk3(fm(ray2.x, s2), fm(ray2.y, s2), fm(ray2.z, s2));
        }
        F3 w1, w2;
        triangulate_match(pr, tp.method, tp.location, xn1, xn2, cp1, cp2, w1, w2);
        F3 c1 = apply(pr.T1f, w1), c2 = apply(pr.T2f, w2);
        F3 r1 = normalize3(rotT(pr.T1f.R, xn1)), r2 = normalize3(rotT(pr.T2f.R, xn2));
        float cp = fdv(dot3(r1, r2), fm(norm3(r1), norm3(r2)));
        bool ok = true;
        if (tp.gate == 1) {                              // Mapping::isValidParallax, Mapping.cc:351-364
            if (c1.z < 0.f || c2.z < 0.f) ok = false;
            if (!(cp <= tp.min_cos)) ok = false;
        } else if (tp.gate == 2) {                       // MonocularMapInitializer.cc:315-360
            bool fin = isfinite(w1.x) && isfinite(w1.y) && isfinite(w1.z) && isfinite(w2.x) && isfinite(w2.y) && isfinite(w2.z);
            bool z1 = (w1.x == 0.f && w1.y == 0.f && w1.z == 0.f), z2 = (w2.x == 0.f && w2.y == 0.f && w2.z == 0.f);
            if (!fin || z1 || z2) ok = false;
            if (c1.z < 0.f || c1.z > tp.depth_limit) ok = false;
            if (c2.z < 0.f || c2.z > tp.depth_limit) ok = false;
            if (tp.check_reproj) {
                float pu, pv;
                cam_project(pr.cam1, c1, pu, pv);
                float ex = fs(a.x, pu), ey = fs(a.y, pv);
                if ((double)fa(fm(ex, ex), fm(ey, ey)) > 5.991) ok = false;
                cam_project(pr.cam2, c2, pu, pv);
                ex = fs(b.x, pu); ey = fs(b.y, pv);
                if ((double)fa(fm(ex, ex), fm(ey, ey)) > 5.991) ok = false;
            }
        }
        X1[3 * (size_t)i + 0] = w1.x; X1[3 * (size_t)i + 1] = w1.y; X1[3 * (size_t)i + 2] = w1.z;
        X2[3 * (size_t)i + 0] = w2.x; X2[3 * (size_t)i + 1] = w2.y; X2[3 * (size_t)i + 2] = w2.z;
        valid[i] = ok ? 1 : 0;
        cosp[i] = cp;
    }
}

// useTriangulationMethod on rays the caller already holds (Geometry.h:66-69 takes xn1 / xn2, not pixels): no camera
// model, no gates -- the reference's function has neither.  xn[n][3] in, X[n][3] out.
__global__ void __launch_bounds__(kThreads)
triangulate_rays_kernel(int n, const float* __restrict__ xn1, const float* __restrict__ xn2, const __grid_constant__ PairDev pr,
                        int method, int location, float* __restrict__ X1, float* __restrict__ X2) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const F3 a = mk3(xn1[3 * (size_t)i], xn1[3 * (size_t)i + 1], xn1[3 * (size_t)i + 2]);
        const F3 b = mk3(xn2[3 * (size_t)i], xn2[3 * (size_t)i + 1], xn2[3 * (size_t)i + 2]);
        F3 w1, w2;
        triangulate_match(pr, method, location, a, b, a, b, w1, w2);
        X1[3 * (size_t)i + 0] = w1.x; X1[3 * (size_t)i + 1] = w1.y; X1[3 * (size_t)i + 2] = w1.z;
        X2[3 * (size_t)i + 0] = w2.x; X2[3 * (size_t)i + 1] = w2.y; X2[3 * (size_t)i + 2] = w2.z;
    }
}

// KeyFrame::setInitialDepthScaleInSimulationImages (KeyFrame.cc:131-153): sum of d / z_c, count
__global__ void __launch_bounds__(kThreads)
depth_scale_kernel(int n, const float* __restrict__ X, const float* __restrict__ dep, const uint8_t* __restrict__ valid,
                   const __grid_constant__ PoseF T, double* __restrict__ part /* [grid][2] */) {
    __shared__ double sm[2 * (kThreads / 32)];
    double v[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float d = dep[i];
        if (!valid[i] || d == 0.f) continue;
        F3 c = apply(T, mk3(X[3 * (size_t)i], X[3 * (size_t)i + 1], X[3 * (size_t)i + 2]));
        v[0] += (double)fdv(d, c.z);
        v[1] += 1.0;
    }
    block_reduce<2>(v, sm);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = v[0]; part[2 * blockIdx.x + 1] = v[1]; }
}

// ================================================================== state set-up
__global__ void __launch_bounds__(kThreads)
init_state_kernel(int n, const float* __restrict__ X1, const float* __restrict__ X2, const int* __restrict__ perm,
                  double* __restrict__ P) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        size_t s = perm ? (size_t)perm[i] : (size_t)i;
        double4* p = reinterpret_cast<double4*>(P);
        p[i] = make_double4((double)X1[3 * s], (double)X1[3 * s + 1], (double)X1[3 * s + 2], 0.0);
        p[(size_t)n + i] = make_double4((double)X2[3 * s], (double)X2[3 * s + 1], (double)X2[3 * s + 2], 0.0);
    }
}

// ================================================================== K2 + K3 reductions (the kernels are in dsc_kernels_ell.cuh)
// per-block partial layout of the linearisation: chi2[3], max diagonal, global gradient bg[8], C_TT packed (21), C_ss (2)
constexpr int kLinPart = 3 + 1 + 8 + 21 + 2;   // 35
__global__ void __launch_bounds__(kThreads)
finalize_linearize_kernel(int nb, const double* __restrict__ part, LinGlobal* __restrict__ out) {
    __shared__ double sm[kThreads / 32];
    for (int k = 0; k < kLinPart; ++k) {
        double v;
        if (k == 3) {
            double m = 0.0;
            for (int i = threadIdx.x; i < nb; i += blockDim.x) m = fmax(m, part[(size_t)i * kLinPart + 3]);
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
            __syncthreads();
            v = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) v = fmax(v, sm[w]);
        } else {
            v = sum_partials(part + k, nb, kLinPart, sm);
        }
        if (threadIdx.x == 0) {
            if (k < 3) out->chi2[k] = v;
            else if (k == 3) out->maxdiag = v;
            else if (k < 12) out->bg[k - 4] = v;
            else if (k < 33) {
                int idx = k - 12, r = 0;
                while (idx >= 6 - r) { idx -= 6 - r; ++r; }
                int c = r + idx;
                out->C[r * 8 + c] = v; out->C[c * 8 + r] = v;
            } else { int s = 6 + (k - 33); out->C[s * 8 + s] = v; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int r = 0; r < 6; ++r) for (int s = 6; s < 8; ++s) { out->C[r * 8 + s] = 0.0; out->C[s * 8 + r] = 0.0; }
        out->C[6 * 8 + 7] = 0.0; out->C[7 * 8 + 6] = 0.0;
        double m = out->maxdiag;
        for (int r = 0; r < 8; ++r) m = fmax(m, fabs(out->C[r * 8 + r]));
        out->maxdiag = m;                                 // computeLambdaInit: max over ALL vertices' diagonals
    }
}

// Packed 6x6 blocks (D, Minv) are stored slice-major, [slice of 32 rows][21][32]: entry k of row i sits at
// blk21(base, i)[k * 32], so a warp reads entry k of its 32 rows as one 256-byte run.
template <typename T>
DSC_D T* blk21(T* base, int i) { return base + ((size_t)(i >> 5) * 21) * 32 + (i & 31); }

// block-Jacobi preconditioner: Minv_i = (D_i + lambda I)^-1 (packed), Ginv = (C + lambda I)^-1
template <bool kRO = true, typename T = double>
DSC_D void precond_block(const double* __restrict__ D, int i, double lambda, T* __restrict__ Minv, int* __restrict__ err,
                         double (&M)[21]) {
    const double* Dp = blk21(D, i);
    double A[36], Ai[36];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) {
            double v = (kRO ? Dp[pk<6>(r, c) * 32] : __ldcg(Dp + pk<6>(r, c) * 32)) + (r == c ? lambda : 0.0);
            A[r * 6 + c] = v; A[c * 6 + r] = v;
        }
    if (!spd_inverse<6>(A, Ai)) {
        atomicExch(err, 1);
#pragma unroll
        for (int k = 0; k < 36; ++k) Ai[k] = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) Ai[r * 6 + r] = 1.0 / fmax(fabs(A[r * 6 + r]), 1e-300);
    }
    T* Mp = blk21(Minv, i);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) {
            const T stored = (T)Ai[r * 6 + c];           // the preconditioner the iterations will apply is the STORED one
            M[pk<6>(r, c)] = (double)stored; Mp[pk<6>(r, c) * 32] = stored;
        }
}
template <bool kRO = true>
DSC_D void precond_global(const LinGlobal* __restrict__ lin, double lambda, double* __restrict__ Ginv, int* __restrict__ err) {
    double A[64], Ai[64];
    for (int k = 0; k < 64; ++k) A[k] = kRO ? lin->C[k] : __ldcg(&lin->C[k]);
    for (int r = 0; r < 8; ++r) A[r * 8 + r] += lambda;
    if (!spd_inverse<8>(A, Ai)) {
        atomicExch(err, 1);
        for (int k = 0; k < 64; ++k) Ai[k] = 0.0;
        for (int r = 0; r < 8; ++r) Ai[r * 8 + r] = 1.0 / fmax(fabs(A[r * 8 + r]), 1e-300);
    }
    for (int k = 0; k < 64; ++k) Ginv[k] = Ai[k];
}

template <typename T>
DSC_D void apply_minv(const T* __restrict__ Mp, const double* r, double* z) {
    double M[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) M[k] = (double)Mp[k * 32];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) s += (a <= c ? M[pk<6>(a, c)] : M[pk<6>(c, a)]) * r[c];
        z[a] = s;
    }
}

// ================================================================== K4: PCG (Chronopoulos-Gear form)
//   z = M^-1 r ; w = A z ; gamma = r.z ; delta = z.w
//   beta = gamma/gamma_prev ; alpha = gamma / (delta - beta gamma / alpha_prev)
//   p = z + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s
// Two kernels per iteration (update, spmv); all scalars stay on the device.
template <typename T>
struct CgVecsT {
    T *x, *r, *z, *w, *p, *s;               // [n][6]   (T = float in the fp32 mode: 24-byte records)
    double *xg, *rg, *zg, *wg, *pg, *sg;    // [8]
};
using CgVecs = CgVecsT<double>;

// Preconditioner and PCG start in one pass over the rows: Minv = (D + lambda I)^-1, r = b, z = Minv r, gamma partial.
// x, p and s are not written: the first cg_update (first = 1) treats them as zero.
// TM: storage type of the preconditioner blocks (float also in the fp64 mode: any fixed SPD M is a valid preconditioner, the
// iterations apply the STORED one, and 84 instead of 168 bytes per row and iteration are streamed)
template <typename T, typename TM = T>
__global__ void __launch_bounds__(kThreads)
cg_init_kernel(int n, const double* __restrict__ b, const double* __restrict__ D, double lambda, const LinGlobal* __restrict__ lin,
               TM* __restrict__ Minv, double* __restrict__ Ginv, int* __restrict__ err, CgVecsT<T> v, double* __restrict__ gpart,
               CgControl* __restrict__ ctl) {
    __shared__ double sm[kThreads / 32];
    double g[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double r[6], z[6], M[21];
        D3 a, c;
        load6(b, i, a, c);
        if (sizeof(T) == 4) { a = d3((float)a.x, (float)a.y, (float)a.z); c = d3((float)c.x, (float)c.y, (float)c.z); }   // r as stored
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = c.x; r[4] = c.y; r[5] = c.z;
        precond_block<true, TM>(D, i, lambda, Minv, err, M);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += (q <= k ? M[pk<6>(q, k)] : M[pk<6>(k, q)]) * r[k];
            z[q] = sizeof(T) == 4 ? (double)(float)sacc : sacc;
        }
        store6(v.r, i, a, c);
        store6(v.z, i, d3(z[0], z[1], z[2]), d3(z[3], z[4], z[5]));
#pragma unroll
        for (int k = 0; k < 6; ++k) g[0] += r[k] * z[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        precond_global(lin, lambda, Ginv, err);
        for (int a = 0; a < 8; ++a) {
            double s = 0.0;
            for (int c = 0; c < 8; ++c) s += Ginv[a * 8 + c] * lin->bg[c];
            v.rg[a] = lin->bg[a]; v.zg[a] = s; v.xg[a] = 0.0; v.pg[a] = 0.0; v.sg[a] = 0.0;
            g[0] += lin->bg[a] * s;
        }
        ctl->iters = 0; ctl->converged = 0; ctl->breakdown = 0;
        ctl->sc[0].gamma_prev = 1.0; ctl->sc[0].alpha_prev = 1.0;
        ctl->sc[1].gamma_prev = 1.0; ctl->sc[1].alpha_prev = 1.0;
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) gpart[blockIdx.x] = g[0];
}

// w = (H + lambda I) z.  The ARAP part streams the per-edge Jacobian record Je = {u, m, g} written by
// linearize_kernel (gradient of e w.r.t. X1i,X2i,X1j,X2j,T_g is  -m-2g | u+2Rg^T g | m-2g | -u+2Rg^T g |
// [2 (X1i+X1j) x g, -4 g]), so with dz1 = z1i-z1j, dz2 = z2i-z2j, sz1 = z1i+z1j, sz2 = z2i+z2j:
//     s_e = u.dz2 - m.dz1 + 2 g.(Rg sz2 - sz1 + z_w x (X1i+X1j) - 2 z_v)
//     Am = sum_e 2W s_e m, Ag = sum_e 2W s_e g, Au = sum_e 2W s_e u          (per vertex, over its CSR row)
//     w_i1 = -Am - 2 Ag,   w_i2 = Au + 2 Rg^T Ag
//     T_g rows: the directed twins carry the same s_e and g, so  w_w = 2 sum_i X1i x Ag_i,  w_v = -2 sum_i Ag_i.
// Layout / work decomposition: sliced ELL.  A slice is 32 consecutive rows (correspondences) = one warp, lane l
// owns row 32 s + l for the whole slice: its z_i / X1_i stay in registers, its three sums Am/Ag/Au accumulate in
// registers, and NO cross-lane reduction is needed.  Column k of the slice is block sliceptr[s] + k: one coalesced
// 128-byte line of neighbour indices and nine coalesced 256-byte lines of Jacobian records.  Rows are degree-sorted
// inside groups of 512 on the host, so the 32 rows of a slice have (almost) the same length; padding slots point
// at the row itself and carry an all-zero record.  Only z_j (48 B) and X1_j (32 B) are gathered (L1/L2: the
// space-filling-curve order keeps neighbours close).  dpart[grid]: partial z.w (global rows included),
// bpart[grid][8]: partial global rows (T_g from the ARAP sums, s1/s2 from the depth edges).
// Bulk asynchronous copies (cp.async.bulk, the non-tensor TMA path) + mbarrier completion, used by cg_spmv_kernel to
// stream the ELL blocks into a per-warp shared-memory ring: the bytes in flight no longer depend on registers.
DSC_D unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DSC_D void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
DSC_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
DSC_D void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
DSC_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "MBAR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra MBAR_DONE;\n"
                 "bra MBAR_WAIT;\n"
                 "MBAR_DONE:\n"
                 "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

constexpr int kURec = 14;                                        // unary record {U1[6], U2[6], kd1, kd2}, stored [slice][kURec][32]
// Per storage type T of the operator data (Je, U) and of the vectors (double, or float in the fp32 mode): ring depth per
// warp (ELL blocks in flight: the float blocks are half as large, so more of them keep the same bytes in flight), bytes
// of one ELL block (Je[9][32] + ecol[32] ints), window bytes (z: 6 T per row, X1: double4 per row).
template <typename T> struct SpmvCfg {
    static constexpr int kStages = sizeof(T) == 8 ? 3 : 6;
    static constexpr int kJeBytes = 9 * 32 * (int)sizeof(T);
    static constexpr int kStageBytes = kJeBytes + 32 * 4;
    static constexpr int kZRow = 6 * (int)sizeof(T);
    static constexpr int kWinBytes = kSortGroup * (kZRow + 32);
    static constexpr int kSmem = kWinBytes + (kThreads / 32) * kStages * kStageBytes;
};
constexpr int kSpmvSmem = SpmvCfg<double>::kSmem;

// kShard: one rank of a point-sharded pair (dsc_shard.cuh) -- the work units cover the rank's own slices only, z holds
// the halo rows the peers pushed (waited for in the prologue), and the LAST block of the launch to finish sends the
// rank's totals of z.w and of the 8 global rows to every rank.
template <typename T, bool kShard = false>
__global__ void __launch_bounds__(kThreads, 2)
cg_spmv_kernel(int n, const double* __restrict__ P, const T* __restrict__ Je, const T* __restrict__ U,
               const int* __restrict__ sliceptr, const int* __restrict__ ecol, const int* __restrict__ part, int nunits,
               const Globals* __restrict__ Gp, const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W,
               double lambda, const T* __restrict__ z, const double* __restrict__ zg, T* __restrict__ w,
               double* __restrict__ dpart, double* __restrict__ bpart,
               const LinGlobal* __restrict__ lin, const CgControl* __restrict__ ctl, const __grid_constant__ ShardDev S) {
    constexpr int kSpmvStages = SpmvCfg<T>::kStages, kSpmvStageBytes = SpmvCfg<T>::kStageBytes, kSpmvWinBytes = SpmvCfg<T>::kWinBytes;
    constexpr int kJeBytes = SpmvCfg<T>::kJeBytes, kZRow = SpmvCfg<T>::kZRow;
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ double sm[9 * (kThreads / 32)];
    __shared__ double Rg[9];
    __shared__ double zgs[8];
    __shared__ unsigned long long bars[(kThreads / 32) * kSpmvStages + 1];
    const T* sz = reinterpret_cast<const T*>(dyn);                               // z of the tile: 6 T per row
    double4* sx = reinterpret_cast<double4*>(dyn + kSortGroup * kZRow);          // X1 of the tile
    if (ctl && (ctl->converged || ctl->breakdown)) return;   // flags are only written by an EARLIER launch
    if (ctl) lambda = ctl->lambda;
    if (kShard) shard_wait(S, SF_Z, shard_sent(S, SF_Z));    // the peers' halo rows of z have arrived
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = kThreads / 32;
    if (threadIdx.x < 9) Rg[threadIdx.x] = Gp->Rg[threadIdx.x];
    if (threadIdx.x < 8) zgs[threadIdx.x] = zg[threadIdx.x];
    if (threadIdx.x <= wpb * kSpmvStages) mbar_init(bars + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned long long* wbar = bars + wpb * kSpmvStages;                         // window barrier
    unsigned long long* rbar = bars + warp * kSpmvStages;                        // this warp's ring barriers
    unsigned char* ring = dyn + kSpmvWinBytes + warp * kSpmvStages * kSpmvStageBytes;
    const D3 zw = d3(zgs[0], zgs[1], zgs[2]);
    const D3 zv2 = d3(2.0 * zgs[3], 2.0 * zgs[4], 2.0 * zgs[5]);
    double acc[9];                                     // T_g border (6), s1, s2 border, z.w
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    // Work units: slice ranges part[u] .. part[u+1], taken round-robin (unit u by block u mod grid, so that at any
    // time the blocks work on neighbouring rows: halo gathers hit L2 and the DRAM streams stay close).  Full rounds
    // are tiles of kSortGroup rows = 16 slices; the host cuts the last, partial round into one equal-work unit per
    // block, so there is no tail round with idle SMs.  The tile's z and X1 are staged in shared memory (two bulk
    // copies) and a neighbour inside the tile is read from there; halo neighbours come from L2.  Warp w walks
    // slices w and 15 - w of the tile (balanced: the slices are degree-sorted inside groups of kSortGroup rows);
    // their ELL blocks form one stream of kSpmvStageBytes records that lane 0 keeps kSpmvStages deep in flight.
    unsigned cons = 0;                                 // ring slots consumed so far by this warp (stage, phase)
    unsigned wphase = 0;
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int s_begin = __ldg(part + unit), s_end = __ldg(part + unit + 1);
    for (int ts = s_begin; ts < s_end; ts += kSortGroup / 32) {
        const int nsl = min(kSortGroup / 32, s_end - ts);
        const int v0 = ts * 32;
        const int nv = min(nsl * 32, n - v0);
        const int lsv[2] = {warp, 2 * wpb - 1 - warp};
        int rb[2], rl[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool ok = lsv[h] < nsl;
            const int sl = ts + lsv[h];
            rb[h] = ok ? __ldg(sliceptr + sl) : 0;
            rl[h] = ok ? __ldg(sliceptr + sl + 1) - rb[h] : 0;
        }
        const int total = rl[0] + rl[1];
        auto issue = [&](int q, unsigned slot, int dep) {   // lane 0: fetch stream element q into ring slot
            if (dep == 0x7ff4d5c1) q = 0;              // (practically) never taken; keeps the dependency alive
            const int bk = q < rl[0] ? rb[0] + q : rb[1] + (q - rl[0]);
            unsigned char* dst = ring + (slot % kSpmvStages) * kSpmvStageBytes;
            unsigned long long* bar = rbar + (slot % kSpmvStages);
            mbar_expect_tx(bar, kSpmvStageBytes);
            bulk_load(dst, Je + (size_t)bk * 288, kJeBytes, bar);
            bulk_load(dst + kJeBytes, ecol + (size_t)bk * 32, 128, bar);
        };
        __syncthreads();                               // everyone is done with the previous window
        if (threadIdx.x == 0) {
            // (bulk copies move multiples of 16 bytes: an odd number of 24-byte float rows is rounded up; the vector
            // buffers are allocated for doubles, so the 8 extra bytes exist)
            const unsigned zbytes = ((unsigned)nv * (unsigned)kZRow + 15u) & ~15u;
            mbar_expect_tx(wbar, zbytes + (unsigned)nv * 32u);
            bulk_load(dyn, z + 6 * (size_t)v0, zbytes, wbar);
            bulk_load(sx, reinterpret_cast<const double4*>(P) + v0, (unsigned)nv * 32u, wbar);
        }
        if (lane == 0)
            for (int q = 0; q < kSpmvStages && q < total; ++q) issue(q, cons + q, 0);
        mbar_wait(wbar, wphase);
        wphase ^= 1u;
        int q = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (lsv[h] >= nsl) continue;                   // (a slice without edges still owns its rows)
        const int il = lsv[h] * 32 + lane;
        const int i = v0 + il;
        const bool act = il < nv;
        const int ilc = act ? il : nv - 1;
        D3 zi1, zi2;
        load6(sz, ilc, zi1, zi2);
        const double4 xi = sx[ilc];
        const D3 X1i = d3(xi.x, xi.y, xi.z);
        D3 Am = d3(0, 0, 0), Ag = d3(0, 0, 0), Au = d3(0, 0, 0);
        for (int e = 0; e < rl[h]; ++e, ++q, ++cons) {
            const unsigned st = cons % kSpmvStages;
            mbar_wait(rbar + st, (cons / kSpmvStages) & 1u);
            const unsigned char* sb = ring + st * kSpmvStageBytes;
            const T* jb = reinterpret_cast<const T*>(sb) + lane;
            const int j = reinterpret_cast<const int*>(sb + kJeBytes)[lane];
            const D3 u = d3((double)jb[0], (double)jb[32], (double)jb[64]);
            const D3 m = d3((double)jb[96], (double)jb[128], (double)jb[160]);
            const D3 g = d3((double)jb[192], (double)jb[224], (double)jb[256]);
            D3 zj1, zj2, X1j;
            const unsigned jl = (unsigned)(j - v0);
            if (jl < (unsigned)nv) {
                load6(sz, (int)jl, zj1, zj2);
                const double4 xj = sx[jl];
                X1j = d3(xj.x, xj.y, xj.z);
            } else {
                load6(z, j, zj1, zj2);
                const double4 xj = ldg256(reinterpret_cast<const double4*>(P) + (size_t)j);
                X1j = d3(xj.x, xj.y, xj.z);
            }
            const D3 S1 = X1i + X1j;
            const D3 t = mul(Rg, zi2 + zj2) - (zi1 + zj1) + cross(zw, S1) - zv2;
            const double s = dot(u, zi2 - zj2) - dot(m, zi1 - zj1) + 2.0 * dot(g, t);
            const double w2 = 2.0 * W.arap_info * s;
            Am = Am + w2 * m; Ag = Ag + w2 * g; Au = Au + w2 * u;
            // The slot is free once every lane holds its record in registers: refill it with element q + kSpmvStages.
            // The copy is issued by the async proxy, which is not ordered after this warp's shared-memory loads by
            // program order alone, so the issue carries a true data dependency on s (= on all ten loads of the slot).
            // (Issuing it earlier, right after the loads, serialises them with the gathers and is 15 % slower.)
            const int dep = __shfl_sync(0xffffffffu, __double2hiint(s), 0);
            if (lane == 0 && q + kSpmvStages < total) issue(q + kSpmvStages, cons + kSpmvStages, dep);
        }
        if (act) {
            // output rows: [-Am - 2 Ag | Au + 2 Rg^T Ag] + U z + kd n z_s + lambda z
            const T* Up = U + ((size_t)(i >> 5) * kURec) * 32 + lane;
            double uu[kURec];
#pragma unroll
            for (int k = 0; k < kURec; ++k) uu[k] = (double)__ldg(Up + k * 32);
            const D3 rg = mulT(Rg, Ag);
            double out[6] = {-Am.x - 2.0 * Ag.x, -Am.y - 2.0 * Ag.y, -Am.z - 2.0 * Ag.z,
                             Au.x + 2.0 * rg.x, Au.y + 2.0 * rg.y, Au.z + 2.0 * rg.z};
            const double zi[6] = {zi1.x, zi1.y, zi1.z, zi2.x, zi2.y, zi2.z};
#pragma unroll
            for (int cam = 0; cam < 2; ++cam) {
                const double* R = cam == 0 ? pr.R1 : pr.R2;
                const double nz = R[6] * zi[cam * 3] + R[7] * zi[cam * 3 + 1] + R[8] * zi[cam * 3 + 2];
                const double kd = uu[12 + cam];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double sum = 0.0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) sum += (r <= c ? uu[cam * 6 + pk<3>(r, c)] : uu[cam * 6 + pk<3>(c, r)]) * zi[cam * 3 + c];
                    out[cam * 3 + r] += sum + kd * R[6 + r] * zgs[6 + cam];
                }
                acc[6 + cam] += kd * nz;                           // s1/s2 rows: kd (n . z)
            }
            double dl = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                out[k] += lambda * zi[k];
                if (sizeof(T) == 4) out[k] = (double)(float)out[k];   // z.w of the w the update kernel will read
                dl += zi[k] * out[k];
            }
            acc[8] += dl;
            // T_g rows (the directed twins carry the same s_e and g): omega 2 (X1i x Ag), upsilon -2 Ag
            const D3 cx = cross(X1i, Ag);
            acc[0] += 2.0 * cx.x; acc[1] += 2.0 * cx.y; acc[2] += 2.0 * cx.z;
            acc[3] -= 2.0 * Ag.x; acc[4] -= 2.0 * Ag.y; acc[5] -= 2.0 * Ag.z;
            store6(w, i, d3(out[0], out[1], out[2]), d3(out[3], out[4], out[5]));
        }
      }
    }
    }
    block_reduce<9>(acc, sm);
    if (threadIdx.x == 0) {
        double dl = acc[8];
        for (int k = 0; k < 8; ++k) { bpart[8 * (size_t)blockIdx.x + k] = acc[k]; dl += zgs[k] * acc[k]; }
        if (blockIdx.x == 0 && (!kShard || S.rank == 0))     // diagonal of the global rows that is not inside the edge sums: (C_ss + lambda)
            for (int k = 0; k < 8; ++k) dl += zgs[k] * ((k >= 6 ? lin->C[k * 8 + k] : 0.0) + lambda) * zgs[k];
        dpart[blockIdx.x] = dl;
    }
    if (kShard) {
        __shared__ double tot[9];
        if (!shard_last_block(S, SF_S)) return;
        double t9[9];                                      // fixed-order totals of dpart[grid] and bpart[grid][8], all threads loading
#pragma unroll
        for (int e = 0; e < 9; ++e) t9[e] = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            t9[0] += __ldcg(dpart + i);
#pragma unroll
            for (int e = 1; e < 9; ++e) t9[e] += __ldcg(bpart + 8 * (size_t)i + (e - 1));
        }
        block_reduce<9>(t9, sm);
        if (threadIdx.x == 0) for (int e = 0; e < 9; ++e) tot[e] = t9[e];
        __syncthreads();
        shard_send(S, SF_S, tot, 9);
    }
}

// One CG step.  par = iteration parity (double-buffered scalars and gamma partials).
template <typename T, typename TM = T>
__global__ void __launch_bounds__(kThreads)
cg_update_kernel(int n, int par, int first, const TM* __restrict__ Minv, const double* __restrict__ Ginv,
                 const LinGlobal* __restrict__ lin, double lambda, CgVecsT<T> v,
                 const double* __restrict__ gpart_in, double* __restrict__ gpart_out,
                 const double* __restrict__ dpart, const double* __restrict__ bpart, int nspmv,
                 CgControl* __restrict__ ctl, double rtol2) {
    __shared__ double sm[kThreads / 32];
    if (ctl->converged || ctl->breakdown) return;            // set by an earlier launch
    lambda = ctl->lambda;
    rtol2 = ctl->rtol2;
    double gamma = sum_partials(gpart_in, gridDim.x, 1, sm);
    double delta = sum_partials(dpart, nspmv, 1, sm);
    if (first && blockIdx.x == 0 && threadIdx.x == 0) ctl->gamma0 = gamma;
    double gamma0 = first ? gamma : ctl->gamma0;
    if (!first && gamma <= rtol2 * gamma0) {                  // same fixed-order sums in every block => same decision
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->converged = 1;
        return;
    }
    CgScalars prev = ctl->sc[par ^ 1];
    double beta = first ? 0.0 : gamma / prev.gamma_prev;
    double denom = first ? delta : delta - beta * gamma / prev.alpha_prev;
    double alpha = gamma / denom;
    bool bad = !(denom > 0.0) || !isfinite(alpha);
    if (bad) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->breakdown = 1;
        return;
    }
    double g[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 z1, z2, w1, w2, p1, p2, s1, s2, x1, x2, r1, r2;
        load6(v.z, i, z1, z2); load6(v.w, i, w1, w2); load6(v.r, i, r1, r2);
        if (first) {                                         // x = p = s = 0 (never written by cg_init_kernel)
            p1 = z1; p2 = z2; s1 = w1; s2 = w2;
            x1 = alpha * p1; x2 = alpha * p2;
        } else {
            load6(v.p, i, p1, p2); load6(v.s, i, s1, s2); load6(v.x, i, x1, x2);
            p1 = z1 + beta * p1; p2 = z2 + beta * p2;
            s1 = w1 + beta * s1; s2 = w2 + beta * s2;
            x1 = x1 + alpha * p1; x2 = x2 + alpha * p2;
        }
        r1 = r1 - alpha * s1; r2 = r2 - alpha * s2;
        if (sizeof(T) == 4) {                                // continue from the residual as it is stored
            r1 = d3((float)r1.x, (float)r1.y, (float)r1.z); r2 = d3((float)r2.x, (float)r2.y, (float)r2.z);
        }
        double r[6] = {r1.x, r1.y, r1.z, r2.x, r2.y, r2.z}, zn[6];
        apply_minv(blk21(Minv, i), r, zn);
        if (sizeof(T) == 4) {
#pragma unroll
            for (int k = 0; k < 6; ++k) zn[k] = (double)(float)zn[k];     // gamma = r.z of the stored z
        }
        store6(v.p, i, p1, p2); store6(v.s, i, s1, s2); store6(v.x, i, x1, x2); store6(v.r, i, r1, r2);
        store6(v.z, i, d3(zn[0], zn[1], zn[2]), d3(zn[3], zn[4], zn[5]));
#pragma unroll
        for (int k = 0; k < 6; ++k) g[0] += r[k] * zn[k];
    }
    if (blockIdx.x == 0) {
        __shared__ double wg[8], rgn[8];
        __syncthreads();
        if (threadIdx.x < 8) {
            int k = threadIdx.x;
            double s = 0.0;
            for (int bk = 0; bk < nspmv; ++bk) s += bpart[8 * (size_t)bk + k];
            double d = (k >= 6 ? lin->C[k * 8 + k] : 0.0) + lambda;
            wg[k] = s + d * v.zg[k];
            double pg = v.zg[k] + beta * v.pg[k];
            double sg = wg[k] + beta * v.sg[k];
            v.pg[k] = pg; v.sg[k] = sg;
            v.xg[k] += alpha * pg;
            double rr = v.rg[k] - alpha * sg;
            v.rg[k] = rr; rgn[k] = rr;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            int k = threadIdx.x;
            double s = 0.0;
            for (int c = 0; c < 8; ++c) s += Ginv[k * 8 + c] * rgn[c];
            v.zg[k] = s;
            wg[k] = s * rgn[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 0; k < 8; ++k) g[0] += wg[k];
            ctl->sc[par].gamma_prev = gamma; ctl->sc[par].alpha_prev = alpha;
            ctl->iters = ctl->iters + 1;
        }
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) gpart_out[blockIdx.x] = g[0];
}

// ------------------------------------------------------------------ fp32 mode: mixed-precision iterative refinement
// The float PCG (cg_*_kernel<float>) solves for a CORRECTION; the solution itself and the residual it restarts from are
// double:   X += x_f ;  r = b - (H + lambda I) X  (double operator, cg_spmv_kernel<double>) ;  solve (H + lambda I) e = r ...
// X <- base + x_f (base = nullptr: X <- x_f); the 8 globals alike.  Also used to peek at the running solution for a
// trial evaluation without disturbing the solve (out != base).
__global__ void __launch_bounds__(kThreads)
refine_accumulate_kernel(int n, const double* base, const float* __restrict__ xf, double* out,
                         const double* baseg, const double* __restrict__ xg, double* outg, int add_xf) {     // (out may be base)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 a = d3(0, 0, 0), c = d3(0, 0, 0), u = d3(0, 0, 0), v = d3(0, 0, 0);
        if (base) load6(base, i, a, c);
        if (add_xf) load6(xf, i, u, v);
        store6(out, i, a + u, c + v);
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) outg[threadIdx.x] = (baseg ? baseg[threadIdx.x] : 0.0) + (add_xf ? xg[threadIdx.x] : 0.0);
}
// Restart of the float PCG from the TRUE residual: r = b - Wd, where Wd = (H + lambda I) X was just applied in double
// (bpart: the block partials of its global rows); z = Minv r; gamma partial (double, from the unrounded r and z); the
// float vectors r, z are what the iterations continue from; x, p, s are treated as zero by the first update.
__global__ void __launch_bounds__(kThreads)
cg_restart_kernel(int n, const double* __restrict__ b, const double* __restrict__ Wd, const float* __restrict__ Minv,
                  const double* __restrict__ Ginv, const LinGlobal* __restrict__ lin, const double* __restrict__ Xg,
                  const double* __restrict__ bpart, int nspmv, CgVecsT<float> v, double* __restrict__ gpart, CgControl* __restrict__ ctl) {
    __shared__ double sm[kThreads / 32];
    double g[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        D3 b1, b2, w1, w2;
        load6(b, i, b1, b2); load6(Wd, i, w1, w2);
        const D3 r1 = b1 - w1, r2 = b2 - w2;
        const double r[6] = {r1.x, r1.y, r1.z, r2.x, r2.y, r2.z};
        double z[6];
        apply_minv(blk21(Minv, i), r, z);
        store6(v.r, i, r1, r2);
        store6(v.z, i, d3(z[0], z[1], z[2]), d3(z[3], z[4], z[5]));
#pragma unroll
        for (int k = 0; k < 6; ++k) g[0] += r[k] * z[k];
    }
    if (blockIdx.x == 0) {
        __shared__ double rgn[8];
        __syncthreads();
        if (threadIdx.x < 8) {
            const int k = threadIdx.x;
            double s = 0.0;
            for (int bk = 0; bk < nspmv; ++bk) s += bpart[8 * (size_t)bk + k];
            const double d = (k >= 6 ? lin->C[k * 8 + k] : 0.0) + ctl->lambda;
            rgn[k] = lin->bg[k] - (s + d * Xg[k]);
            v.rg[k] = rgn[k]; v.xg[k] = 0.0; v.pg[k] = 0.0; v.sg[k] = 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int a = 0; a < 8; ++a) {
                double s = 0.0;
                for (int c = 0; c < 8; ++c) s += Ginv[a * 8 + c] * rgn[c];
                v.zg[a] = s;
                g[0] += rgn[a] * s;
            }
            ctl->iters = 0; ctl->converged = 0;
            ctl->sc[0].gamma_prev = 1.0; ctl->sc[0].alpha_prev = 1.0;
            ctl->sc[1].gamma_prev = 1.0; ctl->sc[1].alpha_prev = 1.0;
        }
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) gpart[blockIdx.x] = g[0];
}
// gamma_true = fixed-order sum of the restart's partials, left in the control block for the host
__global__ void __launch_bounds__(kThreads)
refine_gamma_kernel(int nb, const double* __restrict__ gpart, CgControl* __restrict__ ctl) {
    __shared__ double sm[kThreads / 32];
    const double g = sum_partials(gpart, nb, 1, sm);
    if (threadIdx.x == 0) ctl->gamma_true = g;
}

// host -> CgControl without a (pageable, stream-serialising) memcpy: one thread writes the fields that are >= 0 / set
__global__ void ctl_set_kernel(CgControl* ctl, double lambda, int set_lambda, double rtol2, int set_rtol, int clear_converged) {
    if (set_lambda) ctl->lambda = lambda;
    if (set_rtol) ctl->rtol2 = rtol2;
    if (clear_converged) ctl->converged = 0;
}

// ================================================================== LM trial: x_new = x (+) dx
// Ptrial = P + dx (points), trial globals = exp(dx_T) * T_g, s + ds; scale partial = sum dx (lambda dx + b)
template <bool kRO, typename T = double>
DSC_D double apply_update_rows(int n, int first, int stride, const double* __restrict__ P, const T* __restrict__ x,
                               const double* __restrict__ b, double lambda, double* __restrict__ Ptrial) {
    double acc = 0.0;
    for (int i = first; i < n; i += stride) {
        D3 x1, x2, b1, b2;
        load6_t<kRO>(x, i, x1, x2); load6_t<kRO>(b, i, b1, b2);
        P8 Pi = load_P<kRO>(P, n, i);
        double4* o = reinterpret_cast<double4*>(Ptrial);
        o[i] = make_double4(Pi.a.x + x1.x, Pi.a.y + x1.y, Pi.a.z + x1.z, 0.0);
        o[(size_t)n + i] = make_double4(Pi.b.x + x2.x, Pi.b.y + x2.y, Pi.b.z + x2.z, 0.0);
        acc += dot(x1, lambda * x1 + b1) + dot(x2, lambda * x2 + b2);
    }
    return acc;
}
// the 8 global unknowns of the trial state; returns their share of dx.(lambda dx + b)
DSC_D double apply_update_globals(const Globals& g, const double* xg, const double* bg, double lambda, Globals& o) {
    double upd[6];
    for (int k = 0; k < 6; ++k) upd[k] = xg[k];
    double T[7];
    se3_oplus(g.Tg, upd, T);
    for (int k = 0; k < 7; ++k) o.Tg[k] = T[k];
    o.s1 = g.s1 + xg[6]; o.s2 = g.s2 + xg[7];
    quat_to_rot(o.Tg, o.Rg);
    double acc = 0.0;
    for (int k = 0; k < 8; ++k) acc += xg[k] * (lambda * xg[k] + bg[k]);
    return acc;
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_update_kernel(int n, const double* __restrict__ P, const T* __restrict__ x, const double* __restrict__ xg,
                    const double* __restrict__ b, const LinGlobal* __restrict__ lin, double lambda,
                    const Globals* __restrict__ Gcur, double* __restrict__ Ptrial, Globals* __restrict__ Gtrial,
                    double* __restrict__ part) {
    __shared__ double sm[kThreads / 32];
    double acc[1];
    acc[0] = apply_update_rows<true, T>(n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, P, x, b, lambda, Ptrial);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Globals g = *Gcur, o;
        acc[0] += apply_update_globals(g, xg, lin->bg, lambda, o);
        *Gtrial = o;
    }
    block_reduce<1>(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

// ================================================================== write-back helpers
// points cast to float in the caller's order; update = sum |p_old - p_new| (float norm, g2oBundleAdjustment.cc:978-990)
__global__ void __launch_bounds__(kThreads)
export_kernel(int n, const double* __restrict__ P, const double* __restrict__ P0, const int* __restrict__ perm,
              float* __restrict__ X1, float* __restrict__ X2, double* __restrict__ X1d, double* __restrict__ X2d,
              double* __restrict__ part) {
    __shared__ double sm[kThreads / 32];
    double acc[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        size_t d = perm ? (size_t)perm[i] : (size_t)i;
        P8 a = load_P(P, n, i), o = load_P(P0, n, i);
        float n1[3] = {(float)a.a.x, (float)a.a.y, (float)a.a.z}, n2[3] = {(float)a.b.x, (float)a.b.y, (float)a.b.z};
        float o1[3] = {(float)o.a.x, (float)o.a.y, (float)o.a.z}, o2[3] = {(float)o.b.x, (float)o.b.y, (float)o.b.z};
        for (int k = 0; k < 3; ++k) {
            X1[3 * d + k] = n1[k]; X2[3 * d + k] = n2[k];
        }
        if (X1d) { X1d[3 * d] = a.a.x; X1d[3 * d + 1] = a.a.y; X1d[3 * d + 2] = a.a.z; }
        if (X2d) { X2d[3 * d] = a.b.x; X2d[3 * d + 1] = a.b.y; X2d[3 * d + 2] = a.b.z; }
        float e1 = fsq(fa(fa(fm(fs(o1[0], n1[0]), fs(o1[0], n1[0])), fm(fs(o1[1], n1[1]), fs(o1[1], n1[1]))), fm(fs(o1[2], n1[2]), fs(o1[2], n1[2]))));
        float e2 = fsq(fa(fa(fm(fs(o2[0], n2[0]), fs(o2[0], n2[0])), fm(fs(o2[1], n2[1]), fs(o2[1], n2[1]))), fm(fs(o2[2], n2[2]), fs(o2[2], n2[2]))));
        acc[0] += (double)e1 + (double)e2;
    }
    block_reduce<1>(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

// calculatePixelsStandDev (Utils/Geometry.cc:409-458): per camera sums of squared |obs - project(T * (float)X)|
// part: [grid][4] = {su1, sv1, su2, sv2}
__global__ void __launch_bounds__(kThreads)
pixel_sigma_kernel(int n, const double* __restrict__ P, const float4* __restrict__ uv,
                   const __grid_constant__ PairDev pr, double* __restrict__ part) {
    __shared__ double sm[4 * (kThreads / 32)];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        P8 a = load_P(P, n, i);
        float4 o = uv[i];
        float pu, pv;
        cam_project(pr.cam1, apply(pr.T1f, mk3((float)a.a.x, (float)a.a.y, (float)a.a.z)), pu, pv);
        double eu = fabs((double)o.x - (double)pu), ev = fabs((double)o.y - (double)pv);
        acc[0] += eu * eu; acc[1] += ev * ev;
        cam_project(pr.cam2, apply(pr.T2f, mk3((float)a.b.x, (float)a.b.y, (float)a.b.z)), pu, pv);
        eu = fabs((double)o.z - (double)pu); ev = fabs((double)o.w - (double)pv);
        acc[2] += eu * eu; acc[3] += ev * ev;
    }
    block_reduce<4>(acc, sm);
    if (threadIdx.x == 0) for (int k = 0; k < 4; ++k) part[4 * (size_t)blockIdx.x + k] = acc[k];
}

}  // namespace dsc
