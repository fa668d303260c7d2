// dsc_ba.cuh -- the classic bundle-adjustment paths of the reference: key-frame poses + map points +
// reprojection edges, Levenberg-Marquardt with the points marginalised (g2o BlockSolver_6_3: a TRUE Schur complement).
//   bundleAdjustment / poseOnlyOptimization / localBundleAdjustment   Modules/Optimization/g2oBundleAdjustment.cc:38-444
//   EdgeSE3ProjectXYZ, EdgeSE3ProjectXYZOnlyPose                      g2oTypes.h:150-228, g2oTypes.cc:120-180
// Layout: the observations are sorted by map point (CSR pt_ptr): one thread owns a point, walks its observations and keeps
// the point's 3x3 block in registers.  Per observation the linearisation stores the pose-side blocks
//   A = w Jp^T Jp (21, packed), g = -w Jp^T e (6), W = w Jp^T Jx (6x3)
// (component-major and RANK-major: component c of the s-th observation of point j lives at [c * O + slot], slot = ob_slot[o] =
// first slot of rank s + the number of earlier points with more than s observations -- the threads of a warp, consecutive
// points at the same rank, write whole 32-byte sectors in one instruction; interleaving the ranks left half-written sectors
// in L2 long enough to be evicted: 2.1x the algorithmic DRAM traffic) and per point Hll (6, packed) and bl (3).  The reduced camera system S = Hpp - Hpl (Hll + lambda)^-1 Hlp is summed over
// "entries": pairs of observations of the same point seen from free poses a <= b, sorted by (a, b) on the host, cut into
// chunks of one block each -- every sum is a fixed-order two-stage reduction (no atomics), the per-chunk partials are
// folded in order on the host, which also factorises the small system (6 x free poses).
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

struct BaPose { double R[9]; double t[3]; };
constexpr int kBaDiag = 54;                  // chunk partial of an (a, a) segment: sum A (21), sum g (6), sum W Hinv W^T (21), sum W Hinv bl (6)
constexpr int kBaOff = 36;                   // (a, b), a < b: sum W_a Hinv W_b^T (6x6)
constexpr int kBaChunk = 4096;               // entries per chunk (= per block)

struct BaEntryChunk { int begin, end, diag, pad; };

// 3x3 symmetric (packed xx xy xz yy yz zz) + lambda on the diagonal -> inverse (packed); false if not positive definite
DSC_D bool ba_inv3(const double* h, double lambda, double* o) {
    const double a = h[0] + lambda, b = h[1], c = h[2], d = h[3] + lambda, e = h[4], f = h[5] + lambda;
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (!(det > 0.0) || !(a > 0.0) || !(a * d - b * b > 0.0)) { for (int k = 0; k < 6; ++k) o[k] = 0.0; return false; }
    const double r = 1.0 / det;
    o[0] = c00 * r; o[1] = c01 * r; o[2] = c02 * r;
    o[3] = (a * f - c * c) * r; o[4] = (b * c - a * e) * r; o[5] = (a * d - b * b) * r;
    return true;
}
DSC_D D3 ba_sym3_mul(const double* s, D3 v) {
    return d3(s[0] * v.x + s[1] * v.y + s[2] * v.z, s[1] * v.x + s[3] * v.y + s[4] * v.z, s[2] * v.x + s[4] * v.y + s[5] * v.z);
}

// residual, robust weight and the two Jacobians of one edge (g2oTypes.h:160-178, g2oTypes.cc:120-140)
DSC_D void ba_edge(const CamF& cam, const BaPose& T, D3 X, float u, float v, double isg, double delta, double& rho0, double& w,
                   double* e, double* Jx /*2x3*/, double* Jp /*2x6*/, double& chi2, double& zc) {
    F3 xcf;
    reproj_residual(cam, T.R, T.t, X, u, v, e[0], e[1], xcf);
    const D3 Xc0 = mul(T.R, X);
    const D3 Xc = d3(Xc0.x + T.t[0], Xc0.y + T.t[1], Xc0.z + T.t[2]);
    zc = Xc.z;
    chi2 = isg * (e[0] * e[0] + e[1] * e[1]);
    double rho1 = 1.0;
    rho0 = chi2;
    if (delta > 0.0) huber(chi2, delta, rho0, rho1);
    w = rho1 * isg;
    float Jf[6];
    cam_project_jac(cam, xcf, Jf);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double j0 = -(double)Jf[r * 3], j1 = -(double)Jf[r * 3 + 1], j2 = -(double)Jf[r * 3 + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c) Jx[r * 3 + c] = j0 * T.R[c] + j1 * T.R[3 + c] + j2 * T.R[6 + c];
        Jp[r * 6 + 0] = -j1 * Xc.z + j2 * Xc.y;          // [-[Xc]x | I] (g2oTypes.cc:134-138)
        Jp[r * 6 + 1] = j0 * Xc.z - j2 * Xc.x;
        Jp[r * 6 + 2] = -j0 * Xc.y + j1 * Xc.x;
        Jp[r * 6 + 3] = j0; Jp[r * 6 + 4] = j1; Jp[r * 6 + 5] = j2;
    }
}

// per point: Hll, bl; per observation: A, g, W.  part[grid][2] = {sum of robust chi2, max diagonal of Hll}
__global__ void __launch_bounds__(kThreads)
ba_linearize_kernel(int M, const int* __restrict__ pt_ptr, const int* __restrict__ ob_pose, const float2* __restrict__ ob_uv,
                    const float* __restrict__ ob_isg, const unsigned char* __restrict__ ob_act, const int* __restrict__ ob_slot,
                    const double4* __restrict__ X, const BaPose* __restrict__ poses, const CamF* __restrict__ cams,
                    const unsigned char* __restrict__ pose_free, double delta, int points_fixed, size_t O, double* __restrict__ Hll, double* __restrict__ bl, double* __restrict__ W,
                    double* __restrict__ A, double* __restrict__ g, double* __restrict__ part) {
    __shared__ double sm[2 * (kThreads / 32)];
    double acc[1] = {0.0};
    double mx = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const double4 x4 = X[j];
        const D3 Xw = d3(x4.x, x4.y, x4.z);
        double h[6] = {0, 0, 0, 0, 0, 0}, b3[3] = {0, 0, 0};
        for (int o = pt_ptr[j]; o < pt_ptr[j + 1]; ++o) {
            const size_t sl = (size_t)ob_slot[o];
            double* Wo = W + sl;                        // component c at Wo[c * O]
            double* Ao = A + sl;
            double* go = g + sl;
            const int k = ob_pose[o];
            const bool live = ob_act[o] != 0;
            const bool fr = live && pose_free[k] != 0;
            double e[2] = {0.0, 0.0}, Jx[6] = {0, 0, 0, 0, 0, 0}, Jp[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, rho0 = 0.0, w = 0.0, chi2, zc;
            if (live) {
                const float2 uv = ob_uv[o];
                ba_edge(cams[k], poses[k], Xw, uv.x, uv.y, (double)ob_isg[o], delta, rho0, w, e, Jx, Jp, chi2, zc);
                acc[0] += rho0;
                if (!points_fixed) {
                    int q = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
#pragma unroll
                        for (int c = r; c < 3; ++c) h[q++] += w * (Jx[r] * Jx[c] + Jx[3 + r] * Jx[3 + c]);
                        b3[r] -= w * (Jx[r] * e[0] + Jx[3 + r] * e[1]);
                    }
                }
            }
            int q = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int c = r; c < 6; ++c) Ao[(size_t)(q++) * O] = fr ? w * (Jp[r] * Jp[c] + Jp[6 + r] * Jp[6 + c]) : 0.0;
                go[(size_t)r * O] = fr ? -w * (Jp[r] * e[0] + Jp[6 + r] * e[1]) : 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) Wo[(size_t)(r * 3 + c) * O] = fr && !points_fixed ? w * (Jp[r] * Jx[c] + Jp[6 + r] * Jx[3 + c]) : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) Hll[6 * (size_t)j + k] = h[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) bl[3 * (size_t)j + k] = b3[k];
        mx = fmax(mx, fmax(h[0], fmax(h[3], h[5])));
    }
    block_reduce<1>(acc, sm);
    // max over the block (fixed order is irrelevant for a maximum)
    __shared__ double smx[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) smx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) m = fmax(m, smx[k]);
        part[2 * blockIdx.x] = acc[0];
        part[2 * blockIdx.x + 1] = m;
    }
}

// one block per chunk of entries (en_a / en_b: storage slots of the two observations); part[chunk][kBaDiag]
__global__ void __launch_bounds__(kThreads)
ba_schur_kernel(const BaEntryChunk* __restrict__ chunks, const int* __restrict__ en_a, const int* __restrict__ en_b,
                const int* __restrict__ en_pt, const double* __restrict__ Hll, const double* __restrict__ bl,
                const double* __restrict__ W, const double* __restrict__ A, const double* __restrict__ g, double lambda,
                int with_points, size_t O, double* __restrict__ part) {
    __shared__ double sm[kBaDiag * (kThreads / 32)];
    const BaEntryChunk ch = chunks[blockIdx.x];
    double acc[kBaDiag];
#pragma unroll
    for (int k = 0; k < kBaDiag; ++k) acc[k] = 0.0;
    for (int t = ch.begin + threadIdx.x; t < ch.end; t += blockDim.x) {
        const int oa = en_a[t], ob = en_b[t], j = en_pt[t];
        double Hi[6] = {0, 0, 0, 0, 0, 0};
        if (with_points) ba_inv3(Hll + 6 * (size_t)j, lambda, Hi);
        double Wa[18];
#pragma unroll
        for (int c = 0; c < 18; ++c) Wa[c] = W[(size_t)c * O + oa];
        double WH[18];                                   // W_a Hinv (6x3)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const D3 v = ba_sym3_mul(Hi, d3(Wa[r * 3], Wa[r * 3 + 1], Wa[r * 3 + 2]));
            WH[r * 3] = v.x; WH[r * 3 + 1] = v.y; WH[r * 3 + 2] = v.z;
        }
        if (ch.diag) {
#pragma unroll
            for (int k = 0; k < 21; ++k) acc[k] += A[(size_t)k * O + oa];
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[21 + k] += g[(size_t)k * O + oa];
            int q = 27;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = r; c < 6; ++c) acc[q++] += WH[r * 3] * Wa[c * 3] + WH[r * 3 + 1] * Wa[c * 3 + 1] + WH[r * 3 + 2] * Wa[c * 3 + 2];
            const double* b3 = bl + 3 * (size_t)j;
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[48 + r] += WH[r * 3] * b3[0] + WH[r * 3 + 1] * b3[1] + WH[r * 3 + 2] * b3[2];
        } else {
            double Wb[18];
#pragma unroll
            for (int c = 0; c < 18; ++c) Wb[c] = W[(size_t)c * O + ob];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[r * 6 + c] += WH[r * 3] * Wb[c * 3] + WH[r * 3 + 1] * Wb[c * 3 + 1] + WH[r * 3 + 2] * Wb[c * 3 + 2];
        }
    }
    block_reduce<kBaDiag>(acc, sm);
    if (threadIdx.x == 0)
        for (int k = 0; k < kBaDiag; ++k) part[(size_t)kBaDiag * blockIdx.x + k] = acc[k];
}

// dx_l = (Hll + lambda)^-1 (bl - sum_o W_o^T dp[pose(o)]);  Xt = X + dx_l;  and, in the same pass, the trial's cost at the
// trial poses: part[grid][2] = {sum dx_l . (lambda dx_l + bl), activeRobustChi2(Xt, poses_t)}
__global__ void __launch_bounds__(kThreads)
ba_backsub_kernel(int M, const int* __restrict__ pt_ptr, const int* __restrict__ ob_pose, const int* __restrict__ ob_slot, const double4* __restrict__ X,
                  const double* __restrict__ Hll, const double* __restrict__ bl, const double* __restrict__ W,
                  const double* __restrict__ dP /*[K][6]*/, double lambda, int points_fixed, size_t O, double4* __restrict__ Xt,
                  const float2* __restrict__ ob_uv, const float* __restrict__ ob_isg, const unsigned char* __restrict__ ob_act,
                  const BaPose* __restrict__ poses_t, const CamF* __restrict__ cams, double delta, double* __restrict__ part) {
    __shared__ double sm[2 * (kThreads / 32)];
    double acc[2] = {0.0, 0.0};
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const double4 x4 = X[j];
        D3 dx = d3(0, 0, 0);
        if (!points_fixed) {
            D3 r = d3(bl[3 * (size_t)j], bl[3 * (size_t)j + 1], bl[3 * (size_t)j + 2]);
            const D3 b3 = r;
            for (int o = pt_ptr[j]; o < pt_ptr[j + 1]; ++o) {
                const double* Wo = W + (size_t)ob_slot[o];
                const double* dp = dP + 6 * (size_t)ob_pose[o];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    r.x -= Wo[(size_t)(q * 3) * O] * dp[q]; r.y -= Wo[(size_t)(q * 3 + 1) * O] * dp[q]; r.z -= Wo[(size_t)(q * 3 + 2) * O] * dp[q];
                }
            }
            double Hi[6];
            ba_inv3(Hll + 6 * (size_t)j, lambda, Hi);
            dx = ba_sym3_mul(Hi, r);
            acc[0] += dx.x * (lambda * dx.x + b3.x) + dx.y * (lambda * dx.y + b3.y) + dx.z * (lambda * dx.z + b3.z);
        }
        const D3 Xn = d3(x4.x + dx.x, x4.y + dx.y, x4.z + dx.z);
        Xt[j] = make_double4(Xn.x, Xn.y, Xn.z, 0.0);
        for (int o = pt_ptr[j]; o < pt_ptr[j + 1]; ++o) {
            if (!ob_act[o]) continue;
            const int k = ob_pose[o];
            const float2 uv = ob_uv[o];
            double e0, e1;
            F3 xcf;
            reproj_residual(cams[k], poses_t[k].R, poses_t[k].t, Xn, uv.x, uv.y, e0, e1, xcf);
            const double chi2 = (double)ob_isg[o] * (e0 * e0 + e1 * e1);
            double rho0 = chi2, rho1;
            if (delta > 0.0) huber(chi2, delta, rho0, rho1);
            acc[1] += rho0;
        }
    }
    block_reduce<2>(acc, sm);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc[0]; part[2 * blockIdx.x + 1] = acc[1]; }
}

// activeRobustChi2 of a state: part[grid]
__global__ void __launch_bounds__(kThreads)
ba_cost_kernel(int M, const int* __restrict__ pt_ptr, const int* __restrict__ ob_pose, const float2* __restrict__ ob_uv,
               const float* __restrict__ ob_isg, const unsigned char* __restrict__ ob_act, const double4* __restrict__ X,
               const BaPose* __restrict__ poses, const CamF* __restrict__ cams, double delta, double* __restrict__ part) {
    __shared__ double sm[kThreads / 32];
    double acc[1] = {0.0};
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const double4 x4 = X[j];
        const D3 Xw = d3(x4.x, x4.y, x4.z);
        for (int o = pt_ptr[j]; o < pt_ptr[j + 1]; ++o) {
            if (!ob_act[o]) continue;
            const int k = ob_pose[o];
            const float2 uv = ob_uv[o];
            double e0, e1;
            F3 xcf;
            reproj_residual(cams[k], poses[k].R, poses[k].t, Xw, uv.x, uv.y, e0, e1, xcf);
            const double chi2 = (double)ob_isg[o] * (e0 * e0 + e1 * e1);
            double rho0 = chi2, rho1;
            if (delta > 0.0) huber(chi2, delta, rho0, rho1);
            acc[0] += rho0;
        }
    }
    block_reduce<1>(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
}

// e->chi2() and e->isDepthPositive() of every edge (caller order through ob_orig)
__global__ void __launch_bounds__(kThreads)
ba_edge_kernel(int M, const int* __restrict__ pt_ptr, const int* __restrict__ ob_pose, const float2* __restrict__ ob_uv,
               const float* __restrict__ ob_isg, const int* __restrict__ ob_orig, const double4* __restrict__ X,
               const BaPose* __restrict__ poses, const CamF* __restrict__ cams, double* __restrict__ chi2_out,
               unsigned char* __restrict__ pos_out) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const double4 x4 = X[j];
        const D3 Xw = d3(x4.x, x4.y, x4.z);
        for (int o = pt_ptr[j]; o < pt_ptr[j + 1]; ++o) {
            const int k = ob_pose[o];
            const float2 uv = ob_uv[o];
            double e0, e1;
            F3 xcf;
            reproj_residual(cams[k], poses[k].R, poses[k].t, Xw, uv.x, uv.y, e0, e1, xcf);
            const D3 Xc = mul(poses[k].R, Xw);
            chi2_out[ob_orig[o]] = (double)ob_isg[o] * (e0 * e0 + e1 * e1);
            pos_out[ob_orig[o]] = (Xc.z + poses[k].t[2]) > 0.0 ? 1 : 0;
        }
    }
}

}  // namespace dsc
