// dsc_kernels_ell.cuh -- the per-iteration gather kernels (rotations K7, cost K6, linearise K2+K3) on the same
// layout as the PCG operator: sliced ELL (slice = 32 rows = one warp, lane = row, sums in registers, no cross-lane
// reduction per row) with the tile's points and rotations staged in shared memory, so that a neighbour inside the
// tile (one degree-sorted group of kSortGroup Morton-consecutive rows) is read from shared memory and only halo
// neighbours are gathered from L2.  Padding slots of the ELL point at the row itself and are skipped.
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

constexpr size_t kWinBytes = sizeof(double4) * 3 * kSortGroup;      // X1 | X2 | Q windows
constexpr int kEllThreads = 256;
constexpr int kLinThreads = 256;             // linearise: 234 registers, one block per SM (384 / 512 threads spill and are slower)

struct Window { const double4* x1; const double4* x2; const double4* q; int v0, nv; };

// kRO: load policy of the points (dsc_kernels.cuh, ld256); the rotations Q are constant during a refinement in both forms
template <bool kRO = true>
DSC_D void stage_window(double4* sw, const double* __restrict__ P, const double* __restrict__ Q, int n, int v0, int nv, bool with_q) {
    const double4* p = reinterpret_cast<const double4*>(P);
    const double4* q = reinterpret_cast<const double4*>(Q);
    for (int k = threadIdx.x; k < nv; k += blockDim.x) {
        sw[k] = ld256<kRO>(p + (size_t)v0 + k);
        sw[kSortGroup + k] = ld256<kRO>(p + (size_t)n + v0 + k);
        if (with_q) sw[2 * kSortGroup + k] = ldg256(q + (size_t)v0 + k);
    }
}
template <bool kRO = true>
DSC_D void fetch_point(const Window& w, const double* __restrict__ P, int n, int j, P8& Pj) {
    const unsigned jl = (unsigned)(j - w.v0);
    double4 a, b;
    if (jl < (unsigned)w.nv) { a = w.x1[jl]; b = w.x2[jl]; }
    else {
        a = ld256<kRO>(reinterpret_cast<const double4*>(P) + (size_t)j);
        b = ld256<kRO>(reinterpret_cast<const double4*>(P) + (size_t)n + (size_t)j);
    }
    Pj.a = d3(a.x, a.y, a.z); Pj.b = d3(b.x, b.y, b.z);
}
DSC_D void fetch_quat(const Window& w, const double* __restrict__ Q, int j, double* qj) {
    const unsigned jl = (unsigned)(j - w.v0);
    const double4 u = jl < (unsigned)w.nv ? w.q[jl] : ldg256(reinterpret_cast<const double4*>(Q) + (size_t)j);
    qj[0] = u.x; qj[1] = u.y; qj[2] = u.z; qj[3] = u.w;
}
// Column index and weight of a row's ELL slots, fetched one column ahead of their use so that the load latency
// overlaps the previous edge's arithmetic.
struct SlotStream {
    const int* ecol; const double* ewgt; int b1, lane, jn; double wn;
    DSC_D SlotStream(const int* __restrict__ ec, const double* __restrict__ ew, int b0, int b1_, int lane_)
        : ecol(ec), ewgt(ew), b1(b1_), lane(lane_), jn(0), wn(0.0) {
        if (b0 < b1) { jn = __ldg(ecol + (size_t)b0 * 32 + lane); wn = __ldg(ewgt + (size_t)b0 * 32 + lane); }
    }
    DSC_D void next(int bk, int& j, double& w) {
        j = jn; w = wn;
        if (bk + 1 < b1) { jn = __ldg(ecol + (size_t)(bk + 1) * 32 + lane); wn = __ldg(ewgt + (size_t)(bk + 1) * 32 + lane); }
    }
};

// warp sum of K per-lane values into a per-warp shared accumulator row (lane 0 adds)
template <int K>
DSC_D void warp_accumulate(const double (&v)[K], double* wacc) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) wacc[k] += s;
    }
}

// ------------------------------------------------------------------ K7 computeR (Geometry.cc:549-604)
#ifndef DSC_ROT_BLOCKS
#define DSC_ROT_BLOCKS 3      // 78 registers: three blocks per SM (measured: 204 -> 162 us at 1M; four spill)
#endif
#ifndef DSC_COST_BLOCKS
#define DSC_COST_BLOCKS 2     // (three: 185 -> 169 us at 1M but 28 -> 31 us at 100k; four spill)
#endif
__global__ void __launch_bounds__(kEllThreads, DSC_ROT_BLOCKS)
rotations_ell_kernel(int n, const double* __restrict__ P, const int* __restrict__ sliceptr, const int* __restrict__ ecol,
                     const double* __restrict__ ewgt, double* __restrict__ Q) {
    extern __shared__ double4 sw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int ntiles = (n + kSortGroup - 1) / kSortGroup;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int v0 = tile * kSortGroup, nv = min(kSortGroup, n - v0);
        __syncthreads();
        stage_window(sw, P, P, n, v0, nv, false);
        __syncthreads();
        const Window win{sw, sw + kSortGroup, sw + 2 * kSortGroup, v0, nv};
        for (int ls = warp; ls * 32 < nv; ls += wpb) {
            const int il = ls * 32 + lane, i = v0 + il;
            if (il >= nv) continue;
            const P8 Pi{d3(win.x1[il].x, win.x1[il].y, win.x1[il].z), d3(win.x2[il].x, win.x2[il].y, win.x2[il].z)};
            const int sl = (v0 >> 5) + ls;
            const int b0 = __ldg(sliceptr + sl), b1 = __ldg(sliceptr + sl + 1);
            double S[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) S[k] = 0.0;
            int deg = 0;
            SlotStream slots(ecol, ewgt, b0, b1, lane);
            for (int bk = b0; bk < b1; ++bk) {
                int j; double wv;
                slots.next(bk, j, wv);
                if (j == i) continue;
                P8 Pj;
                fetch_point(win, P, n, j, Pj);
                const D3 d1 = Pi.a - Pj.a, d2 = Pi.b - Pj.b;
                S[0] += wv * d1.x * d2.x; S[1] += wv * d1.x * d2.y; S[2] += wv * d1.x * d2.z;
                S[3] += wv * d1.y * d2.x; S[4] += wv * d1.y * d2.y; S[5] += wv * d1.y * d2.z;
                S[6] += wv * d1.z * d2.x; S[7] += wv * d1.z * d2.y; S[8] += wv * d1.z * d2.z;
                ++deg;
            }
            double q[4] = {0.0, 0.0, 0.0, 1.0};
            if (deg > 0) {
                double R[9];
                rotation_from_covariance(S, R);
                rot_to_quat(R, q);
            }
            reinterpret_cast<double4*>(Q)[i] = make_double4(q[0], q[1], q[2], q[3]);
        }
    }
}

// unary residuals of one observation: reprojection (Huber) and depth; returns rho0 and the depth chi2
DSC_D void unary_cost(const CamF& cam, const double* R, const double* t, D3 X, float u, float v, double isg, double dmeas,
                      double scale, const WeightsDev& W, double& rho0, double& chid) {
    double e0, e1, r1;
    F3 xcf;
    reproj_residual(cam, R, t, X, u, v, e0, e1, xcf);
    huber(isg * W.rep * (e0 * e0 + e1 * e1), W.huber, rho0, r1);
    const D3 xc = mul(R, X);
    const double rr = dmeas / scale - (xc.z + t[2]);
    const double ed = rr * rr * (scale <= 0.0 ? 500.0 : 1.0);
    chid = W.depth_info * ed * ed;
}

// ------------------------------------------------------------------ K6 cost: part[grid][3]
// The tiles first, first + stride, ... of one block; acc[3] = this thread's share of (reprojection, depth, ARAP) chi2.
// sw = the block's dynamic shared window (kWinBytes).
template <bool kRO>
DSC_D void cost_tiles(int first, int stride, int n, const double* __restrict__ P, const double* __restrict__ Q,
                      const float4* __restrict__ uv, const double2* __restrict__ dm, const float2* __restrict__ isg,
                      const int* __restrict__ sliceptr, const int* __restrict__ ecol, const double* __restrict__ ewgt,
                      const Globals& G, const PairDev& pr, const WeightsDev& W, double4* sw, double (&acc)[3],
                      int tile0 = 0, int tile1 = -1 /* tiles [tile0, tile1): the rank's share of a point-sharded pair */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int ntiles = tile1 < 0 ? (n + kSortGroup - 1) / kSortGroup : tile1;
    for (int tile = tile0 + first; tile < ntiles; tile += stride) {
        const int v0 = tile * kSortGroup, nv = min(kSortGroup, n - v0);
        __syncthreads();
        stage_window<kRO>(sw, P, Q, n, v0, nv, true);
        __syncthreads();
        const Window win{sw, sw + kSortGroup, sw + 2 * kSortGroup, v0, nv};
        for (int ls = warp; ls * 32 < nv; ls += wpb) {
            const int il = ls * 32 + lane, i = v0 + il;
            if (il >= nv) continue;
            const P8 Pi{d3(win.x1[il].x, win.x1[il].y, win.x1[il].z), d3(win.x2[il].x, win.x2[il].y, win.x2[il].z)};
            const double qi[4] = {win.q[il].x, win.q[il].y, win.q[il].z, win.q[il].w};
            const int sl = (v0 >> 5) + ls;
            const int b0 = __ldg(sliceptr + sl), b1 = __ldg(sliceptr + sl + 1);
            double ea = 0.0;
            SlotStream slots(ecol, ewgt, b0, b1, lane);
            for (int bk = b0; bk < b1; ++bk) {
                int j; double wv;
                slots.next(bk, j, wv);
                if (j == i) continue;
                P8 Pj;
                double qj[4];
                fetch_point<kRO>(win, P, n, j, Pj);
                fetch_quat(win, Q, j, qj);
                ArapGrad g;
                arap_edge<false>(Pi, Pj, qi, qj, wv, W.inv_area, G, g);
                ea += g.e * g.e;
            }
            acc[2] += W.arap_info * ea;
            const float4 o = uv[i];
            const float2 sg = isg[i];
            const double2 d = dm[i];
            double r0, cd;
            unary_cost(pr.cam1, pr.R1, pr.t1, Pi.a, o.x, o.y, (double)sg.x, d.x, G.s1, W, r0, cd);
            acc[0] += r0; acc[1] += cd;
            unary_cost(pr.cam2, pr.R2, pr.t2, Pi.b, o.z, o.w, (double)sg.y, d.y, G.s2, W, r0, cd);
            acc[0] += r0; acc[1] += cd;
        }
    }
    __syncthreads();
}
__global__ void __launch_bounds__(kEllThreads, DSC_COST_BLOCKS)
cost_ell_kernel(int n, const double* __restrict__ P, const double* __restrict__ Q, const float4* __restrict__ uv,
                const double2* __restrict__ dm, const float2* __restrict__ isg, const int* __restrict__ sliceptr,
                const int* __restrict__ ecol, const double* __restrict__ ewgt, const Globals* __restrict__ Gp,
                const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W, double* __restrict__ part,
                int tile0, int tile1) {
    extern __shared__ double4 sw[];
    __shared__ double sm[3 * (kEllThreads / 32)];
    __shared__ Globals G;
    if (threadIdx.x == 0) G = *Gp;
    double acc[3] = {0.0, 0.0, 0.0};
    cost_tiles<true>(blockIdx.x, gridDim.x, n, P, Q, uv, dm, isg, sliceptr, ecol, ewgt, G, pr, W, sw, acc, tile0, tile1);
    block_reduce<3>(acc, sm);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = acc[0]; part[3 * blockIdx.x + 1] = acc[1]; part[3 * blockIdx.x + 2] = acc[2]; }
}

// ------------------------------------------------------------------ K2 + K3 linearise and assemble
// Per row: gradient b (6), packed 6x6 diagonal block D (21), unary record U (16), ARAP Jacobian records Je of its
// ELL slots (coalesced).  Global rows: T_g gradient from the per-row sum Bg_i = sum_j 2 W e_ij g_ij (the directed
// twins carry the same e and g):  b_w = -2 sum_i X1_i x Bg_i,  b_v = 2 sum_i Bg_i;  C_TT = sum_dir W gT gT^T.
// part[grid][kLinPart]: chi2[3], max diag, bg[8], C_TT packed (21), C_ss (2)  -- reduced by finalize_linearize_kernel.
// kDual: the fp32 mode of the solver (dsc_set_precision) -- the records the PCG streams are ALSO written as float
// (UF, JeF: same layouts, 4-byte values); the double copies serve the fp64 residual of the iterative refinement.
template <bool kRO, bool kDual = false>
DSC_D void linearize_tiles(int first, int stride, int n, const double* __restrict__ P, const double* __restrict__ Q,
                           const float4* __restrict__ uv, const double2* __restrict__ dm, const float2* __restrict__ isg,
                           const int* __restrict__ sliceptr, const int* __restrict__ ecol, const double* __restrict__ ewgt,
                           const Globals& G, const PairDev& pr, const WeightsDev& W,
                           double* __restrict__ b, double* __restrict__ D, double* __restrict__ U, double* __restrict__ Je,
                           double* __restrict__ part_row /* [kLinPart] of this block */, double4* sw,
                           float* __restrict__ UF = nullptr, float* __restrict__ JeF = nullptr, int tile0 = 0, int tile1 = -1) {
    __shared__ double wacc[kLinThreads / 32][kLinPart];
    __shared__ double wmax[kLinThreads / 32];
    __syncthreads();                                   // (a previous phase of the same block may still read the scratch)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int k = lane; k < kLinPart; k += 32) wacc[warp][k] = 0.0;
    if (lane == 0) wmax[warp] = 0.0;
    const int ntiles = tile1 < 0 ? (n + kSortGroup - 1) / kSortGroup : tile1;
    for (int tile = tile0 + first; tile < ntiles; tile += stride) {
        const int v0 = tile * kSortGroup, nv = min(kSortGroup, n - v0);
        __syncthreads();
        stage_window<kRO>(sw, P, Q, n, v0, nv, true);
        __syncthreads();
        const Window win{sw, sw + kSortGroup, sw + 2 * kSortGroup, v0, nv};
        for (int ls = warp; ls * 32 < nv; ls += wpb) {
            const int il = ls * 32 + lane, i = v0 + il;
            const bool act = il < nv;
            const int ilc = act ? il : nv - 1;
            const P8 Pi{d3(win.x1[ilc].x, win.x1[ilc].y, win.x1[ilc].z), d3(win.x2[ilc].x, win.x2[ilc].y, win.x2[ilc].z)};
            const double qi[4] = {win.q[ilc].x, win.q[ilc].y, win.q[ilc].z, win.q[ilc].w};
            const int sl = (v0 >> 5) + ls;
            const int b0 = __ldg(sliceptr + sl), b1 = __ldg(sliceptr + sl + 1);
            double gb[6], Dk[21], cT[21], Bg[3] = {0.0, 0.0, 0.0}, chi_a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) gb[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 21; ++k) { Dk[k] = 0.0; cT[k] = 0.0; }
            SlotStream slots(ecol, ewgt, b0, b1, lane);
            for (int bk = b0; bk < b1; ++bk) {
                int j; double wv;
                slots.next(bk, j, wv);
                if (!act || j == i) continue;                  // padding slot: its Je record stays all-zero
                P8 Pj;
                double qj[4];
                fetch_point<kRO>(win, P, n, j, Pj);
                fetch_quat(win, Q, j, qj);
                ArapGrad g;
                arap_edge<true>(Pi, Pj, qi, qj, wv, W.inv_area, G, g);
                double* jb = Je + (size_t)bk * 288 + lane;
                jb[0] = g.u.x; jb[32] = g.u.y; jb[64] = g.u.z; jb[96] = g.m.x; jb[128] = g.m.y; jb[160] = g.m.z;
                jb[192] = g.g.x; jb[224] = g.g.y; jb[256] = g.g.z;
                if (kDual) {
                    float* jf = JeF + (size_t)bk * 288 + lane;
                    jf[0] = (float)g.u.x; jf[32] = (float)g.u.y; jf[64] = (float)g.u.z; jf[96] = (float)g.m.x; jf[128] = (float)g.m.y;
                    jf[160] = (float)g.m.z; jf[192] = (float)g.g.x; jf[224] = (float)g.g.y; jf[256] = (float)g.g.z;
                }
                const double gi[6] = {g.gi1.x, g.gi1.y, g.gi1.z, g.gi2.x, g.gi2.y, g.gi2.z};
                const double gt[6] = {g.gw.x, g.gw.y, g.gw.z, g.gv.x, g.gv.y, g.gv.z};
                const double we = W.arap_info * g.e;
                chi_a += we * g.e;
                Bg[0] += 2.0 * we * g.g.x; Bg[1] += 2.0 * we * g.g.y; Bg[2] += 2.0 * we * g.g.z;
#pragma unroll
                for (int r = 0; r < 6; ++r) {                  // vertex rows: the directed twin doubles them
                    gb[r] -= 2.0 * we * gi[r];
                    const double s2 = 2.0 * W.arap_info * gi[r];
                    const double st = W.arap_info * gt[r];
#pragma unroll
                    for (int c = r; c < 6; ++c) { Dk[pk<6>(r, c)] += s2 * gi[c]; cT[pk<6>(r, c)] += st * gt[c]; }
                }
            }
            double glob[kLinPart];
#pragma unroll
            for (int k = 0; k < kLinPart; ++k) glob[k] = 0.0;
            double mx = 0.0;
            if (act) {
                const float4 o = uv[i];
                const float2 sg = isg[i];
                const double2 d = dm[i];
                double Urec[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) Urec[k] = 0.0;
#pragma unroll
                for (int cam = 0; cam < 2; ++cam) {
                    const CamF& cm = cam == 0 ? pr.cam1 : pr.cam2;
                    const double* R = cam == 0 ? pr.R1 : pr.R2;
                    const double* t = cam == 0 ? pr.t1 : pr.t2;
                    const D3 X = cam == 0 ? Pi.a : Pi.b;
                    double e0, e1r, rho0, rho1;
                    F3 xcf;
                    reproj_residual(cm, R, t, X, cam == 0 ? o.x : o.z, cam == 0 ? o.y : o.w, e0, e1r, xcf);
                    const double om = (double)(cam == 0 ? sg.x : sg.y) * W.rep;
                    huber(om * (e0 * e0 + e1r * e1r), W.huber, rho0, rho1);
                    glob[0] += rho0;
                    float Jf[6];
                    cam_project_jac(cm, xcf, Jf);
                    double J[6];                               // J = -Jproj * R   (2x3)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            J[r * 3 + c] = -((double)Jf[r * 3] * R[c] + (double)Jf[r * 3 + 1] * R[3 + c] + (double)Jf[r * 3 + 2] * R[6 + c]);
                    const double wr = rho1 * om;
                    double Uu[6];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        gb[cam * 3 + r] -= wr * (J[r] * e0 + J[3 + r] * e1r);
#pragma unroll
                        for (int c = r; c < 3; ++c) Uu[pk<3>(r, c)] = wr * (J[r] * J[c] + J[3 + r] * J[3 + c]);
                    }
                    // depth edge EdgeDepthCorrection (g2oTypes.h:400-416): e = (d/s - z_c)^2 (x500 if s<=0)
                    const double s = cam == 0 ? G.s1 : G.s2;
                    const double dd = cam == 0 ? d.x : d.y;
                    const double kf = s <= 0.0 ? 500.0 : 1.0;
                    const D3 xc = mul(R, X);
                    const double rr = dd / s - (xc.z + t[2]);
                    const double ed = kf * rr * rr;
                    const double alpha = -2.0 * kf * rr;       // de/dX = alpha * R[2,:]
                    const double Js = 2.0 * kf * rr * (-dd / (s * s));
                    glob[1] += W.depth_info * ed * ed;
                    const double nrm[3] = {R[6], R[7], R[8]};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        gb[cam * 3 + r] -= W.depth_info * ed * alpha * nrm[r];
#pragma unroll
                        for (int c = r; c < 3; ++c) Uu[pk<3>(r, c)] += W.depth_info * alpha * alpha * nrm[r] * nrm[c];
                    }
                    glob[10 + cam] -= W.depth_info * ed * Js;          // bg[6 + cam]
                    glob[33 + cam] += W.depth_info * Js * Js;          // C[s s]
                    Urec[12 + cam] = W.depth_info * alpha * Js;        // kd: coupling X <-> s along R[2,:]
#pragma unroll
                    for (int k = 0; k < 6; ++k) Urec[cam * 6 + k] = Uu[k];
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = r; c < 3; ++c) Dk[pk<6>(cam * 3 + r, cam * 3 + c)] += Uu[pk<3>(r, c)];
                }
                store6(b, i, d3(gb[0], gb[1], gb[2]), d3(gb[3], gb[4], gb[5]));
                double* Dp = blk21(D, i);
#pragma unroll
                for (int k = 0; k < 21; ++k) Dp[k * 32] = Dk[k];
                double* Up = U + ((size_t)(i >> 5) * kURec) * 32 + (i & 31);     // slice-major: component k of 32 rows is one 256 B line pair
#pragma unroll
                for (int k = 0; k < kURec; ++k) Up[k * 32] = Urec[k];
                if (kDual) {
                    float* Uf = UF + ((size_t)(i >> 5) * kURec) * 32 + (i & 31);
#pragma unroll
                    for (int k = 0; k < kURec; ++k) Uf[k * 32] = (float)Urec[k];
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) mx = fmax(mx, fabs(Dk[pk<6>(r, r)]));
                glob[2] = chi_a;
                const D3 bgv = d3(Bg[0], Bg[1], Bg[2]);
                const D3 cx = cross(Pi.a, bgv);
                glob[4] = -2.0 * cx.x; glob[5] = -2.0 * cx.y; glob[6] = -2.0 * cx.z;
                glob[7] = 2.0 * Bg[0]; glob[8] = 2.0 * Bg[1]; glob[9] = 2.0 * Bg[2];
#pragma unroll
                for (int k = 0; k < 21; ++k) glob[12 + k] = cT[k];
            }
            warp_accumulate<kLinPart>(glob, wacc[warp]);
            for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
            if (lane == 0) wmax[warp] = fmax(wmax[warp], mx);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double* o = part_row;
        for (int k = 0; k < kLinPart; ++k) {
            double s = 0.0;
            for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += wacc[wv][k];
            o[k] = s;
        }
        double m = 0.0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) m = fmax(m, wmax[wv]);
        o[3] = m;
    }
}

template <bool kDual>
__global__ void __launch_bounds__(kLinThreads, 1)
linearize_ell_kernel(int n, const double* __restrict__ P, const double* __restrict__ Q, const float4* __restrict__ uv,
                     const double2* __restrict__ dm, const float2* __restrict__ isg, const int* __restrict__ sliceptr,
                     const int* __restrict__ ecol, const double* __restrict__ ewgt, const Globals* __restrict__ Gp,
                     const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W,
                     double* __restrict__ b, double* __restrict__ D, double* __restrict__ U, double* __restrict__ Je,
                     double* __restrict__ part, float* __restrict__ UF, float* __restrict__ JeF, int tile0, int tile1) {
    extern __shared__ double4 sw[];
    __shared__ Globals G;
    if (threadIdx.x == 0) G = *Gp;
    linearize_tiles<true, kDual>(blockIdx.x, gridDim.x, n, P, Q, uv, dm, isg, sliceptr, ecol, ewgt, G, pr, W, b, D, U, Je,
                                 part + (size_t)kLinPart * blockIdx.x, sw, UF, JeF, tile0, tile1);
}

}  // namespace dsc
