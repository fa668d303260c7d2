// dsc_small.cuh -- the whole PCG solve of a SMALL problem by ONE thread-block cluster: all iterations inside a launch,
// the CTAs meeting at the hardware cluster barrier twice per iteration instead of the host launching two kernels per
// iteration.  Between the dense path (<= 600 correspondences) and ~16 k correspondences -- the sizes of the reference's
// real sequences and of config 5's pairs -- a PCG iteration of the large-problem kernels is ~15 us of pure launch /
// grid-drain latency for ~1 us of work; here it costs two cluster barriers.
// Same algorithm and the same arithmetic as cg_init / cg_update / cg_spmv (Chronopoulos-Gear PCG, block-Jacobi 6x6 +
// exact 8x8 preconditioner, the ARAP operator applied from the per-edge Jacobian records); everything the cluster
// itself writes during the solve is read back with L2 loads (__ldcg).
// Two users: pcg_cluster_kernel (one solve per launch, driven by the host LM loop of dsc_optimize) and the one-launch
// LM of the batched path (dsc_batch.cuh), where the linearisation of the same launch writes P / Je / U / D / b between
// solves: kRO = false makes those loads L2-coherent too.
#pragma once
#include <cooperative_groups.h>

#include "dsc_kernels.cuh"

namespace dsc {

namespace cg = cooperative_groups;

constexpr int kGridMaxRows = 120000;             // the grid-wide one-launch solve (pcg_grid_kernel) up to this many rows (0 = off; DSC_GRID_MAX_ROWS)
constexpr int kSmallMaxRows = 3072;          // above this the grid-wide variant is faster (measured: 2 k 223 vs 217 LM it/s, 4 k 96 vs 103, 8 k 168 vs 244)
constexpr int kSmallBatch = 4;              // ELL columns whose loads are in flight together
constexpr int kSmallCluster = 16;            // CTAs per cluster (non-portable size; 8 is tried if 16 cannot launch)

DSC_D void load6_l2(const double* V, int i, D3& a, D3& b) { load6_t<false>(V, i, a, b); }
// sum of part[0 .. nb) by every warp on its own: lane c loads entry c (one L2 round trip for all of them,
// a serial loop would pay one per entry), then a fixed butterfly: the same result in every lane of every warp and CTA
DSC_D double sum_small(const double* part, int nb) {
    const int lane = threadIdx.x & 31;
    double s = lane < nb ? __ldcg(part + lane) : 0.0;
    for (int i = lane + 32; i < nb; i += 32) s += __ldcg(part + i);     // (the grid-wide variant has one partial per SM)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// everything one solve touches (one problem)
struct ClusterPcgArgs {
    int n;
    const double* P;            // X1 plane is read by the operator
    const double* Je; const double* U; const int* sliceptr; const int* ecol;
    const double* b; const double* D; const LinGlobal* lin;
    double* Minv; double* Ginv; int* err;
    CgVecs v;
    double *gpart0, *gpart1, *dpart, *bpart;
};
// the scalars of a running solve; identical in every thread of the cluster (all derived from the same partial sums)
struct ClusterPcgState { int k; double gamma0, gprev, aprev; int converged, breakdown; };

// w = (H + lambda I) z and the partial sums of z.w and of the 8 global rows (see cg_spmv_kernel for the formulas).
// Rg: rotation of T_global; pr / W: the pair and the weights (any address space).
template <bool kRO>
DSC_D void cluster_spmv(int rank, int cs, const ClusterPcgArgs& A, const double* Rg, const PairDev& pr, const WeightsDev& W, double lambda) {
    __shared__ double sm[9 * (kThreads / 32)];
    __shared__ double psum[(kThreads / 32) * 9 * 32];
    __shared__ double zgs[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = kThreads / 32;
    const int n = A.n, nslices = (n + 31) / 32;
    const CgVecs& v = A.v;
    __syncthreads();
    if (threadIdx.x < 8) zgs[threadIdx.x] = __ldcg(v.zg + threadIdx.x);
    __syncthreads();
    const D3 zw = d3(zgs[0], zgs[1], zgs[2]);
    const D3 zv2 = d3(2.0 * zgs[3], 2.0 * zgs[4], 2.0 * zgs[5]);
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    // Slices are dealt to the CTAs in contiguous groups of per_cta.  When a CTA has fewer slices than warps, wps
    // warps share a slice: part p walks the column batches p, p + wps, ... and the leader (p = 0) adds the partial
    // sums (through shared memory, in a fixed order) and finishes the rows -- 1000 correspondences are 32 slices
    // for 128 warps, and the phase is as long as its longest warp.
    const int per_cta = (nslices + cs - 1) / cs;
    int wps = 1;
    while (wps * 2 * per_cta <= wpb) wps *= 2;
    const int slot = warp / wps, part = warp % wps;
    for (int s0 = 0; s0 < per_cta; s0 += wpb / wps) {           // (uniform trip count inside the CTA)
        const int sl = rank * per_cta + s0 + slot;
        const bool have = s0 + slot < per_cta && sl < nslices;
        const int slc = have ? sl : 0;
        const int i = slc * 32 + lane;
        const bool act = have && i < n;
        const int ic = act ? i : n - 1;
        D3 zi1, zi2;
        load6_l2(v.z, ic, zi1, zi2);
        const double4 xi = ld256<kRO>(reinterpret_cast<const double4*>(A.P) + ic);
        const D3 X1i = d3(xi.x, xi.y, xi.z);
        D3 Am = d3(0, 0, 0), Ag = d3(0, 0, 0), Au = d3(0, 0, 0);
        const int b0 = __ldg(A.sliceptr + slc), b1 = have ? __ldg(A.sliceptr + slc + 1) : b0;
        // kSmallBatch ELL columns at a time: all their index / record loads are issued together, then all their
        // gathers -- with 8 warps per SM the loop is bound by L2 latency, not by throughput
        for (int bk = b0 + part * kSmallBatch; bk < b1; bk += wps * kSmallBatch) {
            int jq[kSmallBatch];
            D3 uq[kSmallBatch], mq[kSmallBatch], gq[kSmallBatch];
#pragma unroll
            for (int q = 0; q < kSmallBatch; ++q) {
                const bool ok = bk + q < b1;
                const size_t blk = ok ? (size_t)(bk + q) : (size_t)b0;
                const double* jb = A.Je + blk * 288 + lane;
                jq[q] = ok ? __ldg(A.ecol + blk * 32 + lane) : ic;
                const double f = ok ? 1.0 : 0.0;                       // columns past the end contribute exactly 0
                uq[q] = f * d3(ld64<kRO>(jb), ld64<kRO>(jb + 32), ld64<kRO>(jb + 64));
                mq[q] = f * d3(ld64<kRO>(jb + 96), ld64<kRO>(jb + 128), ld64<kRO>(jb + 160));
                gq[q] = f * d3(ld64<kRO>(jb + 192), ld64<kRO>(jb + 224), ld64<kRO>(jb + 256));
            }
            D3 zq1[kSmallBatch], zq2[kSmallBatch], xq[kSmallBatch];
#pragma unroll
            for (int q = 0; q < kSmallBatch; ++q) {
                load6_l2(v.z, jq[q], zq1[q], zq2[q]);
                const double4 xj = ld256<kRO>(reinterpret_cast<const double4*>(A.P) + (size_t)jq[q]);
                xq[q] = d3(xj.x, xj.y, xj.z);
            }
#pragma unroll
            for (int q = 0; q < kSmallBatch; ++q) {
                const D3 S1 = X1i + xq[q];
                const D3 t = mul(Rg, zi2 + zq2[q]) - (zi1 + zq1[q]) + cross(zw, S1) - zv2;
                const double s = dot(uq[q], zi2 - zq2[q]) - dot(mq[q], zi1 - zq1[q]) + 2.0 * dot(gq[q], t);
                const double w2 = 2.0 * W.arap_info * s;
                Am = Am + w2 * mq[q]; Ag = Ag + w2 * gq[q]; Au = Au + w2 * uq[q];
            }
        }
        if (wps > 1) {                                          // partial sums of the slice's other warps
            double* ps = psum + (size_t)warp * 9 * 32 + lane;
            ps[0] = Am.x; ps[32] = Am.y; ps[64] = Am.z; ps[96] = Ag.x; ps[128] = Ag.y; ps[160] = Ag.z;
            ps[192] = Au.x; ps[224] = Au.y; ps[256] = Au.z;
            __syncthreads();
            if (part == 0)
                for (int q = 1; q < wps; ++q) {
                    const double* pq = psum + (size_t)(warp + q) * 9 * 32 + lane;
                    Am = Am + d3(pq[0], pq[32], pq[64]); Ag = Ag + d3(pq[96], pq[128], pq[160]);
                    Au = Au + d3(pq[192], pq[224], pq[256]);
                }
            __syncthreads();
        }
        if (act && part == 0) {
            const double* Up = A.U + ((size_t)(i >> 5) * kURec) * 32 + lane;
            double uu[kURec];
#pragma unroll
            for (int k = 0; k < kURec; ++k) uu[k] = ld64<kRO>(Up + k * 32);
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
                acc[6 + cam] += kd * nz;
            }
            double dl = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) { out[k] += lambda * zi[k]; dl += zi[k] * out[k]; }
            acc[8] += dl;
            const D3 cx = cross(X1i, Ag);
            acc[0] += 2.0 * cx.x; acc[1] += 2.0 * cx.y; acc[2] += 2.0 * cx.z;
            acc[3] -= 2.0 * Ag.x; acc[4] -= 2.0 * Ag.y; acc[5] -= 2.0 * Ag.z;
            store6(v.w, i, d3(out[0], out[1], out[2]), d3(out[3], out[4], out[5]));
        }
    }
    block_reduce<9>(acc, sm);
    if (threadIdx.x == 0) {
        double dl = acc[8];
        for (int k = 0; k < 8; ++k) { A.bpart[8 * rank + k] = acc[k]; dl += zgs[k] * acc[k]; }
        if (rank == 0)
            for (int k = 0; k < 8; ++k) dl += zgs[k] * ((k >= 6 ? __ldcg(&A.lin->C[k * 8 + k]) : 0.0) + lambda) * zgs[k];
        A.dpart[rank] = dl;
    }
}

// start of a solve: preconditioner, r = b, z = M^-1 r, gamma partial, first operator application
template <bool kRO, typename Group>
DSC_D void cluster_pcg_begin(Group& cluster, const ClusterPcgArgs& A, const double* Rg, const PairDev& pr,
                             const WeightsDev& W, double lambda, ClusterPcgState& st) {
    __shared__ double sm[kThreads / 32];
    const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
    const int n = A.n;
    const CgVecs& v = A.v;
    double g[1] = {0.0};
    for (int i = rank * kThreads + threadIdx.x; i < n; i += cs * kThreads) {
        double r[6], z[6], M[21];
        D3 a, c;
        load6_t<kRO>(A.b, i, a, c);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = c.x; r[4] = c.y; r[5] = c.z;
        precond_block<kRO>(A.D, i, lambda, A.Minv, A.err, M);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int t = 0; t < 6; ++t) sacc += (q <= t ? M[pk<6>(q, t)] : M[pk<6>(t, q)]) * r[t];
            z[q] = sacc;
        }
        store6(v.r, i, a, c);
        store6(v.z, i, d3(z[0], z[1], z[2]), d3(z[3], z[4], z[5]));
#pragma unroll
        for (int t = 0; t < 6; ++t) g[0] += r[t] * z[t];
    }
    if (rank == 0 && threadIdx.x == 0) {
        precond_global<kRO>(A.lin, lambda, A.Ginv, A.err);
        for (int a = 0; a < 8; ++a) {
            double s = 0.0;
            const double bga = ld64<kRO>(&A.lin->bg[a]);
            for (int c = 0; c < 8; ++c) s += A.Ginv[a * 8 + c] * ld64<kRO>(&A.lin->bg[c]);
            v.rg[a] = bga; v.zg[a] = s; v.xg[a] = 0.0; v.pg[a] = 0.0; v.sg[a] = 0.0;
            g[0] += bga * s;
        }
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) A.gpart0[rank] = g[0];
    st.k = 0; st.gamma0 = 0.0; st.gprev = 1.0; st.aprev = 1.0; st.converged = 0; st.breakdown = 0;
    cluster.sync();
    cluster_spmv<kRO>(rank, cs, A, Rg, pr, W, lambda);
    cluster.sync();
}

// iterations: update (Chronopoulos-Gear step), barrier, operator, barrier -- until gamma <= rtol2 gamma0 (converged),
// breakdown, or max_iters updates in total.  Resumable: call again with a tighter rtol2 to continue the same solve.
template <bool kRO, typename Group>
DSC_D void cluster_pcg_run(Group& cluster, const ClusterPcgArgs& A, const double* Rg, const PairDev& pr,
                           const WeightsDev& W, double lambda, double rtol2, int max_iters, ClusterPcgState& st) {
    __shared__ double sm[kThreads / 32];
    __shared__ double wg[8], rgn[8];
    const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = A.n;
    const CgVecs& v = A.v;
    double* gp[2] = {A.gpart0, A.gpart1};
    st.converged = 0;
    while (st.k < max_iters) {
        const int par = st.k & 1;
        const bool first = st.k == 0;
        const double gamma = sum_small(gp[par], cs);
        const double delta = sum_small(A.dpart, cs);
        if (first) st.gamma0 = gamma;
        if (!first && gamma <= rtol2 * st.gamma0) { st.converged = 1; break; }      // the same sums in every thread: uniform
        const double beta = first ? 0.0 : gamma / st.gprev;
        const double denom = first ? delta : delta - beta * gamma / st.aprev;
        const double alpha = gamma / denom;
        if (!(denom > 0.0) || !isfinite(alpha)) { st.breakdown = 1; break; }
        double g[1] = {0.0};
        for (int i = rank * kThreads + threadIdx.x; i < n; i += cs * kThreads) {
            D3 z1, z2, w1, w2, p1, p2, s1, s2, x1, x2, r1, r2;
            load6_l2(v.z, i, z1, z2); load6_l2(v.w, i, w1, w2); load6_l2(v.r, i, r1, r2);
            if (first) {
                p1 = z1; p2 = z2; s1 = w1; s2 = w2;
                x1 = alpha * p1; x2 = alpha * p2;
            } else {
                load6_l2(v.p, i, p1, p2); load6_l2(v.s, i, s1, s2); load6_l2(v.x, i, x1, x2);
                p1 = z1 + beta * p1; p2 = z2 + beta * p2;
                s1 = w1 + beta * s1; s2 = w2 + beta * s2;
                x1 = x1 + alpha * p1; x2 = x2 + alpha * p2;
            }
            r1 = r1 - alpha * s1; r2 = r2 - alpha * s2;
            double r[6] = {r1.x, r1.y, r1.z, r2.x, r2.y, r2.z}, zn[6], M[21];
            const double* Mp = blk21(A.Minv, i);
#pragma unroll
            for (int q = 0; q < 21; ++q) M[q] = __ldcg(Mp + q * 32);
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                double sacc = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) sacc += (a <= c ? M[pk<6>(a, c)] : M[pk<6>(c, a)]) * r[c];
                zn[a] = sacc;
            }
            store6(v.p, i, p1, p2); store6(v.s, i, s1, s2); store6(v.x, i, x1, x2); store6(v.r, i, r1, r2);
            store6(v.z, i, d3(zn[0], zn[1], zn[2]), d3(zn[3], zn[4], zn[5]));
#pragma unroll
            for (int q = 0; q < 6; ++q) g[0] += r[q] * zn[q];
        }
        if (rank == 0) {
            __syncthreads();
            double bsum = 0.0;                          // warp 0: sum over the CTAs of bpart[c][q], q = lane % 8
            if (warp == 0) {
                for (int c = lane >> 3; c < cs; c += 4) bsum += __ldcg(A.bpart + 8 * c + (lane & 7));
                bsum += __shfl_xor_sync(0xffffffffu, bsum, 8);
                bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
            }
            if (threadIdx.x < 8) {
                const int q = threadIdx.x;
                const double s = bsum;
                const double d = (q >= 6 ? __ldcg(&A.lin->C[q * 8 + q]) : 0.0) + lambda;
                const double zq = __ldcg(v.zg + q);
                wg[q] = s + d * zq;
                const double pg = first ? zq : zq + beta * __ldcg(v.pg + q);
                const double sg = first ? wg[q] : wg[q] + beta * __ldcg(v.sg + q);
                v.pg[q] = pg; v.sg[q] = sg;
                v.xg[q] = (first ? 0.0 : __ldcg(v.xg + q)) + alpha * pg;
                const double rr = __ldcg(v.rg + q) - alpha * sg;
                v.rg[q] = rr; rgn[q] = rr;
            }
            __syncthreads();
            if (threadIdx.x < 8) {
                const int q = threadIdx.x;
                double s = 0.0;
                for (int c = 0; c < 8; ++c) s += __ldcg(A.Ginv + q * 8 + c) * rgn[c];
                v.zg[q] = s;
                wg[q] = s * rgn[q];
            }
            __syncthreads();
            if (threadIdx.x == 0)
                for (int q = 0; q < 8; ++q) g[0] += wg[q];
        }
        block_reduce<1>(g, sm);
        if (threadIdx.x == 0) gp[par ^ 1][rank] = g[0];
        st.gprev = gamma; st.aprev = alpha;
        ++st.k;
        cluster.sync();
        cluster_spmv<kRO>(rank, cs, A, Rg, pr, W, lambda);
        cluster.sync();
    }
}

// fresh = 1: start a solve (preconditioner, r = b, z = M^-1 r, first operator application); fresh = 0: resume the solve
// whose state is in the vectors and in ctl (pause / resume of the early rejection).  Runs until converged
// (gamma <= rtol2 gamma0), breakdown, or max_iters updates in total.  gpart[2][cs], dpart[cs], bpart[cs][8].
__global__ void __launch_bounds__(kThreads, 1)
pcg_cluster_kernel(int n, int fresh, int max_iters, const double* __restrict__ P, const double* __restrict__ Je,
                   const double* __restrict__ U, const int* __restrict__ sliceptr, const int* __restrict__ ecol,
                   const Globals* __restrict__ Gp, const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W,
                   const double* __restrict__ b, const double* __restrict__ D, const LinGlobal* __restrict__ lin,
                   double* Minv, double* Ginv, int* err, CgVecs v, double* gpart0, double* gpart1, double* dpart, double* bpart,
                   CgControl* ctl) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double Rg[9];
    if (threadIdx.x < 9) Rg[threadIdx.x] = Gp->Rg[threadIdx.x];
    const double lambda = __ldcg(&ctl->lambda);
    const double rtol2 = __ldcg(&ctl->rtol2);
    __syncthreads();
    ClusterPcgArgs A{n, P, Je, U, sliceptr, ecol, b, D, lin, Minv, Ginv, err, v, gpart0, gpart1, dpart, bpart};
    ClusterPcgState st;
    if (fresh) {
        cluster_pcg_begin<true>(cluster, A, Rg, pr, W, lambda, st);
    } else {
        st.k = __ldcg(&ctl->iters);
        st.gamma0 = __ldcg(&ctl->gamma0);
        st.gprev = __ldcg(&ctl->sc[(st.k & 1) ^ 1].gamma_prev);
        st.aprev = __ldcg(&ctl->sc[(st.k & 1) ^ 1].alpha_prev);
        st.converged = 0; st.breakdown = 0;
    }
    cluster_pcg_run<true>(cluster, A, Rg, pr, W, lambda, rtol2, max_iters, st);
    // every CTA has read the control block before any CTA can get here (at least one cluster barrier lies between)
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        ctl->iters = st.k;
        ctl->gamma0 = st.gamma0;
        ctl->sc[(st.k & 1) ^ 1].gamma_prev = st.gprev; ctl->sc[(st.k & 1) ^ 1].alpha_prev = st.aprev;
        ctl->converged = st.converged;
        ctl->breakdown = st.breakdown;
    }
}

// The same solve by the WHOLE device: a cooperative launch of one CTA per SM, grid barriers instead of cluster barriers.
// For pairs between the cluster path and a few hundred thousand correspondences an iteration of the per-iteration kernels
// is two launches of ~20-30 us each whose duration is latency, not bandwidth; here it is two grid barriers.
__global__ void __launch_bounds__(kThreads, 1)
pcg_grid_kernel(int n, int fresh, int max_iters, const double* __restrict__ P, const double* __restrict__ Je,
                const double* __restrict__ U, const int* __restrict__ sliceptr, const int* __restrict__ ecol,
                const Globals* __restrict__ Gp, const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W,
                const double* __restrict__ b, const double* __restrict__ D, const LinGlobal* __restrict__ lin,
                double* Minv, double* Ginv, int* err, CgVecs v, double* gpart0, double* gpart1, double* dpart, double* bpart,
                CgControl* ctl) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double Rg[9];
    if (threadIdx.x < 9) Rg[threadIdx.x] = Gp->Rg[threadIdx.x];
    const double lambda = __ldcg(&ctl->lambda);
    const double rtol2 = __ldcg(&ctl->rtol2);
    __syncthreads();
    ClusterPcgArgs A{n, P, Je, U, sliceptr, ecol, b, D, lin, Minv, Ginv, err, v, gpart0, gpart1, dpart, bpart};
    ClusterPcgState st;
    if (fresh) {
        cluster_pcg_begin<true>(grid, A, Rg, pr, W, lambda, st);
    } else {
        st.k = __ldcg(&ctl->iters);
        st.gamma0 = __ldcg(&ctl->gamma0);
        st.gprev = __ldcg(&ctl->sc[(st.k & 1) ^ 1].gamma_prev);
        st.aprev = __ldcg(&ctl->sc[(st.k & 1) ^ 1].alpha_prev);
        st.converged = 0; st.breakdown = 0;
        grid.sync();                                   // every CTA has read the control block before rank 0 may rewrite it below
    }
    cluster_pcg_run<true>(grid, A, Rg, pr, W, lambda, rtol2, max_iters, st);
    if (grid.block_rank() == 0 && threadIdx.x == 0) {
        ctl->iters = st.k;
        ctl->gamma0 = st.gamma0;
        ctl->sc[(st.k & 1) ^ 1].gamma_prev = st.gprev; ctl->sc[(st.k & 1) ^ 1].alpha_prev = st.aprev;
        ctl->converged = st.converged;
        ctl->breakdown = st.breakdown;
    }
}

}  // namespace dsc
