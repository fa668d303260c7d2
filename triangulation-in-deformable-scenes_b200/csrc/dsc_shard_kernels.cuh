// dsc_shard_kernels.cuh -- the PCG kernels of a point-sharded frame pair (dsc_shard.cuh has the protocol).  Same
// arithmetic as cg_init_kernel / cg_update_kernel / apply_update_kernel over the rank's own rows [row_begin, row_end),
// plus, fused into the same launches: the push of halo rows into the peers' buffers (NVLink stores) and the exchange of
// the rank totals of gamma (last block of the launch).  The operator is cg_spmv_kernel<double, true> (dsc_kernels.cuh).
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

// store a 6-vector / a point record of row i into the buffers of every rank that holds the row as a halo row
DSC_D void shard_push6(const ShardDev& S, double* const* bufs, unsigned mask, int i, D3 a, D3 b) {
    for (; mask; mask &= mask - 1) store6(bufs[__ffs((int)mask) - 1], i, a, b);
}

// the rank's total of count per-block partials part[grid][count] (fixed order) -> every rank's mailbox; called by all
// blocks after their partial has been written, the last block to arrive does the work
template <int kCount>
DSC_D void shard_finish(const ShardDev& S, int cls, const double* part, double* smem /* [kCount][kThreads / 32] */, bool pushed) {
    __shared__ double tot[kCount];
    if (!shard_last_block(S, cls, pushed)) return;
    double t[kCount];
#pragma unroll
    for (int e = 0; e < kCount; ++e) t[e] = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)           // (every thread loads: a handful of L2 round trips in all)
#pragma unroll
        for (int e = 0; e < kCount; ++e) t[e] += __ldcg(part + (size_t)i * kCount + e);
    block_reduce<kCount>(t, smem);
    if (threadIdx.x == 0) for (int e = 0; e < kCount; ++e) tot[e] = t[e];
    __syncthreads();
    shard_send(S, cls, tot, kCount);
}

__global__ void __launch_bounds__(kThreads)
shard_cg_init_kernel(const __grid_constant__ ShardDev S, int n, const double* __restrict__ b, const double* __restrict__ D, double lambda,
                     const LinGlobal* __restrict__ lin, double* __restrict__ Minv, double* __restrict__ Ginv, int* __restrict__ err,
                     CgVecs v, int zpar, double* __restrict__ gpart, CgControl* __restrict__ ctl) {
    __shared__ double sm[kThreads / 32];
    // every rank has finished the operator applications of the previous solve: the z buffers may be overwritten
    shard_wait(S, SF_S, shard_sent(S, SF_S));
    double* zout = S.zbuf[zpar][S.rank];
    double* const* zpeer = S.zbuf[zpar];
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    double g[1] = {0.0};
    bool pushed = false;
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        double r[6], z[6], M[21];
        D3 a, c;
        load6(b, i, a, c);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = c.x; r[4] = c.y; r[5] = c.z;
        precond_block(D, i, lambda, Minv, err, M);
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += (q <= k ? M[pk<6>(q, k)] : M[pk<6>(k, q)]) * r[k];
            z[q] = sacc;
        }
        const D3 z1 = d3(z[0], z[1], z[2]), z2 = d3(z[3], z[4], z[5]);
        store6(v.r, i, a, c);
        store6(zout, i, z1, z2);
        const unsigned em = S.exportmask[i];
        pushed |= em != 0;
        shard_push6(S, zpeer, em, i, z1, z2);
#pragma unroll
        for (int k = 0; k < 6; ++k) g[0] += r[k] * z[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {            // the 8 globals: every rank computes the same copy
        precond_global(lin, lambda, Ginv, err);
        for (int a = 0; a < 8; ++a) {
            double s = 0.0;
            for (int c = 0; c < 8; ++c) s += Ginv[a * 8 + c] * lin->bg[c];
            v.rg[a] = lin->bg[a]; v.zg[a] = s; v.xg[a] = 0.0; v.pg[a] = 0.0; v.sg[a] = 0.0;
            if (S.rank == 0) g[0] += lin->bg[a] * s;      // ... and rank 0 counts their share of gamma
        }
        ctl->iters = 0; ctl->converged = 0; ctl->breakdown = 0;
        ctl->sc[0].gamma_prev = 1.0; ctl->sc[0].alpha_prev = 1.0;
        ctl->sc[1].gamma_prev = 1.0; ctl->sc[1].alpha_prev = 1.0;
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) gpart[blockIdx.x] = g[0];
    shard_finish<1>(S, SF_Z, gpart, sm, pushed);
}

// One CG step over the rank's rows.  z is double-buffered by iteration parity: read from zbuf[zpar], written (own rows
// locally, halo rows into the peers) to zbuf[zpar ^ 1].
__global__ void __launch_bounds__(kThreads, 3)
shard_cg_update_kernel(const __grid_constant__ ShardDev S, int n, int par, int first, const double* __restrict__ Minv,
                       const double* __restrict__ Ginv, const LinGlobal* __restrict__ lin, CgVecs v, int zpar,
                       double* __restrict__ gpart_out, CgControl* __restrict__ ctl) {
    __shared__ double sm[kThreads / 32];
    if (ctl->converged || ctl->breakdown) return;            // set by an earlier launch, the same on every rank
    const double lambda = ctl->lambda, rtol2 = ctl->rtol2;
    const unsigned long long seq_s = shard_sent(S, SF_S), seq_z = shard_sent(S, SF_Z);
    shard_wait(S, SF_S, seq_s);                              // (the operator launch before this one has waited for seq_z)
    const double gamma = shard_total(S, SF_Z, seq_z, 0);
    const double delta = shard_total(S, SF_S, seq_s, 0);
    if (first && blockIdx.x == 0 && threadIdx.x == 0) ctl->gamma0 = gamma;
    const double gamma0 = first ? gamma : ctl->gamma0;
    if (!first && gamma <= rtol2 * gamma0) {                  // the same totals on every rank => the same decision
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->converged = 1;
        return;
    }
    const CgScalars prev = ctl->sc[par ^ 1];
    const double beta = first ? 0.0 : gamma / prev.gamma_prev;
    const double denom = first ? delta : delta - beta * gamma / prev.alpha_prev;
    const double alpha = gamma / denom;
    if (!(denom > 0.0) || !isfinite(alpha)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->breakdown = 1;
        return;
    }
    const double* zin = S.zbuf[zpar][S.rank];
    double* zout = S.zbuf[zpar ^ 1][S.rank];
    double* const* zpeer = S.zbuf[zpar ^ 1];
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    double g[1] = {0.0};
    bool pushed = false;
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        D3 z1, z2, w1, w2, p1, p2, s1, s2, x1, x2, r1v, r2v;
        load6(zin, i, z1, z2); load6(v.w, i, w1, w2); load6(v.r, i, r1v, r2v);
        if (first) {
            p1 = z1; p2 = z2; s1 = w1; s2 = w2;
            x1 = alpha * p1; x2 = alpha * p2;
        } else {
            load6(v.p, i, p1, p2); load6(v.s, i, s1, s2); load6(v.x, i, x1, x2);
            p1 = z1 + beta * p1; p2 = z2 + beta * p2;
            s1 = w1 + beta * s1; s2 = w2 + beta * s2;
            x1 = x1 + alpha * p1; x2 = x2 + alpha * p2;
        }
        r1v = r1v - alpha * s1; r2v = r2v - alpha * s2;
        double r[6] = {r1v.x, r1v.y, r1v.z, r2v.x, r2v.y, r2v.z}, zn[6];
        apply_minv(blk21(Minv, i), r, zn);
        const D3 zn1 = d3(zn[0], zn[1], zn[2]), zn2 = d3(zn[3], zn[4], zn[5]);
        store6(v.p, i, p1, p2); store6(v.s, i, s1, s2); store6(v.x, i, x1, x2); store6(v.r, i, r1v, r2v);
        store6(zout, i, zn1, zn2);
        const unsigned em = S.exportmask[i];
        pushed |= em != 0;
        shard_push6(S, zpeer, em, i, zn1, zn2);
#pragma unroll
        for (int k = 0; k < 6; ++k) g[0] += r[k] * zn[k];
    }
    if (blockIdx.x == 0) {                                    // the 8 global rows, identically on every rank
        __shared__ double wg[8], rgn[8];
        __syncthreads();
        if (threadIdx.x < 8) {
            const int k = threadIdx.x;
            const double s = shard_total(S, SF_S, seq_s, 1 + k);
            const double d = (k >= 6 ? lin->C[k * 8 + k] : 0.0) + lambda;
            wg[k] = s + d * v.zg[k];
            const double pg = v.zg[k] + beta * v.pg[k];
            const double sg = wg[k] + beta * v.sg[k];
            v.pg[k] = pg; v.sg[k] = sg;
            v.xg[k] += alpha * pg;
            const double rr = v.rg[k] - alpha * sg;
            v.rg[k] = rr; rgn[k] = rr;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            const int k = threadIdx.x;
            double s = 0.0;
            for (int c = 0; c < 8; ++c) s += Ginv[k * 8 + c] * rgn[c];
            v.zg[k] = s;
            wg[k] = s * rgn[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (S.rank == 0) for (int k = 0; k < 8; ++k) g[0] += wg[k];
            ctl->sc[par].gamma_prev = gamma; ctl->sc[par].alpha_prev = alpha;
            ctl->iters = ctl->iters + 1;
        }
    }
    block_reduce<1>(g, sm);
    if (threadIdx.x == 0) gpart_out[blockIdx.x] = g[0];
    shard_finish<1>(S, SF_Z, gpart_out, sm, pushed);
}

// trial state of the rank's rows (+ halo rows into the peers' trial buffers Pbuf[pidx]); part[grid]: partial of
// dx.(lambda dx + b); the exchange kernel that follows (SF_P) sums it over the ranks and publishes the halo rows
__global__ void __launch_bounds__(kThreads)
shard_apply_update_kernel(const __grid_constant__ ShardDev S, int n, const double* __restrict__ P, const double* __restrict__ x,
                          const double* __restrict__ xg, const double* __restrict__ b, const LinGlobal* __restrict__ lin, double lambda,
                          const Globals* __restrict__ Gcur, int pidx, Globals* __restrict__ Gtrial, double* __restrict__ part) {
    __shared__ double sm[kThreads / 32];
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    double4* own = reinterpret_cast<double4*>(S.Pbuf[pidx][S.rank]);
    double acc[1] = {0.0};
    for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += gridDim.x * blockDim.x) {
        D3 x1, x2, b1, b2;
        load6(x, i, x1, x2); load6(b, i, b1, b2);
        const P8 Pi = load_P(P, n, i);
        const double4 a = make_double4(Pi.a.x + x1.x, Pi.a.y + x1.y, Pi.a.z + x1.z, 0.0);
        const double4 c = make_double4(Pi.b.x + x2.x, Pi.b.y + x2.y, Pi.b.z + x2.z, 0.0);
        own[i] = a; own[(size_t)n + i] = c;
        for (unsigned m = S.exportmask[i]; m; m &= m - 1) {
            double4* dst = reinterpret_cast<double4*>(S.Pbuf[pidx][__ffs((int)m) - 1]);
            dst[i] = a; dst[(size_t)n + i] = c;
        }
        acc[0] += dot(x1, lambda * x1 + b1) + dot(x2, lambda * x2 + b2);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Globals g = *Gcur, o;
        const double share = apply_update_globals(g, xg, lin->bg, lambda, o);
        if (S.rank == 0) acc[0] += share;
        *Gtrial = o;
    }
    block_reduce<1>(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = acc[0];
    __threadfence_system();
}

}  // namespace dsc
