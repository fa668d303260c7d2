// dsc_batch.cuh -- the batched path (BASELINE.json configs[4]: thousands of independent ~10k-correspondence frame
// pairs): ONE launch refines a whole batch.  The grid is a set of thread-block clusters; a cluster takes frame pairs
// from a device-side queue and runs the COMPLETE Levenberg-Marquardt refinement of a pair -- every linearisation,
// every PCG solve, every trial evaluation, the lambda schedule and the accept / reject decisions of g2o's
// OptimizationAlgorithmLevenberg (Modules/Optimization/g2oBundleAdjustment.cc:619-628,959-962 + upstream g2o, as in
// dsc_optimize) -- without returning to the host: a 10k-correspondence pair costs no launch at all instead of the
// ~15 000 of the per-phase kernels.  Phases are separated by hardware cluster barriers; every decision is taken
// redundantly by all CTAs of the cluster from the same fixed-order sums of per-CTA partials, so control flow is
// uniform across the cluster and no flag has to be broadcast.
// The phases are the SAME device functions as the per-phase kernels (linearize_tiles, cost_tiles, apply_update_rows,
// cluster_pcg_*): same arithmetic; data another CTA of the launch writes is loaded L2-coherently (kRO = false).
// A pair's working set (~1.8 kB per correspondence: ~18 MB at 10k) lives in L2 while its cluster works on it: the more
// CTAs per cluster, the fewer pairs in flight and the larger the share of the 126 MB L2 each of them gets.
#pragma once
#include "dsc_kernels_ell.cuh"
#include "dsc_small.cuh"

namespace dsc {

// device-visible descriptor of one frame pair of a batch; every pointer is that pair's own buffer
struct BatchProblem {
    int n;
    double *P, *Ptrial;                         // current / trial state (the kernel leaves the result in P)
    const double* Q;
    const float4* uv; const double2* dm; const float2* isg;
    const int *sliceptr, *ecol; const double* ewgt;
    double *Je, *U, *D, *Minv, *b;
    CgVecs v; double* Ginv;
    Globals* G;                                 // in: initial globals, out: refined globals
    LinGlobal* lin; int* err;
    double *part;                               // [cs][kLinPart] linearisation partials, then [cs][4] trial / cost partials
    double *gpart0, *gpart1, *dpart, *bpart;
    PairDev pair; WeightsDev W;
};
struct BatchParams {
    int n_iters, max_pcg;
    double rtol;
    int early_levels;
    double early_rtol[4], early_margin[4];
};
struct BatchIterRec { double chi2_before, chi2_after, lambda; int trials, accepted, pcg_iters; };   // == dsc_iter_record
struct BatchResult { int iterations, total_trials, total_pcg_iters, terminated, early_rejects, pcg_unconverged, status, pad; double final_chi2; };

constexpr int kBatchEarlyWorthIters = 16;       // as dsc_optimize: pauses are skipped while full solves are this short

// LinGlobal from the per-CTA partials of linearize_tiles (what finalize_linearize_kernel does for a grid); every CTA
// computes the same copy in its own shared memory
DSC_D void batch_reduce_lin(const double* part, int cs, LinGlobal& out, double* tmp /* [kLinPart] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    __syncthreads();
    for (int k = warp; k < kLinPart; k += wpb) {
        double v = lane < cs ? __ldcg(part + (size_t)lane * kLinPart + k) : 0.0;
        if (k == 3) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); }
        else { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); }
        if (lane == 0) tmp[k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 64; ++k) out.C[k] = 0.0;
        for (int k = 0; k < 3; ++k) out.chi2[k] = tmp[k];
        for (int k = 0; k < 8; ++k) out.bg[k] = tmp[4 + k];
        int idx = 12;
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c, ++idx) { out.C[r * 8 + c] = tmp[idx]; out.C[c * 8 + r] = tmp[idx]; }
        out.C[6 * 8 + 6] = tmp[33]; out.C[7 * 8 + 7] = tmp[34];
        double m = tmp[3];
        for (int r = 0; r < 8; ++r) m = fmax(m, fabs(out.C[r * 8 + r]));
        out.maxdiag = m;                                  // computeLambdaInit: max over ALL vertices' diagonals
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1)
lm_batch_kernel(const BatchProblem* __restrict__ probs, int nprob, const __grid_constant__ BatchParams prm, int* queue,
                BatchIterRec* __restrict__ recs, BatchResult* __restrict__ res) {
    extern __shared__ double4 sw[];                        // the tile window of linearise / cost (kWinBytes)
    __shared__ PairDev pr;
    __shared__ WeightsDev W;
    __shared__ Globals G, Gt;
    __shared__ LinGlobal lin;
    __shared__ double tmp[kLinPart];
    __shared__ double sm[3 * (kThreads / 32)];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), cs = (int)cluster.num_blocks();
    const int cluster_id = blockIdx.x / cs;
    for (;;) {
        // ---- next frame pair of the queue (one atomic per pair, published to the cluster through its slot)
        if (rank == 0 && threadIdx.x == 0) {
            const int nxt = atomicAdd(queue, 1);
            *reinterpret_cast<volatile int*>(queue + 1 + cluster_id) = nxt;
        }
        cluster.sync();
        const int p = __ldcg(queue + 1 + cluster_id);
        if (p >= nprob) break;
        const BatchProblem& B = probs[p];
        const int n = B.n;
        {   // the pair, the weights and the globals into shared memory
            const int* src = reinterpret_cast<const int*>(&B.pair);
            int* dst = reinterpret_cast<int*>(&pr);
            for (int k = threadIdx.x; k < (int)(sizeof(PairDev) / sizeof(int)); k += blockDim.x) dst[k] = __ldg(src + k);
            if (threadIdx.x == 0) { W = B.W; G = *B.G; }
        }
        __syncthreads();
        ClusterPcgArgs A{n, B.P, B.Je, B.U, B.sliceptr, B.ecol, B.b, B.D, B.lin, B.Minv, B.Ginv, B.err, B.v, B.gpart0, B.gpart1, B.dpart, B.bpart};
        double* Pc = B.P;
        double* Pt = B.Ptrial;
        double* tpart = B.part + (size_t)cs * kLinPart;    // [cs][4]: trial scale, cost x 3
        double lambda = 0.0, ni = 2.0, current = 0.0;
        bool expect_long = true;
        BatchResult r;
        r.iterations = 0; r.total_trials = 0; r.total_pcg_iters = 0; r.terminated = 0; r.early_rejects = 0; r.pcg_unconverged = 0;
        r.status = 0; r.pad = 0; r.final_chi2 = 0.0;
        for (int it = 0; it < prm.n_iters && n > 0; ++it) {
            // ---- linearise (b, D, U, Je, partial chi2 / max diag / global block) and reduce
            linearize_tiles<false>(rank, cs, n, Pc, B.Q, B.uv, B.dm, B.isg, B.sliceptr, B.ecol, B.ewgt, G, pr, W, B.b, B.D, B.U, B.Je,
                                   B.part + (size_t)rank * kLinPart, sw);
            cluster.sync();
            batch_reduce_lin(B.part, cs, lin, tmp);
            if (rank == 0) {                               // the solver's rank-0 code reads the global block from memory
                double* dst = reinterpret_cast<double*>(B.lin);
                const double* src = reinterpret_cast<const double*>(&lin);
                for (int k = threadIdx.x; k < (int)(sizeof(LinGlobal) / sizeof(double)); k += blockDim.x) dst[k] = src[k];
                __threadfence();
                __syncthreads();
            }
            current = lin.chi2[0] + lin.chi2[1] + lin.chi2[2];
            if (!isfinite(current)) { r.status = -5; break; }                 // DSC_ERR_NONFINITE (uniform: same sums everywhere)
            if (it == 0) { lambda = 1e-5 * lin.maxdiag; ni = 2.0; }             // computeLambdaInit, tau = 1e-5
            BatchIterRec rec;
            rec.chi2_before = current; rec.lambda = lambda; rec.pcg_iters = 0;
            double rho = 0.0;
            int q = 0;
            bool accepted = false;
            do {
                ClusterPcgState st;
                A.P = Pc;
                cluster_pcg_begin<false>(cluster, A, G.Rg, pr, W, lambda, st);
                double temp = 1.7976931348623157e308, scale = 1e-3;
                bool rejected_early = false, solved = true;
                for (int level = 0; level <= prm.early_levels; ++level) {
                    const bool last = level == prm.early_levels;
                    const double tol = last ? prm.rtol : prm.early_rtol[level];
                    if (!last && (!(tol > prm.rtol) || !expect_long)) continue;
                    cluster_pcg_run<false>(cluster, A, G.Rg, pr, W, lambda, tol * tol, prm.max_pcg, st);
                    if (st.breakdown || !st.converged) { solved = false; if (!st.breakdown) r.pcg_unconverged++; break; }
                    // ---- trial state x (+) dx, its robust chi2 and the rho denominator dx.(lambda dx + b) + 1e-3
                    double acc[1];
                    acc[0] = apply_update_rows<false>(n, rank * kThreads + threadIdx.x, cs * kThreads, Pc, B.v.x, B.b, lambda, Pt);
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        double xg[8];
                        for (int k = 0; k < 8; ++k) xg[k] = __ldcg(B.v.xg + k);
                        const double share = apply_update_globals(G, xg, lin.bg, lambda, Gt);
                        if (rank == 0) acc[0] += share;
                    }
                    block_reduce<1>(acc, sm);
                    if (threadIdx.x == 0) tpart[4 * rank] = acc[0];
                    cluster.sync();
                    double c3[3] = {0.0, 0.0, 0.0};
                    cost_tiles<false>(rank, cs, n, Pt, B.Q, B.uv, B.dm, B.isg, B.sliceptr, B.ecol, B.ewgt, Gt, pr, W, sw, c3);
                    block_reduce<3>(c3, sm);
                    if (threadIdx.x == 0) { tpart[4 * rank + 1] = c3[0]; tpart[4 * rank + 2] = c3[1]; tpart[4 * rank + 3] = c3[2]; }
                    cluster.sync();
                    {
                        const int lane = threadIdx.x & 31;
                        double s[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double v = lane < cs ? __ldcg(tpart + 4 * lane + k) : 0.0;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            s[k] = v;
                        }
                        scale = s[0] + 1e-3;
                        temp = s[1] + s[2] + s[3];
                    }
                    if (!last && isfinite(temp) && (current - temp) / scale < -prm.early_margin[level]) {
                        rejected_early = true;
                        r.early_rejects++;
                        break;
                    }
                }
                if (!solved) { temp = 1.7976931348623157e308; scale = 1e-3; }
                rec.pcg_iters += st.k; r.total_pcg_iters += st.k;
                if (!rejected_early && solved) expect_long = st.k > kBatchEarlyWorthIters;
                rho = (current - temp) / scale;
                if (rho > 0 && isfinite(temp)) {
                    double alpha = 1.0 - pow(2.0 * rho - 1.0, 3);
                    alpha = fmin(alpha, 2.0 / 3.0);
                    lambda *= fmax(1.0 / 3.0, alpha);
                    ni = 2.0;
                    current = temp;
                    double* t = Pc; Pc = Pt; Pt = t;
                    __syncthreads();
                    if (threadIdx.x == 0) G = Gt;
                    __syncthreads();
                    accepted = true;
                } else {
                    lambda *= ni;
                    ni *= 2.0;
                }
                ++q;
            } while (rho < 0 && q < 10);
            rec.trials = q; rec.accepted = accepted ? 1 : 0; rec.chi2_after = current;
            r.total_trials += q; r.iterations = it + 1;
            if (rank == 0 && threadIdx.x == 0) recs[(size_t)p * prm.n_iters + it] = rec;
            if (q == 10 || rho == 0) { r.terminated = 1; break; }
        }
        // ---- leave the result where the host expects it: state in B.P, globals in B.G
        cluster.sync();                                    // every CTA is done reading Pc / Pt
        if (Pc != B.P && n > 0) {
            const double4* s4 = reinterpret_cast<const double4*>(Pc);
            double4* d4 = reinterpret_cast<double4*>(B.P);
            for (int i = rank * kThreads + threadIdx.x; i < 2 * n; i += cs * kThreads) d4[i] = ld256<false>(s4 + i);
        }
        if (rank == 0 && threadIdx.x == 0) {
            *B.G = G;
            r.final_chi2 = current;
            res[p] = r;
        }
        cluster.sync();                                    // the queue slot may be rewritten now
    }
}

}  // namespace dsc
