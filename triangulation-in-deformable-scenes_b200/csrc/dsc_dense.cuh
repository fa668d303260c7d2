// dsc_dense.cuh -- direct solve of the LM step for SMALL problems (the reference's own sizes: ~10^2 .. 10^3 points).
// The reference factorises (H + lambda I) with g2o's LinearSolverEigen (sparse Cholesky,
// Modules/Optimization/g2oBundleAdjustment.cc:619-628).  Below ~500 correspondences the PCG path is bound by launch
// latency (2 launches x ~10^3 iterations per trial), while the whole matrix is a few MB: here it is assembled densely
// on the device (m = 6 n + 8 unknowns, column-major, lower triangle) and factorised by a blocked right-looking
// Cholesky (32-wide panels: potf2 / trsm / syrk kernels), then two triangular solves.  Same unknown layout as the PCG:
// rows 6 i .. 6 i + 5 = correspondence i (X1 | X2), the last 8 rows = T_g (6), s1, s2.
#pragma once
#include "dsc_kernels.cuh"

namespace dsc {

constexpr int kDenseNB = 32;

DSC_D size_t dense_at(int m, int r, int c) { return (size_t)c * m + r; }      // column-major

// H (lower triangle) from the linearisation: diagonal blocks D, ARAP off-diagonal blocks from the per-edge Jacobian
// records (block (i, j) = 2 W g_i g_j^T: the directed twins of an edge contribute equally), the 8-wide border and C.
__global__ void __launch_bounds__(kThreads)
dense_assemble_kernel(int n, int m, const double* __restrict__ P, const double* __restrict__ D, const double* __restrict__ U,
                      const double* __restrict__ Je, const int* __restrict__ sliceptr, const int* __restrict__ ecol,
                      const Globals* __restrict__ Gp, const __grid_constant__ PairDev pr, const __grid_constant__ WeightsDev W,
                      const LinGlobal* __restrict__ lin, double* __restrict__ H) {
    __shared__ double Rg[9];
    if (threadIdx.x < 9) Rg[threadIdx.x] = Gp->Rg[threadIdx.x];
    __syncthreads();
    const int gbase = 6 * n;                                    // first global row
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* Dp = blk21(D, i);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) H[dense_at(m, 6 * i + r, 6 * i + c)] = Dp[pk<6>(c, r) * 32];
        const double4 xi = ldg256(reinterpret_cast<const double4*>(P) + i);
        const D3 X1i = d3(xi.x, xi.y, xi.z);
        double bt[6][6];                                        // border block: rows T_g (6), columns this vertex (6)
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) bt[a][c] = 0.0;
        const int sl = i >> 5, lane = i & 31;
        for (int bk = sliceptr[sl]; bk < sliceptr[sl + 1]; ++bk) {
            const int j = ecol[(size_t)bk * 32 + lane];
            if (j == i) continue;                               // padding slot
            const double* jb = Je + (size_t)bk * 288 + lane;
            const D3 u = d3(jb[0], jb[32], jb[64]), mm = d3(jb[96], jb[128], jb[160]), g = d3(jb[192], jb[224], jb[256]);
            const D3 v2 = 2.0 * mulT(Rg, g);
            const double gi[6] = {-mm.x - 2.0 * g.x, -mm.y - 2.0 * g.y, -mm.z - 2.0 * g.z, u.x + v2.x, u.y + v2.y, u.z + v2.z};
            const double gj[6] = {mm.x - 2.0 * g.x, mm.y - 2.0 * g.y, mm.z - 2.0 * g.z, -u.x + v2.x, -u.y + v2.y, -u.z + v2.z};
            const double4 xj = ldg256(reinterpret_cast<const double4*>(P) + j);
            const D3 S1 = d3(X1i.x + xj.x, X1i.y + xj.y, X1i.z + xj.z);
            const D3 cw = 2.0 * cross(S1, g);
            const double gt[6] = {cw.x, cw.y, cw.z, -4.0 * g.x, -4.0 * g.y, -4.0 * g.z};
            const double w2 = 2.0 * W.arap_info;
            if (j < i) {                                        // lower triangle: block (rows i, columns j)
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = 0; c < 6; ++c) H[dense_at(m, 6 * i + r, 6 * j + c)] = w2 * gi[r] * gj[c];
            }
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c) bt[a][c] += w2 * gt[a] * gi[c];
        }
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
            for (int c = 0; c < 6; ++c) H[dense_at(m, gbase + a, 6 * i + c)] = bt[a][c];
        // depth-scale rows: kd_c * R_c[2,:] on the X_c columns
        const double* Up = U + ((size_t)(i >> 5) * kURec) * 32 + lane;
#pragma unroll
        for (int cam = 0; cam < 2; ++cam) {
            const double* R = cam == 0 ? pr.R1 : pr.R2;
            const double kd = Up[(12 + cam) * 32];
#pragma unroll
            for (int c = 0; c < 3; ++c) H[dense_at(m, gbase + 6 + cam, 6 * i + 3 * cam + c)] = kd * R[6 + c];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 64) {
        const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
        if (c <= r) H[dense_at(m, gbase + r, gbase + c)] = lin->C[r * 8 + c];
    }
}

// A <- H + lambda I (lower triangle only; the strictly upper part is never read)
__global__ void dense_shift_kernel(int m, const double* __restrict__ H, double lambda, double* __restrict__ A) {
    const size_t total = (size_t)m * m;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(k / m), r = (int)(k % m);
        if (r >= c) A[k] = H[k] + (r == c ? lambda : 0.0);
    }
}

// ---- blocked right-looking Cholesky, lower, in place
// Panel step k0: every block factors the 32 x 32 diagonal block itself in shared memory (redundantly: cheaper than one
// more launch), then solves its rows of the panel, L21 = A21 L11^-T (one thread per row, the row kept in shared memory
// so that all loops stay rolled: straight-line unrolled code of this size is instruction-fetch bound); block 0 keeps
// L11 and raises fail[0] when a pivot is not positive.  Every block READS the unfactored A11 in this launch, so block 0
// must not overwrite it here unless it is alone: with more than one block L11 goes to the side buffer `L11` and the
// trailing update of the same step (dense_syrk_kernel, stream-ordered after this launch) copies it into A.
constexpr int kPanelThreads = 128;
__global__ void __launch_bounds__(kPanelThreads)
dense_panel_kernel(int m, int k0, int nb, double* __restrict__ A, double* __restrict__ L11, int* __restrict__ fail) {
    __shared__ double T[kDenseNB][kDenseNB + 1];
    __shared__ double inv[kDenseNB];
    __shared__ double xs[kDenseNB][kPanelThreads];
    __shared__ int bad;
    const int tr = threadIdx.x % kDenseNB, tg = threadIdx.x / kDenseNB;       // row, column group (4 groups)
    if (threadIdx.x == 0) bad = 0;
    for (int c = tg; c < kDenseNB; c += kPanelThreads / kDenseNB)
        T[tr][c] = (tr < nb && c < nb && tr >= c) ? A[dense_at(m, k0 + tr, k0 + c)] : 0.0;
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        // every thread derives 1 / sqrt(pivot) itself (no serial section); T[k][k] keeps the pivot until the loop ends
        const double d = T[k][k];
        const bool good = d > 0.0 && isfinite(d);
        const double ik = good ? 1.0 / sqrt(d) : 1.0;
        if (threadIdx.x == 0) { inv[k] = ik; if (!good) bad = 1; }
        if (tg == 0 && tr > k) T[tr][k] *= ik;                                   // column k of L11 below the diagonal
        __syncthreads();
        for (int c = k + 1 + tg; c < nb; c += kPanelThreads / kDenseNB)
            if (tr >= c) T[tr][c] -= T[tr][k] * T[c][k];
        __syncthreads();
    }
    if (threadIdx.x < nb) T[threadIdx.x][threadIdx.x] *= inv[threadIdx.x];       // pivot * 1/sqrt(pivot) = sqrt(pivot)
    __syncthreads();
    if (blockIdx.x == 0) {
        if (bad && threadIdx.x == 0) fail[0] = 1;
        const bool alone = gridDim.x == 1;
        for (int c = tg; c < nb; c += kPanelThreads / kDenseNB)
            if (tr < nb && tr >= c) {
                if (alone) A[dense_at(m, k0 + tr, k0 + c)] = T[tr][c];
                else L11[tr * kDenseNB + c] = T[tr][c];
            }
    }
    const int row = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    for (int c = 0; c < nb; ++c) xs[c][threadIdx.x] = A[dense_at(m, row, k0 + c)];
    for (int c = 0; c < nb; ++c) {
        double sacc = xs[c][threadIdx.x];
        for (int t = 0; t < c; ++t) sacc -= xs[t][threadIdx.x] * T[c][t];
        sacc *= inv[c];                                                          // 1 / L11[c][c]
        xs[c][threadIdx.x] = sacc;
        A[dense_at(m, row, k0 + c)] = sacc;
    }
}
// trailing update A22 -= L21 L21^T, lower part, 32 x 32 tile per block; block (0, 0) also stores the panel's L11 from
// the side buffer into A when the panel launch had more than one block (see dense_panel_kernel)
__global__ void __launch_bounds__(kDenseNB * 8)
dense_syrk_kernel(int m, int k0, int nb, double* __restrict__ A, const double* __restrict__ L11, int copy_l11) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (copy_l11 && bi == 0 && bj == 0)
        for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
            const int r = e / nb, c = e % nb;
            if (r >= c) A[dense_at(m, k0 + r, k0 + c)] = L11[r * kDenseNB + c];
        }
    if (bj > bi) return;
    __shared__ double Li[kDenseNB][kDenseNB + 1], Lj[kDenseNB][kDenseNB + 1];
    const int base = k0 + nb;
    const int tx = threadIdx.x % kDenseNB, ty = threadIdx.x / kDenseNB;      // ty 0..7
    for (int q = ty; q < kDenseNB; q += 8) {
        const int ri = base + bi * kDenseNB + tx, rj = base + bj * kDenseNB + tx;
        Li[tx][q] = (ri < m && q < nb) ? A[dense_at(m, ri, k0 + q)] : 0.0;
        Lj[tx][q] = (rj < m && q < nb) ? A[dense_at(m, rj, k0 + q)] : 0.0;
    }
    __syncthreads();
    const int r = base + bi * kDenseNB + tx;
    for (int q = ty; q < kDenseNB; q += 8) {
        const int c = base + bj * kDenseNB + q;
        if (r < m && c < m && r >= c) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < kDenseNB; ++t) s += Li[tx][t] * Lj[q][t];
            A[dense_at(m, r, c)] -= s;
        }
    }
}
// x = L^-T L^-1 b in one block of 32 warps, 32 unknowns at a time: the 32 x 32 triangle by warp 0 (shuffles), the
// rest of the column block as a matrix-vector product by all threads.  y lives in shared memory (m <= kDenseMaxM).
constexpr int kDenseMaxM = 6 * 1000 + 8;        // DSC_DENSE_MAX correspondences: 48 KB of shared memory
__global__ void __launch_bounds__(1024)
dense_solve_kernel(int m, const double* __restrict__ L, const double* __restrict__ b, double* __restrict__ x) {
    extern __shared__ double y[];
    __shared__ double part[kDenseNB];
    __shared__ double T[kDenseNB][kDenseNB + 1];                // the diagonal block of the current step (no global
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // loads inside the sequential 32-step solves)
    for (int k = threadIdx.x; k < m; k += blockDim.x) y[k] = b[k];
    for (int k0 = 0; k0 < m; k0 += kDenseNB) {                  // forward: L y = b
        const int nb = min(kDenseNB, m - k0);
        T[lane][warp] = (lane < nb && warp < nb && lane >= warp) ? L[dense_at(m, k0 + lane, k0 + warp)] : 0.0;
        if (lane == warp) T[lane][lane] = lane < nb ? 1.0 / T[lane][lane] : 1.0;     // reciprocal pivots: no division in the chain
        __syncthreads();
        if (warp == 0) {
            double val = lane < nb ? y[k0 + lane] : 0.0;
            for (int t = 0; t < nb; ++t) {
                const double yt = __shfl_sync(0xffffffffu, val, t) * T[t][t];
                if (lane == t) val = yt;
                if (lane > t && lane < nb) val -= T[lane][t] * yt;
            }
            if (lane < nb) y[k0 + lane] = val;
        }
        __syncthreads();
        for (int r = k0 + nb + threadIdx.x; r < m; r += blockDim.x) {
            double sacc = 0.0;
            for (int c = 0; c < nb; ++c) sacc += L[dense_at(m, r, k0 + c)] * y[k0 + c];
            y[r] -= sacc;
        }
        __syncthreads();
    }
    for (int k0 = ((m - 1) / kDenseNB) * kDenseNB; k0 >= 0; k0 -= kDenseNB) {      // backward: L^T x = y
        const int nb = min(kDenseNB, m - k0);
        T[lane][warp] = (lane < nb && warp < nb && lane >= warp) ? L[dense_at(m, k0 + lane, k0 + warp)] : 0.0;
        if (lane == warp) T[lane][lane] = lane < nb ? 1.0 / T[lane][lane] : 1.0;
        if (warp < nb) {                                        // warp c: sum over the rows below of L[r][k0 + c] x[r]
            double sacc = 0.0;
            for (int r = k0 + nb + lane; r < m; r += 32) sacc += L[dense_at(m, r, k0 + warp)] * y[r];
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) part[warp] = sacc;
        }
        __syncthreads();
        if (warp == 0) {
            double val = lane < nb ? y[k0 + lane] - part[lane] : 0.0;
            for (int t = nb - 1; t >= 0; --t) {
                const double xt = __shfl_sync(0xffffffffu, val, t) * T[t][t];
                if (lane == t) val = xt;
                if (lane < t) val -= T[t][lane] * xt;
            }
            if (lane < nb) y[k0 + lane] = val;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < m; k += blockDim.x) x[k] = y[k];
}

// the LM right-hand side in the dense order: [b (6 n) | bg (8)]
__global__ void dense_rhs_kernel(int n, const double* __restrict__ b, const LinGlobal* __restrict__ lin, double* __restrict__ rhs) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 6 * n + 8; k += gridDim.x * blockDim.x)
        rhs[k] = k < 6 * n ? b[k] : lin->bg[k - 6 * n];
}
// solution back into the PCG's vectors: x [n][6] and xg[8]
__global__ void dense_scatter_kernel(int n, const double* __restrict__ sol, double* __restrict__ x, double* __restrict__ xg) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < 6 * n + 8; k += gridDim.x * blockDim.x) {
        if (k < 6 * n) x[k] = sol[k]; else xg[k - 6 * n] = sol[k];
    }
}

}  // namespace dsc
