// dsc_graph.cuh -- device side of dsc_set_graph: internal numbering and the sliced-ELL form of the neighbour graph.
// The caller's CSR (Delaunay adjacency + cot weights of Modules/Utils/Geometry.cc:258-368, or a k-NN graph) is copied to
// the device as it is; everything else happens here:
//   1. Morton code of KF1's world (x, y) per correspondence, radix sort (cub) -> perm (internal -> caller index)
//   2. inside every group of kSortGroup consecutive rows: stable sort by degree, descending (equal-length ELL slices)
//   3. slice widths (max degree of 32 rows), exclusive scan -> sliceptr
//   4. one thread per row: renumber the neighbours, order them (own tile first, halo last, ascending inside each part),
//      write column indices and weights into the row's ELL slots, pad with (self, 0)
//   5. observations gathered into the internal order (permute_obs_kernel)
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "dsc_kernels.cuh"

namespace dsc {

DSC_HD unsigned part1by1_dev(unsigned v) {
    v &= 0x0000ffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// Graph checks of dsc_set_graph (the reference's mesh adjacency always passes them): column in range, no self loop, no
// duplicate edge, the reverse edge exists and carries the same weight.  err = the largest code found (the more
// fundamental defect wins): 4 out of range / self loop, 3 duplicate edge, 2 weights not symmetric, 1 not symmetric.
// Rows need not be sorted.
__global__ void validate_graph_kernel(int n, const int* __restrict__ rowptr0, const int* __restrict__ col0,
                                      const double* __restrict__ w0, int* __restrict__ err) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int code = 0;
        const int e0 = rowptr0[i], e1 = rowptr0[i + 1];
        for (int e = e0; e < e1; ++e) {
            const int j = col0[e];
            if (j < 0 || j >= n || j == i) { code = max(code, 4); continue; }
            for (int q = e0; q < e; ++q) if (col0[q] == j) code = max(code, 3);
            int found = -1;
            for (int q = rowptr0[j]; q < rowptr0[j + 1]; ++q) if (col0[q] == i) { found = q; break; }
            if (found < 0) code = max(code, 1);
            else if (w0[found] != w0[e]) code = max(code, 2);
        }
        if (code) atomicMax(err, code);
    }
}

// key = morton(x, y) << 32 | caller index  (the same key the host version sorted)
// bounding box of KF1's (x, y) over the finite points: two stages, no host round trip.  part[grid][4] = xmin, xmax,
// ymin, ymax per block; bbox_final_kernel folds them into box[4] = xmin, sx, ymin, sy (sx = 65535 / extent, 0 if flat).
__global__ void __launch_bounds__(kThreads)
bbox_kernel(int n, const float* __restrict__ X1, float* __restrict__ part) {
    __shared__ float sm[4][kThreads / 32];
    float v[4] = {1e30f, -1e30f, 1e30f, -1e30f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = X1[3 * (size_t)i], y = X1[3 * (size_t)i + 1];
        if (isfinite(x) && isfinite(y)) { v[0] = fminf(v[0], x); v[1] = fmaxf(v[1], x); v[2] = fminf(v[2], y); v[3] = fmaxf(v[3], y); }
    }
    for (int o = 16; o > 0; o >>= 1) {
        v[0] = fminf(v[0], __shfl_xor_sync(0xffffffffu, v[0], o)); v[1] = fmaxf(v[1], __shfl_xor_sync(0xffffffffu, v[1], o));
        v[2] = fminf(v[2], __shfl_xor_sync(0xffffffffu, v[2], o)); v[3] = fmaxf(v[3], __shfl_xor_sync(0xffffffffu, v[3], o));
    }
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 4; ++k) sm[k][threadIdx.x >> 5] = v[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) {
            v[0] = fminf(v[0], sm[0][w]); v[1] = fmaxf(v[1], sm[1][w]); v[2] = fminf(v[2], sm[2][w]); v[3] = fmaxf(v[3], sm[3][w]);
        }
        for (int k = 0; k < 4; ++k) part[4 * blockIdx.x + k] = v[k];
    }
}
__global__ void bbox_final_kernel(int nb, const float* __restrict__ part, float* __restrict__ box) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
    for (int b = 0; b < nb; ++b) {
        xmin = fminf(xmin, part[4 * b]); xmax = fmaxf(xmax, part[4 * b + 1]);
        ymin = fminf(ymin, part[4 * b + 2]); ymax = fmaxf(ymax, part[4 * b + 3]);
    }
    box[0] = xmin; box[1] = xmax > xmin ? __fdiv_rn(65535.0f, __fsub_rn(xmax, xmin)) : 0.f;
    box[2] = ymin; box[3] = ymax > ymin ? __fdiv_rn(65535.0f, __fsub_rn(ymax, ymin)) : 0.f;
}
__global__ void morton_key_kernel(int n, const float* __restrict__ X1 /* [n][3], caller order */, const float* __restrict__ box,
                                  unsigned long long* __restrict__ key) {
    const float xmin = box[0], sx = box[1], ymin = box[2], sy = box[3];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = X1[3 * (size_t)i], y = X1[3 * (size_t)i + 1];
        const unsigned qx = isfinite(x) ? (unsigned)fminf(65535.0f, fmaxf(0.0f, __fmul_rn(__fsub_rn(x, xmin), sx))) : 0u;
        const unsigned qy = isfinite(y) ? (unsigned)fminf(65535.0f, fmaxf(0.0f, __fmul_rn(__fsub_rn(y, ymin), sy))) : 0u;
        key[i] = ((unsigned long long)(part1by1_dev(qx) | (part1by1_dev(qy) << 1)) << 32) | (unsigned)i;
    }
}
__global__ void perm_from_key_kernel(int n, const unsigned long long* __restrict__ key, int* __restrict__ perm) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) perm[i] = (int)(key[i] & 0xffffffffull);
}

// stable sort of perm[g0 .. g0 + kSortGroup) by degree, descending: rank = #(larger degree) + #(equal degree before me)
__global__ void __launch_bounds__(kSortGroup)
degree_sort_kernel(int n, const int* __restrict__ rowptr0, int* __restrict__ perm) {
    __shared__ int sdeg[kSortGroup];
    const int g0 = blockIdx.x * kSortGroup;
    const int m = min(kSortGroup, n - g0);
    const int t = threadIdx.x;
    int p = 0, d = -1;
    if (t < m) { p = perm[g0 + t]; d = rowptr0[p + 1] - rowptr0[p]; }
    sdeg[t] = d;
    __syncthreads();
    if (t < m) {
        int rank = 0;
        for (int u = 0; u < m; ++u) {
            const int du = sdeg[u];
            rank += (du > d || (du == d && u < t)) ? 1 : 0;
        }
        perm[g0 + rank] = p;
    }
}
__global__ void inverse_perm_kernel(int n, const int* __restrict__ perm, int* __restrict__ inv) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) inv[perm ? perm[i] : i] = i;
}
// width[s] = longest row of slice s (rows in internal order); width[nslices] = 0 so that the scan yields the total
__global__ void slice_width_kernel(int n, int nslices, const int* __restrict__ perm, const int* __restrict__ rowptr0,
                                   int* __restrict__ width) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s <= nslices; s += warps) {
        const int i = s * 32 + lane;
        int d = 0;
        if (s < nslices && i < n) { const int p = perm ? perm[i] : i; d = rowptr0[p + 1] - rowptr0[p]; }
        for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
        if (lane == 0) width[s] = d;
    }
}

// Row i (internal) = caller row perm[i]: neighbours renumbered with inv, sorted by (outside own tile, index), written
// to the ELL slots (sliceptr[i / 32] + k) * 32 + i % 32; the rest of the slice width is padded with (i, 0).
__global__ void __launch_bounds__(kThreads)
ell_fill_kernel(int n, const int* __restrict__ perm, const int* __restrict__ inv, const int* __restrict__ rowptr0,
                const int* __restrict__ col0, const double* __restrict__ w0, const int* __restrict__ sliceptr,
                int* __restrict__ ecol, double* __restrict__ ewgt) {
    const int npad = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += gridDim.x * blockDim.x) {
        const int sl = i >> 5, lane = i & 31;
        const int b0 = sliceptr[sl], width = sliceptr[sl + 1] - b0;
        int deg = 0, e0 = 0;
        if (i < n) { const int p = perm ? perm[i] : i; e0 = rowptr0[p]; deg = rowptr0[p + 1] - e0; }
        const int t0 = (i / kSortGroup) * kSortGroup, t1 = t0 + kSortGroup;
        const unsigned kHalo = 0x80000000u;
        for (int k = 0; k < deg; ++k) {                 // insertion sort by (halo flag, new index), in place in the ELL
            const int v = inv ? inv[col0[e0 + k]] : col0[e0 + k];
            const double wv = w0[e0 + k];
            const unsigned key = (unsigned)v | ((v >= t0 && v < t1) ? 0u : kHalo);
            int q = k - 1;
            while (q >= 0) {
                const size_t at = ((size_t)b0 + q) * 32 + lane;
                const unsigned kq = (unsigned)ecol[at];
                if (kq <= key) break;
                ecol[at + 32] = (int)kq; ewgt[at + 32] = ewgt[at];
                --q;
            }
            const size_t at = ((size_t)b0 + q + 1) * 32 + lane;
            ecol[at] = (int)key; ewgt[at] = wv;
        }
        for (int k = 0; k < deg; ++k) {                 // strip the flag
            const size_t at = ((size_t)b0 + k) * 32 + lane;
            ecol[at] = (int)((unsigned)ecol[at] & ~kHalo);
        }
        for (int k = deg; k < width; ++k) {
            const size_t at = ((size_t)b0 + k) * 32 + lane;
            ecol[at] = i < n ? i : 0; ewgt[at] = 0.0;
        }
    }
}

// observations, depth measurements and inverse variances gathered into the internal order
// (the raw arrays are the caller's own, copied as they are: uv1 / uv2 [n][2] float, d1 / d2 [n] double, isg1 / isg2 [n]
// float or NULL = 1; interleaving them for the kernels happens here, on the device, not in a host loop)
__global__ void permute_obs_kernel(int n, const int* __restrict__ perm, const float2* __restrict__ ruv1, const float2* __restrict__ ruv2,
                                   const double* __restrict__ rd1, const double* __restrict__ rd2, const float* __restrict__ risg1,
                                   const float* __restrict__ risg2, float4* __restrict__ uv, double2* __restrict__ dm,
                                   float2* __restrict__ isg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = perm ? perm[i] : i;
        const float2 a = ruv1[s], b = ruv2[s];
        uv[i] = make_float4(a.x, a.y, b.x, b.y);
        dm[i] = make_double2(rd1[s], rd2[s]);
        isg[i] = make_float2(risg1 ? risg1[s] : 1.0f, risg2 ? risg2[s] : 1.0f);
    }
}

}  // namespace dsc
