// dsc_knn.cuh -- symmetrised k-nearest-neighbour graph of the correspondences in the plane (x, y) on the GPU
// (graph set-up; the synthetic 100k / 1M configurations use this graph instead of the reference's
// Delaunay adjacency, Modules/Utils/Geometry.cc:317-368).  Uniform grid + ring search:
//   1. bounding box, cell size for ~2 points per cell, cell id per point
//   2. counting sort of the points by cell (histogram with integer atomics, exclusive scan, scatter)
//   3. one thread per point: visit the cells ring by ring, keep the k nearest in a sorted register/local array,
//      stop when the next ring cannot contain a closer point; ties broken by the smaller index
//   4. symmetrise: i ~ j if j in kNN(i) or i in kNN(j); rows are sorted ascending, so the CSR is deterministic
// Distances are fp64 of the float32 coordinates, exactly what a CPU k-d tree on the same floats computes.
#pragma once
#include "dsc_math.cuh"

namespace dsc {

constexpr int kKnnMax = 32;

struct KnnGrid { double x0, y0, inv_cell; int nx, ny; };

__global__ void knn_cell_kernel(int n, const float* __restrict__ X /* [n][3] */, KnnGrid g, int* __restrict__ cell, int* __restrict__ count) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cx = min(g.nx - 1, max(0, (int)(((double)X[3 * (size_t)i] - g.x0) * g.inv_cell)));
        int cy = min(g.ny - 1, max(0, (int)(((double)X[3 * (size_t)i + 1] - g.y0) * g.inv_cell)));
        int c = cy * g.nx + cx;
        cell[i] = c;
        atomicAdd(count + c, 1);
    }
}

// exclusive scan of m ints, three phases (block sums -> scan of sums by one block -> add)
constexpr int kScanBlock = 1024;
__global__ void scan_block_kernel(int m, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums) {
    __shared__ int s[kScanBlock];
    int i = blockIdx.x * kScanBlock + threadIdx.x;
    int v = i < m ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < kScanBlock; d <<= 1) {
        int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < m) out[i] = s[threadIdx.x] - v;
    if (threadIdx.x == kScanBlock - 1) sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void scan_sums_kernel(int nb, int* __restrict__ sums) {       // one block, serial over chunks
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanBlock) {
        __shared__ int s[kScanBlock];
        int i = base + threadIdx.x;
        int v = i < nb ? sums[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < kScanBlock; d <<= 1) {
            int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry += s[threadIdx.x];
        __syncthreads();
    }
}
__global__ void scan_add_kernel(int m, int* __restrict__ out, const int* __restrict__ sums) {
    int i = blockIdx.x * kScanBlock + threadIdx.x;
    if (i < m) out[i] += sums[blockIdx.x];
}

__global__ void knn_scatter_kernel(int n, const int* __restrict__ cell, const int* __restrict__ start, int* __restrict__ cursor,
                                   int* __restrict__ order) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int c = cell[i];
        order[start[c] + atomicAdd(cursor + c, 1)] = i;
    }
}

// k nearest neighbours of every point (self excluded); nbr[i][0..k) sorted by (distance, index)
__global__ void __launch_bounds__(128)
knn_search_kernel(int n, int k, const float* __restrict__ X, KnnGrid g, const int* __restrict__ start, const int* __restrict__ order,
                  int ncells, int* __restrict__ nbr) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int i = order[t];                          // walk the points in cell order: neighbouring threads search the same cells
        const double xi = (double)X[3 * (size_t)i], yi = (double)X[3 * (size_t)i + 1];
        const int cx = min(g.nx - 1, max(0, (int)((xi - g.x0) * g.inv_cell)));
        const int cy = min(g.ny - 1, max(0, (int)((yi - g.y0) * g.inv_cell)));
        double bd[kKnnMax];
        int bi[kKnnMax];
        int have = 0;
        const double cellw = 1.0 / g.inv_cell;
        const int rmax = max(g.nx, g.ny);
        for (int r = 0; r <= rmax; ++r) {
            if (have == k) {
                // every point of ring r is at least (r - 1) cells + the distance to the own cell border away
                const double lo = fmin(fmin(xi - (g.x0 + cx * cellw), (g.x0 + (cx + 1) * cellw) - xi),
                                       fmin(yi - (g.y0 + cy * cellw), (g.y0 + (cy + 1) * cellw) - yi)) + (r - 1) * cellw;
                if (lo > 0.0 && lo * lo > bd[k - 1]) break;
            }
            for (int dy = -r; dy <= r; ++dy) {
                const int yy = cy + dy;
                if (yy < 0 || yy >= g.ny) continue;
                const int step = (dy == -r || dy == r) ? 1 : 2 * r;      // only the border of the ring
                for (int dx = -r; dx <= r; dx += (step > 0 ? step : 1)) {
                    const int xx = cx + dx;
                    if (xx < 0 || xx >= g.nx) continue;
                    const int c = yy * g.nx + xx;
                    const int e1 = c + 1 < ncells ? start[c + 1] : n;
                    for (int e = start[c]; e < e1; ++e) {
                        const int j = order[e];
                        if (j == i) continue;
                        const double ddx = (double)X[3 * (size_t)j] - xi, ddy = (double)X[3 * (size_t)j + 1] - yi;
                        const double d = ddx * ddx + ddy * ddy;
                        if (have == k && !(d < bd[k - 1] || (d == bd[k - 1] && j < bi[k - 1]))) continue;
                        int p = have < k ? have : k - 1;                // insertion into the sorted list
                        while (p > 0 && (bd[p - 1] > d || (bd[p - 1] == d && bi[p - 1] > j))) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                        bd[p] = d; bi[p] = j;
                        if (have < k) ++have;
                    }
                    if (r == 0) break;
                }
            }
        }
        for (int q = 0; q < k; ++q) nbr[(size_t)i * k + q] = q < have ? bi[q] : -1;
    }
}

DSC_D bool knn_contains(const int* __restrict__ nbr, int k, int i, int j) {
    for (int q = 0; q < k; ++q) if (nbr[(size_t)i * k + q] == j) return true;
    return false;
}
// degree of the symmetrised graph: k own neighbours + the points that list i without being listed by i
__global__ void knn_extra_kernel(int n, int k, const int* __restrict__ nbr, int* __restrict__ deg) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        for (int q = 0; q < k; ++q) {
            const int i = nbr[(size_t)j * k + q];
            if (i >= 0 && !knn_contains(nbr, k, i, j)) atomicAdd(deg + i, 1);
        }
}
__global__ void knn_owncount_kernel(int n, int k, const int* __restrict__ nbr, int* __restrict__ deg) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int c = 0;
        for (int q = 0; q < k; ++q) c += nbr[(size_t)i * k + q] >= 0 ? 1 : 0;
        deg[i] = c;
    }
}
__global__ void knn_fill_kernel(int n, int k, const int* __restrict__ nbr, const int* __restrict__ rowptr, int* __restrict__ cursor,
                                int* __restrict__ col) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        for (int q = 0; q < k; ++q) {
            const int i = nbr[(size_t)j * k + q];
            if (i < 0) continue;
            col[rowptr[j] + atomicAdd(cursor + j, 1)] = i;                                  // own neighbour
            if (!knn_contains(nbr, k, i, j)) col[rowptr[i] + atomicAdd(cursor + i, 1)] = j;   // reverse edge
        }
}
__global__ void knn_sortrows_kernel(int n, const int* __restrict__ rowptr, int* __restrict__ col) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int a = rowptr[i], b = rowptr[i + 1];
        for (int p = a + 1; p < b; ++p) {
            const int v = col[p];
            int q = p - 1;
            while (q >= a && col[q] > v) { col[q + 1] = col[q]; --q; }
            col[q + 1] = v;
        }
    }
}

}  // namespace dsc
