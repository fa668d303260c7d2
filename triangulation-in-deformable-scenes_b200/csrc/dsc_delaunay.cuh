// dsc_delaunay.cuh -- the reference's neighbour graph on the GPU: 2-D Delaunay adjacency of the KF1 map
// points' world (x, y), cotangent edge weights, mesh area, triangle count
//   ComputeDelaunayTriangulation3D   Modules/Utils/Geometry.cc:317-368 (Qhull "d Qbb Qt" on (x, y))
//   ComputeEdgeWeightsCot            Modules/Utils/Geometry.cc:272-298 (mean over the adjacent triangles of a.b / |a x b|, >= 0)
//   GetSurfaceArea / triangles_.size()  g2oBundleAdjustment.cc:657-662,942-946
// Not an incremental triangulation: every point builds its own VORONOI CELL -- the intersection of the half planes of the
// points around it -- walking the cells of a uniform grid ring by ring and clipping a small convex polygon held in local
// memory; the generators of the polygon's edges are the point's Delaunay neighbours, in counter-clockwise order, and two
// consecutive ones close a Delaunay triangle.  A cell is CERTIFIED once every point not yet seen is farther away than
// twice the cell's radius (no such point can cut it).  Cells that reach outside the point cloud (hull points: ~sqrt(n) of
// them) never certify from nearby points alone; they are finished in a second pass from (a) the certified points that
// list them -- adjacency is symmetric and certified stars are exact -- and (b) each other, which together contain all
// their true neighbours.  Unbounded cells are cells that keep a side of a start box 2^40 times larger than the point
// cloud; the thin triangles along the hull, which a finite super triangle loses (host/Mesh.h, K = 4096), are found like
// any other, as Qhull finds them.  Predicates are plain double on coordinates relative to the cell's own point.
#pragma once
#include "dsc_kernels.cuh"
#include "dsc_knn.cuh"

namespace dsc {

constexpr int kDlMaxV = 32;                 // polygon vertices / star size a cell may reach
constexpr int kDlRingCap = 10;              // rings of grid cells a first-pass cell may visit before it is left to pass two
constexpr int kDlMaxExtra = 96;             // certified neighbours handed to an uncertified point
constexpr int kDlWsSecond = 32768;          // second-pass cells the context's workspace has room for (more: allocated on the spot)

struct DlCell {
    double vx[kDlMaxV], vy[kDlMaxV];        // vertices, counter-clockwise, relative to the cell's point
    double la[kDlMaxV], lb[kDlMaxV], lc[kDlMaxV];   // line of the edge from vertex i to vertex i + 1:  la x + lb y = lc
    int eid[kDlMaxV];                       // its generator (< 0: a side of the start box, i.e. "unbounded")
    int m;
    double R2;                              // largest squared vertex distance
    bool overflow;
};

// The cell starts as a box so large that it stands for the whole plane (2^40 x the extent of the point cloud): a hull
// point's cell keeps box sides, which is how "unbounded" is recorded.  Vertices are never interpolated between old
// vertices -- they are the intersection of the two generators' bisector LINES, so a vertex near the point is exact to
// rounding however far the vertices it replaces were.
DSC_D void dl_init(DlCell& c, double B) {
    c.m = 4; c.overflow = false;
    c.vx[0] = -B; c.vy[0] = -B; c.vx[1] = B; c.vy[1] = -B; c.vx[2] = B; c.vy[2] = B; c.vx[3] = -B; c.vy[3] = B;
    c.la[0] = 0.0; c.lb[0] = -1.0; c.lc[0] = B;     // y = -B
    c.la[1] = 1.0; c.lb[1] = 0.0; c.lc[1] = B;      // x =  B
    c.la[2] = 0.0; c.lb[2] = 1.0; c.lc[2] = B;      // y =  B
    c.la[3] = -1.0; c.lb[3] = 0.0; c.lc[3] = B;     // x = -B
    for (int i = 0; i < 4; ++i) c.eid[i] = -1 - i;
    c.R2 = 2.0 * B * B;
}
DSC_D void dl_meet(double a1, double b1, double c1, double a2, double b2, double c2, double& x, double& y) {
    const double det = a1 * b2 - a2 * b1;
    x = (c1 * b2 - c2 * b1) / det;
    y = (a1 * c2 - a2 * c1) / det;
}
// intersect the cell with the half plane of points closer to the origin than to q = (qx, qy); id = generator index
DSC_D void dl_clip(DlCell& c, double qx, double qy, int id) {
    const double q2 = qx * qx + qy * qy;
    if (!(q2 > 0.0) || q2 >= 4.0 * c.R2) return;           // coincident point, or too far to reach the cell
    const double h = 0.5 * q2;
    double s[kDlMaxV];
    bool any = false;
    for (int i = 0; i < c.m; ++i) {
        if (c.eid[i] == id) return;                        // already a generator (pass two meets some points twice): its own
        s[i] = c.vx[i] * qx + c.vy[i] * qy - h;            // vertices lie ON its bisector, round-off must not cut there again
        any |= s[i] > 0.0;
    }
    if (!any) return;
    int a = -1, b = -1;                                    // a: inside -> outside at edge a;  b: outside -> inside at edge b
    for (int i = 0; i < c.m; ++i) {
        const int j = i + 1 == c.m ? 0 : i + 1;
        if (s[i] <= 0.0 && s[j] > 0.0) a = i;
        if (s[i] > 0.0 && s[j] <= 0.0) b = i;
    }
    if (a < 0 || b < 0) return;                            // (cannot happen for a cell that contains its point)
    double nx[kDlMaxV], ny[kDlMaxV], na[kDlMaxV], nb[kDlMaxV], nc[kDlMaxV];
    int ne[kDlMaxV];
    int k = 0;
    const int first_in = b + 1 == c.m ? 0 : b + 1;
    int kept = a - first_in; if (kept < 0) kept += c.m;    // inside vertices first_in .. a
    if (kept + 3 > kDlMaxV) { c.overflow = true; return; }
    for (int t = 0, i = first_in; t <= kept; ++t, i = (i + 1 == c.m ? 0 : i + 1)) {
        nx[k] = c.vx[i]; ny[k] = c.vy[i]; na[k] = c.la[i]; nb[k] = c.lb[i]; nc[k] = c.lc[i]; ne[k] = c.eid[i]; ++k;
    }
    // leave through edge a (the kept vertex a keeps that edge); the new vertex starts the new generator's edge ...
    dl_meet(c.la[a], c.lb[a], c.lc[a], qx, qy, h, nx[k], ny[k]);
    na[k] = qx; nb[k] = qy; nc[k] = h; ne[k] = id; ++k;
    // ... which ends where it meets edge b, whose generator continues from there
    dl_meet(qx, qy, h, c.la[b], c.lb[b], c.lc[b], nx[k], ny[k]);
    na[k] = c.la[b]; nb[k] = c.lb[b]; nc[k] = c.lc[b]; ne[k] = c.eid[b]; ++k;
    double r2 = 0.0;
    for (int i = 0; i < k; ++i) {
        c.vx[i] = nx[i]; c.vy[i] = ny[i]; c.la[i] = na[i]; c.lb[i] = nb[i]; c.lc[i] = nc[i]; c.eid[i] = ne[i];
        r2 = fmax(r2, nx[i] * nx[i] + ny[i] * ny[i]);
    }
    c.m = k; c.R2 = r2;
}

// Pass one (list == nullptr): every point, certified or not.  Pass two (list = the uncertified points): the same ring
// walk, then the certified points that list the point (extra), then every other uncertified point.
// star[p][0 .. deg[p]): generators in counter-clockwise order (< 0: a side of the start box = unbounded); flag[p]:
// 1 certified, 0 not, 2 overflow.
__global__ void __launch_bounds__(128)
delaunay_cells_kernel(int n, const float* __restrict__ X, KnnGrid g, const int* __restrict__ start, const int* __restrict__ order, int ncells,
                      double box, const unsigned char* __restrict__ dup, const int* __restrict__ list, int nlist, const int* __restrict__ extra,
                      const int* __restrict__ nextra, int* __restrict__ star, int* __restrict__ deg, int* __restrict__ flag) {
    const int total = list ? nlist : n;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int i = list ? list[t] : order[t];          // pass one walks the points in cell order: neighbouring threads search the same cells
        if (dup[i]) { deg[i] = 0; flag[i] = 1; continue; } // no vertex of the mesh
        const double xi = (double)X[3 * (size_t)i], yi = (double)X[3 * (size_t)i + 1];
        const int cx = min(g.nx - 1, max(0, (int)((xi - g.x0) * g.inv_cell)));
        const int cy = min(g.ny - 1, max(0, (int)((yi - g.y0) * g.inv_cell)));
        DlCell c;
        dl_init(c, box);
        const double cellw = 1.0 / g.inv_cell;
        const int rmax = min(kDlRingCap, max(g.nx, g.ny));
        bool certified = false;
        for (int r = 0; r <= max(g.nx, g.ny); ++r) {
            // every point of ring r is at least (r - 1) cells + the distance to the own cell's border away
            const double lo = fmin(fmin(xi - (g.x0 + cx * cellw), (g.x0 + (cx + 1) * cellw) - xi),
                                   fmin(yi - (g.y0 + cy * cellw), (g.y0 + (cy + 1) * cellw) - yi)) + (r - 1) * cellw;
            if (lo > 0.0 && lo * lo > 4.0 * c.R2) { certified = true; break; }
            if (r > rmax) break;
            for (int dy = -r; dy <= r; ++dy) {
                const int yy = cy + dy;
                if (yy < 0 || yy >= g.ny) continue;
                const int step = (dy == -r || dy == r) ? 1 : 2 * r;      // only the border of the ring
                for (int dx = -r; dx <= r; dx += (step > 0 ? step : 1)) {
                    const int xx = cx + dx;
                    if (xx < 0 || xx >= g.nx) continue;
                    const int cc = yy * g.nx + xx;
                    const int e1 = cc + 1 < ncells ? start[cc + 1] : n;
                    for (int e = start[cc]; e < e1; ++e) {
                        const int j = order[e];
                        if (j != i && !dup[j]) dl_clip(c, (double)X[3 * (size_t)j] - xi, (double)X[3 * (size_t)j + 1] - yi, j);
                    }
                    if (r == 0) break;
                }
            }
        }
        if (list) {                                        // pass two: whoever else can be a neighbour
            const int ne = min(nextra[i], kDlMaxExtra);
            for (int k = 0; k < ne; ++k) {
                const int j = extra[(size_t)t * kDlMaxExtra + k];
                dl_clip(c, (double)X[3 * (size_t)j] - xi, (double)X[3 * (size_t)j + 1] - yi, j);
            }
            for (int k = 0; k < nlist; ++k) {
                const int j = list[k];
                if (j != i && !dup[j]) dl_clip(c, (double)X[3 * (size_t)j] - xi, (double)X[3 * (size_t)j + 1] - yi, j);
            }
            certified = nextra[i] <= kDlMaxExtra;
        }
        // the star: one entry per polygon edge (a generator owns at most one edge: dl_clip never re-adds one)
        int* sp = star + (size_t)i * kDlMaxV;
        for (int k = 0; k < c.m; ++k) sp[k] = c.eid[k];
        deg[i] = c.m;
        flag[i] = c.overflow ? 2 : (certified ? 1 : 0);
    }
}

// Points with the same (x, y) as a point of smaller index are DUPLICATES: like Qhull and the host triangulator, the mesh
// has one vertex per location -- a duplicate gets no cell and is nobody's generator.  Coincident points share a grid cell.
__global__ void delaunay_duplicates_kernel(int n, const float* __restrict__ X, KnnGrid g, const int* __restrict__ start,
                                           const int* __restrict__ order, int ncells, unsigned char* __restrict__ dup) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float xf = X[3 * (size_t)i], yf = X[3 * (size_t)i + 1];
        const int cx = min(g.nx - 1, max(0, (int)(((double)xf - g.x0) * g.inv_cell)));
        const int cy = min(g.ny - 1, max(0, (int)(((double)yf - g.y0) * g.inv_cell)));
        const int cc = cy * g.nx + cx;
        const int e1 = cc + 1 < ncells ? start[cc + 1] : n;
        unsigned char d = 0;
        for (int e = start[cc]; e < e1; ++e) {
            const int j = order[e];
            if (j < i && X[3 * (size_t)j] == xf && X[3 * (size_t)j + 1] == yf) { d = 1; break; }
        }
        dup[i] = d;
    }
}

// first and second moments of (x, y): the grid is laid over mean +- 4 sigma (inside the bounding box), so a few far
// outliers -- badly triangulated points -- cannot blow the cells up; points outside fall into the border cells.
// part[grid][4] = sum x, sum y, sum x^2, sum y^2
__global__ void __launch_bounds__(kThreads)
delaunay_moments_kernel(int n, const float* __restrict__ X, double cx, double cy, double* __restrict__ part) {
    __shared__ double sm[4 * (kThreads / 32)];
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = (double)X[3 * (size_t)i] - cx, y = (double)X[3 * (size_t)i + 1] - cy;
        a[0] += x; a[1] += y; a[2] += x * x; a[3] += y * y;
    }
    block_reduce<4>(a, sm);
    if (threadIdx.x == 0) for (int k = 0; k < 4; ++k) part[4 * (size_t)blockIdx.x + k] = a[k];
}

// certified point p lists uncertified u  =>  p is a candidate for u's cell.  slot[u] = position of u in the list of
// uncertified points; extra[slot][...] filled with atomics (sorted afterwards: the cells are clipped in a fixed order).
__global__ void delaunay_extra_kernel(int n, const int* __restrict__ star, const int* __restrict__ deg, const int* __restrict__ flag,
                                      const int* __restrict__ slot, int* __restrict__ extra, int* __restrict__ nextra) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        if (flag[p] != 1) continue;
        for (int k = 0; k < deg[p]; ++k) {
            const int u = star[(size_t)p * kDlMaxV + k];
            if (u < 0 || u >= n || flag[u] == 1) continue;
            const int at = atomicAdd(nextra + u, 1);
            if (at < kDlMaxExtra) extra[(size_t)slot[u] * kDlMaxExtra + at] = p;
        }
    }
}
__global__ void delaunay_sort_extra_kernel(int nlist, const int* __restrict__ list, int* __restrict__ extra, const int* __restrict__ nextra) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nlist; t += gridDim.x * blockDim.x) {
        int* e = extra + (size_t)t * kDlMaxExtra;
        const int m = min(nextra[list[t]], kDlMaxExtra);
        for (int a = 1; a < m; ++a) { const int v = e[a]; int b = a - 1; while (b >= 0 && e[b] > v) { e[b + 1] = e[b]; --b; } e[b + 1] = v; }
    }
}
__global__ void delaunay_flag_kernel(int n, const int* __restrict__ flag, int* __restrict__ isu) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) isu[i] = flag[i] == 1 ? 0 : 1;
}
__global__ void delaunay_compact_kernel(int n, const int* __restrict__ isu, const int* __restrict__ pos, int* __restrict__ list, int* __restrict__ slot) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (isu[i]) { list[pos[i]] = i; slot[i] = pos[i]; }
}

// cot of the angle at o between a and b, ComputeEdgeWeightsCot (Geometry.cc:284-290) without contraction: both ends of an
// edge (and the host build, compiled without FMA) get the same bits
DSC_D double dl_cot(const float* __restrict__ X, int a, int b, int o) {
    const double ox = (double)X[3 * (size_t)o], oy = (double)X[3 * (size_t)o + 1], oz = (double)X[3 * (size_t)o + 2];
    const double ax = __dsub_rn((double)X[3 * (size_t)a], ox), ay = __dsub_rn((double)X[3 * (size_t)a + 1], oy), az = __dsub_rn((double)X[3 * (size_t)a + 2], oz);
    const double bx = __dsub_rn((double)X[3 * (size_t)b], ox), by = __dsub_rn((double)X[3 * (size_t)b + 1], oy), bz = __dsub_rn((double)X[3 * (size_t)b + 2], oz);
    const double cx = __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by)), cy = __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz)),
                 cz = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    const double d = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
    const double c2 = __dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz));
    return __ddiv_rn(d, __dsqrt_rn(c2));
}
DSC_D double dl_tri_area(const float* __restrict__ X, int a, int b, int c) {
    const double ox = (double)X[3 * (size_t)a], oy = (double)X[3 * (size_t)a + 1], oz = (double)X[3 * (size_t)a + 2];
    const double ux = (double)X[3 * (size_t)b] - ox, uy = (double)X[3 * (size_t)b + 1] - oy, uz = (double)X[3 * (size_t)b + 2] - oz;
    const double vx = (double)X[3 * (size_t)c] - ox, vy = (double)X[3 * (size_t)c + 1] - oy, vz = (double)X[3 * (size_t)c + 2] - oz;
    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

// Edges of a point from its star.  write = 0: rowdeg[p] = number of mesh edges at p; part[grid][2] = triangles owned by
// this block's points (a triangle belongs to its smallest vertex) and their 3-D area.  write = 1: col / w of the row,
// ascending.  An edge p-q is a mesh edge iff at least one of the two triangles beside it has three real vertices.
__global__ void __launch_bounds__(kThreads)
delaunay_edges_kernel(int n, const float* __restrict__ X, const int* __restrict__ star, const int* __restrict__ deg, double min_weight,
                      int write, const int* __restrict__ rowptr, int* __restrict__ rowdeg, int* __restrict__ col, double* __restrict__ w,
                      double* __restrict__ part) {
    __shared__ double sm[2 * (kThreads / 32)];
    double acc[2] = {0.0, 0.0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int m = deg[p];
        const int* sp = star + (size_t)p * kDlMaxV;
        int nq = 0;
        int qs[kDlMaxV];
        double ws[kDlMaxV];
        for (int k = 0; k < m; ++k) {
            const int q = sp[k];
            if (q < 0 || q >= n) continue;
            const int prev = sp[k == 0 ? m - 1 : k - 1], next = sp[k + 1 == m ? 0 : k + 1];
            const bool tp = prev >= 0 && prev < n && prev != q, tn = next >= 0 && next < n && next != q;
            if (!tp && !tn) continue;
            if (write) {
                double sum = 0.0;
                int cnt = 0;
                if (tp) { sum = __dadd_rn(sum, dl_cot(X, p, q, prev)); ++cnt; }
                if (tn && !(tp && next == prev)) { sum = __dadd_rn(sum, dl_cot(X, p, q, next)); ++cnt; }
                double wv = __ddiv_rn(sum, (double)cnt);
                if (wv < min_weight) wv = min_weight;
                int at = nq;                                   // insertion sort by neighbour index
                while (at > 0 && qs[at - 1] > q) { qs[at] = qs[at - 1]; ws[at] = ws[at - 1]; --at; }
                qs[at] = q; ws[at] = wv;
            } else if (tn && p < q && p < next) {              // triangle (p, q, next): counted once, by its smallest vertex
                acc[0] += 1.0; acc[1] += dl_tri_area(X, p, q, next);
            }
            ++nq;
        }
        if (write) { const int e0 = rowptr[p]; for (int k = 0; k < nq; ++k) { col[e0 + k] = qs[k]; w[e0 + k] = ws[k]; } }
        else rowdeg[p] = nq;
    }
    if (!write) {
        block_reduce<2>(acc, sm);
        if (threadIdx.x == 0) { part[2 * blockIdx.x] = acc[0]; part[2 * blockIdx.x + 1] = acc[1]; }
    }
}

}  // namespace dsc
