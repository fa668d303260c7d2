// Mesh.h -- the neighbour graph of arapOptimization built on the host:
//   ComputeDelaunayTriangulation3D  Modules/Utils/Geometry.cc:317-368 (2-D Delaunay of world (x,y); Qhull "d Qbb Qt")
//   TriangleMesh::ComputeAdjacencyList / GetSurfaceArea (Open3D, absent from the reference tree)
//   ComputeEdgeWeightsCot           Modules/Utils/Geometry.cc:272-298
// Qhull is not available, so the triangulation is an incremental Bowyer-Watson (walk + cavity) with long double
// predicates; a Delaunay triangulation is unique for points in general position, so the adjacency equals Qhull's
// there (tests/test_host_shim.py compares with scipy's Qhull).  Replaces the reference's O(N^2) createVectorMap
// (:300-315): mesh vertex k IS position k.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace dsc_host {

struct Graph {
    std::vector<int32_t> rowptr, col;
    std::vector<double> w;
    double area = 0.0;
    long long n_triangles = 0;
};

class Delaunay2D {
public:
    // xy: n points (x0,y0,x1,y1,...).  Returns triangles as vertex index triples.
    static std::vector<std::array<int, 3>> triangulate(const double* xy, int n) {
        Delaunay2D d(xy, n);
        d.run();
        return d.result();
    }

private:
    struct Tri { int v[3]; int nb[3]; bool alive; };
    const double* xy_;
    int n_;
    std::vector<long double> px_, py_;
    std::vector<Tri> t_;

    Delaunay2D(const double* xy, int n) : xy_(xy), n_(n) {}

    long double orient(int a, int b, int c) const {
        return (px_[b] - px_[a]) * (py_[c] - py_[a]) - (py_[b] - py_[a]) * (px_[c] - px_[a]);
    }
    bool in_circle(const Tri& t, int p) const {
        long double ax = px_[t.v[0]] - px_[p], ay = py_[t.v[0]] - py_[p];
        long double bx = px_[t.v[1]] - px_[p], by = py_[t.v[1]] - py_[p];
        long double cx = px_[t.v[2]] - px_[p], cy = py_[t.v[2]] - py_[p];
        long double det = (ax * ax + ay * ay) * (bx * cy - cx * by) - (bx * bx + by * by) * (ax * cy - cx * ay) +
                          (cx * cx + cy * cy) * (ax * by - bx * ay);
        return det > 0.0L;
    }
    static uint32_t spread(uint32_t v) {
        v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f;
        v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v;
    }

    void run() {
        px_.resize(n_ + 3); py_.resize(n_ + 3);
        long double xmin = 1e300L, xmax = -1e300L, ymin = 1e300L, ymax = -1e300L;
        for (int i = 0; i < n_; ++i) {
            px_[i] = xy_[2 * i]; py_[i] = xy_[2 * i + 1];
            xmin = std::min(xmin, px_[i]); xmax = std::max(xmax, px_[i]); ymin = std::min(ymin, py_[i]); ymax = std::max(ymax, py_[i]);
        }
        if (n_ < 3) return;
        long double dx = xmax - xmin, dy = ymax - ymin, dm = std::max(dx, dy), cx = 0.5L * (xmin + xmax), cy = 0.5L * (ymin + ymax);
        if (!(dm > 0)) return;
        const long double K = 4096.0L;
        px_[n_] = cx - K * dm; py_[n_] = cy - K * dm;
        px_[n_ + 1] = cx + K * dm; py_[n_ + 1] = cy - K * dm;
        px_[n_ + 2] = cx; py_[n_ + 2] = cy + K * dm;
        t_.reserve(2 * (size_t)n_ + 16);
        t_.push_back(Tri{{n_, n_ + 1, n_ + 2}, {-1, -1, -1}, true});
        // insertion order: Morton curve (short walks)
        std::vector<std::pair<uint32_t, int>> order(n_);
        for (int i = 0; i < n_; ++i) {
            uint32_t qx = (uint32_t)(65535.0L * (px_[i] - xmin) / dm), qy = (uint32_t)(65535.0L * (py_[i] - ymin) / dm);
            order[i] = {spread(qx) | (spread(qy) << 1), i};
        }
        std::sort(order.begin(), order.end());
        int last = 0;
        std::vector<int> cavity, stack;
        std::vector<char> mark;
        struct BEdge { int a, b, outer; };
        std::vector<BEdge> border;
        for (int oi = 0; oi < n_; ++oi) {
            int p = order[oi].second;
            // walk to a triangle containing p
            int cur = last;
            while (!t_[cur].alive) cur = (int)t_.size() - 1;
            for (int guard = 0; guard < (int)t_.size() + 8; ++guard) {
                const Tri& T = t_[cur];
                int go = -1;
                for (int e = 0; e < 3; ++e) {
                    int a = T.v[(e + 1) % 3], b = T.v[(e + 2) % 3];
                    if (orient(a, b, p) < 0 && T.nb[e] >= 0) { go = T.nb[e]; break; }
                }
                if (go < 0) break;
                cur = go;
            }
            // duplicate of an existing vertex: skip (no mesh vertex => no neighbours, like a merged Qhull point)
            bool dup = false;
            for (int e = 0; e < 3; ++e) { int v = t_[cur].v[e]; if (px_[v] == px_[p] && py_[v] == py_[p]) dup = true; }
            if (dup) continue;
            // cavity = connected set of triangles whose circumcircle contains p
            cavity.clear(); stack.clear();
            if (mark.size() < t_.size()) mark.resize(t_.size() * 2 + 16, 0);
            stack.push_back(cur); mark[cur] = 1;
            while (!stack.empty()) {
                int c = stack.back(); stack.pop_back();
                cavity.push_back(c);
                for (int e = 0; e < 3; ++e) {
                    int nb = t_[c].nb[e];
                    if (nb >= 0 && !mark[nb] && in_circle(t_[nb], p)) { mark[nb] = 1; stack.push_back(nb); }
                }
            }
            border.clear();
            for (int c : cavity)
                for (int e = 0; e < 3; ++e) {
                    int nb = t_[c].nb[e];
                    if (nb < 0 || !mark[nb]) border.push_back(BEdge{t_[c].v[(e + 1) % 3], t_[c].v[(e + 2) % 3], nb});
                }
            for (int c : cavity) { t_[c].alive = false; mark[c] = 0; }
            // fan of new triangles (p, a, b); link across the border and around p
            int base = (int)t_.size();
            std::map<int, int> start_at, end_at;         // vertex -> new triangle whose edge starts / ends there
            for (size_t k = 0; k < border.size(); ++k) {
                Tri nt{{p, border[k].a, border[k].b}, {border[k].outer, -1, -1}, true};
                int id = base + (int)k;
                if (border[k].outer >= 0) {
                    Tri& o = t_[border[k].outer];
                    for (int e = 0; e < 3; ++e) {
                        int a = o.v[(e + 1) % 3], b = o.v[(e + 2) % 3];
                        if ((a == border[k].b && b == border[k].a)) o.nb[e] = id;
                    }
                }
                t_.push_back(nt);
                start_at[border[k].a] = id; end_at[border[k].b] = id;
            }
            for (size_t k = 0; k < border.size(); ++k) {
                int id = base + (int)k;
                // edge (b, p) is opposite vertex a (index 1); neighbour = triangle starting at b
                t_[id].nb[1] = start_at.count(border[k].b) ? start_at[border[k].b] : -1;
                // edge (p, a) is opposite vertex b (index 2); neighbour = triangle ending at a
                t_[id].nb[2] = end_at.count(border[k].a) ? end_at[border[k].a] : -1;
            }
            if (mark.size() < t_.size()) mark.resize(t_.size() * 2 + 16, 0);
            last = base;
        }
    }
    std::vector<std::array<int, 3>> result() const {
        std::vector<std::array<int, 3>> out;
        for (const Tri& t : t_)
            if (t.alive && t.v[0] < n_ && t.v[1] < n_ && t.v[2] < n_) out.push_back({t.v[0], t.v[1], t.v[2]});
        return out;
    }
};

// adjacency list (ascending), cot weights (mean over adjacent triangles, clamped to >= min_weight), 3-D area
inline Graph mesh_graph(const std::vector<std::array<double, 3>>& V, const std::vector<std::array<int, 3>>& tri, double min_weight = 0.0) {
    Graph g;
    int n = (int)V.size();
    g.n_triangles = (long long)tri.size();
    std::map<std::pair<int, int>, std::pair<double, int>> acc;
    auto sub = [](const std::array<double, 3>& a, const std::array<double, 3>& b) { return std::array<double, 3>{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
    auto dot = [](const std::array<double, 3>& a, const std::array<double, 3>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    auto crs = [](const std::array<double, 3>& a, const std::array<double, 3>& b) {
        return std::array<double, 3>{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    };
    for (const auto& t : tri) {
        auto c0 = crs(sub(V[t[1]], V[t[0]]), sub(V[t[2]], V[t[0]]));
        g.area += 0.5 * std::sqrt(dot(c0, c0));
        for (int e = 0; e < 3; ++e) {
            int a = t[e], b = t[(e + 1) % 3], o = t[(e + 2) % 3];
            auto va = sub(V[a], V[o]), vb = sub(V[b], V[o]);
            auto cr = crs(va, vb);
            double cot = dot(va, vb) / std::sqrt(dot(cr, cr));
            auto key = std::make_pair(std::min(a, b), std::max(a, b));
            auto& s = acc[key];
            s.first += cot; s.second += 1;
        }
    }
    std::vector<std::vector<std::pair<int, double>>> rows(n);
    for (auto& kv : acc) {
        double w = kv.second.second > 0 ? kv.second.first / kv.second.second : 0.0;
        if (w < min_weight) w = min_weight;
        rows[kv.first.first].push_back({kv.first.second, w});
        rows[kv.first.second].push_back({kv.first.first, w});
    }
    g.rowptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) { std::sort(rows[i].begin(), rows[i].end()); g.rowptr[i + 1] = g.rowptr[i] + (int)rows[i].size(); }
    g.col.reserve(g.rowptr[n]); g.w.reserve(g.rowptr[n]);
    for (int i = 0; i < n; ++i) for (auto& pr : rows[i]) { g.col.push_back(pr.first); g.w.push_back(pr.second); }
    return g;
}

}  // namespace dsc_host
