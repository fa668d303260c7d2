// Optimization.cc -- host shim: gathers the reference's Map into structure-of-arrays buffers, calls the CUDA
// library through the C ABI (include/dsc.h) and scatters the result back.  No numerical work of the hot path
// happens here; there is no CPU fallback (a missing GPU throws).
#include "Optimization.h"

#include <exception>
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <set>
#include <stdexcept>
#include <unordered_map>

#include "../../include/dsc.h"
#include "Mesh.h"
#ifdef DSC_IN_REFERENCE_TREE
#include "compat.h"          // only dsc_host::pose34 (the value types come from the real Eigen / Sophus / OpenCV)
#endif

namespace {

dsc_ctx* g_ctx = nullptr;
std::vector<dsc_host::LmRecord> g_trace;
dsc_host::PhaseTimes g_times;
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
dsc_pcg_params g_pcg{1e-10, 6000, 64};

dsc_ctx* ctx() {
    if (!g_ctx) {
        const char* d = std::getenv("DSC_DEVICE");
        int st = dsc_create(d ? std::atoi(d) : 0, &g_ctx);
        if (st != DSC_OK) throw std::runtime_error(std::string("dsc_create failed: ") + dsc_status_string(st));
    }
    return g_ctx;
}
void ck(int st, const char* what) {
    if (st != DSC_OK) throw std::runtime_error(std::string(what) + ": " + dsc_last_error(g_ctx));
}
int method_id(const std::string& m) {        // Geometry.cc:220-228: anything else is NRSLAM
    if (m == "Classic") return DSC_TRI_CLASSIC;
    if (m == "ORBSLAM") return DSC_TRI_ORBSLAM;
    if (m == "DepthMeasurement") return DSC_TRI_DEPTH;
    return DSC_TRI_NRSLAM;
}
int location_id(const std::string& l) {
    if (l == "TwoPoints") return DSC_LOC_TWOPOINTS;
    if (l == "FarPoints") return DSC_LOC_FARPOINTS;
    return DSC_LOC_INRAYS;
}
// The reference's CameraModel has no model tag: the dynamic type decides (KannalaBrandt8 carries 8 parameters,
// PinHole 4: Calibration/KannalaBrandt8.h:29-36, PinHole.h:36-46), the parameter count is the fallback.
int model_id(CameraModel* c) {
    if (dynamic_cast<KannalaBrandt8*>(c)) return DSC_CAM_KB8;
    if (dynamic_cast<PinHole*>(c)) return DSC_CAM_PINHOLE;
    return c->getNumberOfParameters() >= 8 ? DSC_CAM_KB8 : DSC_CAM_PINHOLE;
}
dsc_pair make_pair(KeyFrame& k1, KeyFrame& k2) {
    dsc_pair p{};
    auto c1 = k1.getCalibration(), c2 = k2.getCalibration();
    p.cam1.model = model_id(c1.get()); p.cam2.model = model_id(c2.get());
    for (int i = 0; i < 8; ++i) {
        p.cam1.params[i] = i < c1->getNumberOfParameters() ? c1->getParameter(i) : 0.f;
        p.cam2.params[i] = i < c2->getNumberOfParameters() ? c2->getParameter(i) : 0.f;
    }
    dsc_host::pose34(k1.getPose(), p.T1w);
    dsc_host::pose34(k2.getPose(), p.T2w);
    return p;
}
// Sophus::SE3f -> (qx,qy,qz,qw,tx,ty,tz) in double, Eigen's matrix->quaternion rule
void se3_to_7(const Sophus::SE3f& T, double* o) {
    auto Rm = T.rotationMatrix();
    double R[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = (double)Rm(r, c);
    double q[4], t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0); q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t; q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t; q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
    for (int k = 0; k < 4; ++k) o[k] = q[k];
    auto tr = T.translation();
    o[4] = tr[0]; o[5] = tr[1]; o[6] = tr[2];
}
Sophus::SE3f se3_from_7(const double* v) {
    double x = v[0], y = v[1], z = v[2], w = v[3];
    Eigen::Matrix3f R;
    R(0, 0) = (float)(1 - 2 * (y * y + z * z)); R(0, 1) = (float)(2 * (x * y - z * w)); R(0, 2) = (float)(2 * (x * z + y * w));
    R(1, 0) = (float)(2 * (x * y + z * w)); R(1, 1) = (float)(1 - 2 * (x * x + z * z)); R(1, 2) = (float)(2 * (y * z - x * w));
    R(2, 0) = (float)(2 * (x * z - y * w)); R(2, 1) = (float)(2 * (y * z + x * w)); R(2, 2) = (float)(1 - 2 * (x * x + y * y));
    return Sophus::SE3f(R, Eigen::Vector3f((float)v[4], (float)v[5], (float)v[6]));
}

// One key-frame pair gathered from the Map (g2oBundleAdjustment.cc:640-957 without the g2o objects)
struct PairProblem {
    KeyFrame_ kf1, kf2;
    ID kf1Id = 0, kf2Id = 0;
    std::vector<MapPoint_> mp1, mp2;
    std::vector<float> X1, X2, uv1, uv2, isg1, isg2;
    std::vector<double> d1, d2;
    dsc_host::Graph graph;
    long long mesh_edges = -1;           // >= 0: the mesh was built on the device (graph.rowptr / col / w are empty)
    double Tg[7];
};

// The reference reads depths through KeyFrame::getDepthMeasure(x, y, false), which throws when the key frame has no
// depth image (KeyFrame.cc:181-184) -- the case of the simulation, whose key frames only carry per-key-point depths
// (SURVEY.md 3.1).  Probed once per key frame with the reference's own call.
bool has_depth_image(KeyFrame& kf) {
    if (kf.getKeyPoints().empty()) return false;
    try {
        cv::Point2f p = kf.getKeyPoint(0).pt;
        (void)kf.getDepthMeasure(p.x, p.y, false);
        return true;
    } catch (const std::out_of_range&) {
        return true;                       // an image exists, the probe pixel was outside it
    } catch (const std::exception&) {
        return false;
    }
}

// What the device holds after the last refinement: lets calculatePixelsStandDev read the statistic without
// gathering the Map and rebuilding the mesh again (it is valid while the same MapPoints still sit at the written-back
// positions).
struct Resident {
    bool valid = false;
    Map* map = nullptr;
    ID kf1 = 0, kf2 = 0;
    std::vector<MapPoint_> mp1, mp2;
    std::vector<float> X1, X2;
} g_res;

bool resident_pair(Map* pMap, KeyFrame_ pKF1, ID kf1ID, KeyFrame_ pKF2, ID kf2ID) {
    if (!g_res.valid || g_res.map != pMap || g_res.kf1 != kf1ID || g_res.kf2 != kf2ID) return false;
    auto& v1 = pKF1->getMapPoints();
    auto& v2 = pKF2->getMapPoints();
    size_t slots = std::min(v1.size(), v2.size()), both = 0;
    for (size_t k = 0; k < slots; ++k) both += (v1[k] && v2[k]) ? 1 : 0;
    if (both < g_res.mp1.size()) return false;
    for (size_t i = 0; i < g_res.mp1.size(); ++i) {
        auto a = g_res.mp1[i]->getWorldPosition(), b = g_res.mp2[i]->getWorldPosition();
        for (int k = 0; k < 3; ++k)
            if (a[k] != g_res.X1[3 * i + k] || b[k] != g_res.X2[3 * i + k]) return false;
    }
    return true;
}

bool host_mesh_forced() { return std::getenv("DSC_HOST_MESH") != nullptr; }
bool host_mesh(PairProblem& pp);

bool gather_pair(Map* pMap, KeyFrame_ pKF1, ID kf1ID, KeyFrame_ pKF2, ID kf2ID, PairProblem& pp, bool with_mesh = true) {
    pp.kf1 = pKF1; pp.kf2 = pKF2; pp.kf1Id = kf1ID; pp.kf2Id = kf2ID;
    const bool img1 = has_depth_image(*pKF1), img2 = has_depth_image(*pKF2);
    auto& v1 = pKF1->getMapPoints();
    auto& v2 = pKF2->getMapPoints();
    size_t slots = std::min(v1.size(), v2.size());
    // Pass 1 (threads): which slots take part and where their observations are -- read-only look-ups in the Map.
    // Pass 2 (threads): the SoA arrays, every correspondence at its final position (the order of the serial loop).
    std::vector<int> o1(slots), o2(slots);
    const long long nslots = (long long)slots;
#pragma omp parallel for schedule(static)
    for (long long mpIndex = 0; mpIndex < nslots; ++mpIndex) {
        const MapPoint_& a = v1[mpIndex];
        const MapPoint_& b = v2[mpIndex];
        int i1 = -1, i2 = -1;
        if (a && b) {                                                         // :728-729
            i1 = pMap->isMapPointInKeyFrame(a->getId(), kf1ID);               // :765-768
            i2 = pMap->isMapPointInKeyFrame(b->getId(), kf2ID);
        }
        o1[mpIndex] = (i1 < 0 || i2 < 0) ? -1 : i1;
        o2[mpIndex] = i2;
    }
    std::vector<size_t> at(slots + 1, 0);
    for (size_t k = 0; k < slots; ++k) at[k + 1] = at[k] + (o1[k] >= 0 ? 1 : 0);
    const size_t cnt = at[slots];
    pp.mp1.resize(cnt); pp.mp2.resize(cnt);
    pp.X1.resize(3 * cnt); pp.X2.resize(3 * cnt); pp.uv1.resize(2 * cnt); pp.uv2.resize(2 * cnt);
    pp.isg1.resize(cnt); pp.isg2.resize(cnt); pp.d1.resize(cnt); pp.d2.resize(cnt);
    std::exception_ptr thrown;
#pragma omp parallel for schedule(static)
    for (long long mpIndex = 0; mpIndex < nslots; ++mpIndex) {
        if (o1[mpIndex] < 0) continue;
        try {
            const size_t q = at[mpIndex];
            const int i1 = o1[mpIndex], i2 = o2[mpIndex];
            const MapPoint_& a = v1[mpIndex];
            const MapPoint_& b = v2[mpIndex];
            cv::KeyPoint k1 = pKF1->getKeyPoint((size_t)i1), k2 = pKF2->getKeyPoint((size_t)i2);
            auto p1 = a->getWorldPosition(), p2 = b->getWorldPosition();
            pp.mp1[q] = a; pp.mp2[q] = b;
            for (int k = 0; k < 3; ++k) { pp.X1[3 * q + k] = p1[k]; pp.X2[3 * q + k] = p2[k]; }
            pp.uv1[2 * q] = k1.pt.x; pp.uv1[2 * q + 1] = k1.pt.y;
            pp.uv2[2 * q] = k2.pt.x; pp.uv2[2 * q + 1] = k2.pt.y;
            pp.isg1[q] = pKF1->getInvSigma2(k1.octave);                       // :781
            pp.isg2[q] = pKF2->getInvSigma2(k2.octave);
            // :816,846 read the depth image; simulated key frames only carry per-key-point depths
            pp.d1[q] = img1 ? pKF1->getDepthMeasure(k1.pt.x, k1.pt.y, false) : (double)pKF1->getDepthMeasure((size_t)i1);
            pp.d2[q] = img2 ? pKF2->getDepthMeasure(k2.pt.x, k2.pt.y, false) : (double)pKF2->getDepthMeasure((size_t)i2);
        } catch (...) {
#pragma omp critical
            if (!thrown) thrown = std::current_exception();
        }
    }
    if (thrown) std::rethrow_exception(thrown);
    int n = (int)pp.mp1.size();
    se3_to_7(pMap->getGlobalKeyFramesTransformation(kf1ID, kf2ID), pp.Tg);    // :664 (identity the first time)
    if (!with_mesh) return n > 0;
    if (n < 3) return false;
    // mesh on KF1's positions: 2-D Delaunay of world (x,y), adjacency, cot weights, area (:653-662) -- built on the
    // device from the uploaded points (dsc_set_graph_delaunay, upload_pair); the host triangulator below only serves
    // DSC_HOST_MESH=1 and inputs the device construction refuses (DSC_ERR_GRAPH: e.g. many co-circular points)
    if (!host_mesh_forced()) return true;
    return host_mesh(pp);
}

bool host_mesh(PairProblem& pp) {
    const int n = (int)pp.mp1.size();
    std::vector<double> xy(2 * (size_t)n);
    std::vector<std::array<double, 3>> V(n);
    for (int i = 0; i < n; ++i) {
        xy[2 * i] = (double)pp.X1[3 * i]; xy[2 * i + 1] = (double)pp.X1[3 * i + 1];
        V[i] = {(double)pp.X1[3 * i], (double)pp.X1[3 * i + 1], (double)pp.X1[3 * i + 2]};
    }
    auto tri = dsc_host::Delaunay2D::triangulate(xy.data(), n);
    pp.graph = dsc_host::mesh_graph(V, tri, 0.0);
    return pp.graph.n_triangles > 0 && pp.graph.area > 0.0;
}

// false: the pair has no mesh (fewer than 3 points, all collinear) -- the caller skips it, as gather_pair's callers do
bool upload_pair(PairProblem& pp, bool with_graph = true) {
    dsc_ctx* c = ctx();
    g_res.valid = false;
    dsc_pair pr = make_pair(*pp.kf1, *pp.kf2);
    int n = (int)pp.mp1.size();
    ck(dsc_problem_upload(c, &pr, n, pp.X1.data(), pp.X2.data(), pp.uv1.data(), pp.uv2.data(), pp.d1.data(), pp.d2.data(),
                          pp.isg1.data(), pp.isg2.data(), pp.kf1->getEstimatedDepthScale(), pp.kf2->getEstimatedDepthScale(), pp.Tg),
       "dsc_problem_upload");
    if (!with_graph) return true;
    bool on_device = !host_mesh_forced();
    if (on_device) {
        long long ntri = 0, ne = 0;
        int st = dsc_set_graph_delaunay(c, 0.0, 1, &pp.graph.area, &ntri, &ne);
        pp.graph.n_triangles = ntri;
        pp.mesh_edges = ne;
        if (st == DSC_ERR_GRAPH) {                         // degenerate for the device construction: the host triangulator decides
            on_device = false;
            if (!host_mesh(pp)) return false;
        } else ck(st, "dsc_set_graph_delaunay");
    }
    if (!on_device)
        ck(dsc_set_graph(c, n, pp.graph.rowptr.data(), pp.graph.col.data(), pp.graph.w.data(), pp.graph.area, pp.graph.n_triangles, 1 | 2),
           "dsc_set_graph");
    ck(dsc_compute_rotations(c), "dsc_compute_rotations");                    // :687-688 computeR
    ck(dsc_set_pcg(c, &g_pcg), "dsc_set_pcg");
    // Early rejection of clearly bad LM trials (dsc.h): same trace, same result (tests/test_gpu_parity.py), 7x fewer PCG
    // iterations on large pairs.  DSC_NO_EARLY_REJECT=1 runs every solve to the tolerance, as g2o does.
    static const double rt[2] = {1e-3, 1e-4}, mg[2] = {1.0, 0.5};
    ck(dsc_set_early_reject(c, std::getenv("DSC_NO_EARLY_REJECT") ? 0 : 2, rt, mg), "dsc_set_early_reject");
    return true;
}

dsc_weights weights(double rep, double glob, double arap, double alpha, double beta, float depthError) {
    dsc_weights w{};
    w.rep = rep; w.global = glob; w.arap = arap; w.alpha = alpha; w.beta = beta; w.depth_sigma = depthError;
    return w;
}

void run_lm(const dsc_weights& w, int nIter) {
    std::vector<dsc_iter_record> rec((size_t)std::max(1, nIter));
    dsc_opt_stats st{};
    ck(dsc_optimize(ctx(), &w, nIter, rec.data(), &st), "dsc_optimize");
    g_trace.clear();
    for (int i = 0; i < st.iterations; ++i)
        g_trace.push_back({rec[i].chi2_before, rec[i].chi2_after, rec[i].lambda, rec[i].trials, rec[i].accepted, rec[i].pcg_iters});
}

void write_back(Map* pMap, PairProblem& pp, double* optimizationUpdate) {
    int n = (int)pp.mp1.size();
    std::vector<float> X1(3 * (size_t)n), X2(3 * (size_t)n);
    double scales[2], Tg[7], upd = 0.0;
    ck(dsc_download(ctx(), X1.data(), X2.data(), nullptr, nullptr, scales, Tg, &upd), "dsc_download");
    pp.kf1->setEstimatedDepthScale(scales[0]);                                // :967-972
    pp.kf2->setEstimatedDepthScale(scales[1]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {                                             // :978-990 (distinct MapPoints: independent writes)
        Eigen::Vector3f a(X1[3 * i], X1[3 * i + 1], X1[3 * i + 2]), b(X2[3 * i], X2[3 * i + 1], X2[3 * i + 2]);
        pp.mp1[i]->setWorldPosition(a);
        pp.mp2[i]->setWorldPosition(b);
    }
    if (optimizationUpdate) *optimizationUpdate += upd;
    pMap->insertGlobalKeyFramesTransformation(0, 1, se3_from_7(Tg));          // :1007 (ids hard-coded upstream)
    g_res.valid = true; g_res.map = pMap; g_res.kf1 = pp.kf1Id; g_res.kf2 = pp.kf2Id;
    g_res.mp1 = std::move(pp.mp1); g_res.mp2 = std::move(pp.mp2); g_res.X1 = std::move(X1); g_res.X2 = std::move(X2);   // pp ends here
}

// The weight search's refinements on the batched path: K replicas of the gathered pair (same initial state, mesh and
// rotations) stay resident in one dsc_batch; a Nelder-Mead step refines as many of them as it has candidate weights in
// ONE launch and reads their pixel sigmas (nloptOptimization.cc:5-37 per candidate).  Selected by DSC_WEIGHT_SEARCH=batch
// (see deformationOptimization): on one GPU the sequential search on the main context measured faster.
constexpr int kSearchReplicas = 4;
struct SearchBatch {
    dsc_batch* bt = nullptr;
    ~SearchBatch() { if (bt) dsc_batch_destroy(bt); }
    void bck(int st, const char* what) { if (st != DSC_OK) throw std::runtime_error(std::string(what) + ": " + dsc_batch_last_error(bt)); }
    void upload(PairProblem& pp) {
        const int K = kSearchReplicas, n = (int)pp.mp1.size();
        std::vector<int32_t> rp, col;
        std::vector<double> w;
        if (pp.mesh_edges >= 0) {                                   // the mesh lies on the device: fetch its CSR once
            rp.resize((size_t)n + 1); col.resize((size_t)pp.mesh_edges); w.resize((size_t)pp.mesh_edges);
            ck(dsc_delaunay_download(ctx(), rp.data(), col.data(), w.data()), "dsc_delaunay_download");
        } else { rp = pp.graph.rowptr; col = pp.graph.col; w = pp.graph.w; }
        const size_t E = col.size();
        int dev = 0;
        if (const char* d = std::getenv("DSC_DEVICE")) dev = std::atoi(d);
        if (!bt && dsc_batch_create(dev, &bt) != DSC_OK) throw std::runtime_error("dsc_batch_create failed");
        std::vector<dsc_batch_pair> pairs(K);
        std::vector<long long> po(K + 1), eo(K + 1);
        auto rep = [&](auto& v) { auto one = v; v.reserve(one.size() * K); for (int k = 1; k < K; ++k) v.insert(v.end(), one.begin(), one.end()); };
        std::vector<float> X1 = pp.X1, X2 = pp.X2, uv1 = pp.uv1, uv2 = pp.uv2, i1 = pp.isg1, i2 = pp.isg2;
        std::vector<double> d1 = pp.d1, d2 = pp.d2;
        rep(X1); rep(X2); rep(uv1); rep(uv2); rep(i1); rep(i2); rep(d1); rep(d2); rep(rp); rep(col); rep(w);
        for (int k = 0; k < K; ++k) {
            pairs[k].pair = make_pair(*pp.kf1, *pp.kf2);
            pairs[k].scale1 = pp.kf1->getEstimatedDepthScale(); pairs[k].scale2 = pp.kf2->getEstimatedDepthScale();
            for (int q = 0; q < 7; ++q) pairs[k].Tg7[q] = pp.Tg[q];
            pairs[k].area = pp.graph.area; pairs[k].n_triangles = pp.graph.n_triangles;
            po[k] = (long long)k * n; eo[k] = (long long)k * (long long)E;
        }
        po[K] = (long long)K * n; eo[K] = (long long)K * (long long)E;
        bck(dsc_batch_upload(bt, K, pairs.data(), po.data(), X1.data(), X2.data(), uv1.data(), uv2.data(), d1.data(), d2.data(), i1.data(), i2.data(),
                             eo.data(), rp.data(), col.data(), w.data(), 1), "dsc_batch_upload");
        bck(dsc_batch_set_pcg(bt, &g_pcg), "dsc_batch_set_pcg");
        static const double rt[2] = {1e-3, 1e-4}, mg[2] = {1.0, 0.5};
        bck(dsc_batch_set_early_reject(bt, std::getenv("DSC_NO_EARLY_REJECT") ? 0 : 2, rt, mg), "dsc_batch_set_early_reject");
    }
    // objective (ln sigma_C1)^2 + (ln sigma_C2)^2 of every candidate (rep, global, arap), refined together
    std::vector<double> evaluate(const std::vector<std::vector<double>>& cand, double alpha, double beta, float depthSigma, int nIter) {
        std::vector<double> out;
        for (size_t at = 0; at < cand.size(); at += kSearchReplicas) {
            const int m = (int)std::min((size_t)kSearchReplicas, cand.size() - at);
            std::vector<dsc_weights> w(m);
            for (int k = 0; k < m; ++k) w[k] = weights(cand[at + k][0], cand[at + k][1], cand[at + k][2], alpha, beta, depthSigma);
            std::vector<double> sg(2 * (size_t)m);
            bck(dsc_batch_set_active(bt, m), "dsc_batch_set_active");
            bck(dsc_batch_reset_state(bt), "dsc_batch_reset_state");
            bck(dsc_batch_optimize(bt, w.data(), m, nIter, nullptr, nullptr, nullptr), "dsc_batch_optimize");
            bck(dsc_batch_pixel_sigma(bt, sg.data()), "dsc_batch_pixel_sigma");
            for (int k = 0; k < m; ++k) out.push_back(std::pow(std::log(sg[2 * k]), 2) + std::pow(std::log(sg[2 * k + 1]), 2));
        }
        return out;
    }
};

template <typename F>
void for_each_pair(Map* pMap, F&& f) {                                        // :640-645 (pKF1 = k2, pKF2 = k1)
    auto& kfs = pMap->getKeyFrames();
    for (auto k1 = kfs.begin(); k1 != kfs.end(); ++k1)
        for (auto k2 = std::next(k1); k2 != kfs.end(); ++k2) f(k2->second, k2->first, k1->second, k1->first);
}

// Nelder-Mead over the free coordinates of (rep, global, arap) inside box bounds: the stand-in for
// nlopt::opt(nlopt::LN_NELDERMEAD, 3) (g2oBundleAdjustment.cc:491-515).  NLopt is absent and unpinned; restated:
// fixed coordinates (lb == ub) are eliminated, initial step = NLopt's default rule, standard coefficients
// (1, 2, 0.5, 0.5), points clamped to the box, stop on xtol_rel / xtol_abs of the simplex or maxeval.
//
// The objective is handed over as evalMany(points) -> values: the refinements of one Nelder-Mead step are independent
// given the same initial state, so with `speculative` the step's candidates -- reflection,
// expansion, outside and inside contraction -- are refined TOGETHER (one launch of the batched path), and the d shrink
// points likewise.  The decisions read the values in the order of the sequential algorithm, and only the evaluations that
// algorithm would have made count towards maxeval and appear in `consumed`: both modes walk the same simplices.
struct NmEval { std::vector<double> x; double f; };
template <typename F>
double nelder_mead(std::vector<double>& x, const std::vector<double>& lb, const std::vector<double>& ub, double xtol_rel,
                   double xtol_abs, int maxeval, bool speculative, F&& evalMany, std::vector<NmEval>* consumed = nullptr,
                   int* launches = nullptr) {
    using Pt = std::vector<double>;
    std::vector<int> freeIdx;
    for (size_t i = 0; i < x.size(); ++i) if (ub[i] > lb[i]) freeIdx.push_back((int)i);
    for (size_t i = 0; i < x.size(); ++i) x[i] = std::min(ub[i], std::max(lb[i], x[i]));
    const size_t d = freeIdx.size();
    int evals = 0, nlaunch = 0;
    auto full = [&](const Pt& y) { Pt f = x; for (size_t k = 0; k < d; ++k) f[freeIdx[k]] = y[k]; return f; };
    auto run = [&](const std::vector<Pt>& ys) {                // one launch: all of ys
        std::vector<Pt> fs;
        for (auto& y : ys) fs.push_back(full(y));
        ++nlaunch;
        return evalMany(fs);
    };
    auto consume = [&](const Pt& y, double f) { ++evals; if (consumed) consumed->push_back({full(y), f}); return f; };
    const double inf = std::numeric_limits<double>::infinity();
    if (d == 0) { double f = run({Pt{}})[0]; consume(Pt{}, f); if (launches) *launches = nlaunch; return f; }
    std::vector<Pt> S(d + 1, Pt(d));
    std::vector<double> fv(d + 1, inf);
    for (size_t k = 0; k < d; ++k) S[0][k] = x[freeIdx[k]];
    for (size_t k = 0; k < d; ++k) {
        int i = freeIdx[k];
        double step = (ub[i] - lb[i]) * 0.25;
        if (ub[i] - x[i] < step && ub[i] > x[i]) step = (ub[i] - x[i]) * 0.75;
        if (x[i] - lb[i] < step && x[i] > lb[i]) step = (x[i] - lb[i]) * 0.75;
        if (!(step > 0) || !std::isfinite(step)) step = x[i] != 0 ? std::fabs(x[i]) : 1.0;
        S[k + 1] = S[0];
        S[k + 1][k] = (S[0][k] + step <= ub[i]) ? S[0][k] + step : S[0][k] - step;
    }
    auto clamp = [&](Pt& y) { for (size_t k = 0; k < d; ++k) y[k] = std::min(ub[freeIdx[k]], std::max(lb[freeIdx[k]], y[k])); };
    {   // the initial simplex: d + 1 independent refinements
        size_t m = std::min(d + 1, (size_t)std::max(0, maxeval));
        std::vector<Pt> ys(S.begin(), S.begin() + m);
        if (speculative) { auto f = run(ys); for (size_t k = 0; k < m; ++k) fv[k] = consume(S[k], f[k]); }
        else for (size_t k = 0; k < m; ++k) fv[k] = consume(S[k], run({S[k]})[0]);
    }
    while (evals < maxeval) {
        std::vector<size_t> o(d + 1);
        for (size_t k = 0; k <= d; ++k) o[k] = k;
        std::sort(o.begin(), o.end(), [&](size_t a, size_t b) { return fv[a] < fv[b]; });
        size_t lo = o[0], hi = o[d], nhi = o[d > 0 ? d - 1 : 0];
        bool conv = true;
        for (size_t k = 0; k < d; ++k) {
            double mn = S[0][k], mx = S[0][k];
            for (size_t j = 1; j <= d; ++j) { mn = std::min(mn, S[j][k]); mx = std::max(mx, S[j][k]); }
            if (mx - mn > xtol_abs && mx - mn > xtol_rel * std::fabs(S[lo][k])) conv = false;
        }
        if (conv) break;
        Pt c(d, 0.0);
        for (size_t j = 0; j <= d; ++j) if (j != hi) for (size_t k = 0; k < d; ++k) c[k] += S[j][k] / d;
        auto along = [&](double t) { Pt y(d); for (size_t k = 0; k < d; ++k) y[k] = c[k] + t * (S[hi][k] - c[k]); clamp(y); return y; };
        const Pt xr = along(-1.0), xe = along(-2.0), xco = along(-0.5), xci = along(0.5);
        std::vector<double> spec;                                // values of {xr, xe, xco, xci} when refined together
        if (speculative) spec = run({xr, xe, xco, xci});
        auto value = [&](int which, const Pt& y) { return consume(y, speculative ? spec[which] : run({y})[0]); };
        const double fr = value(0, xr);
        if (fr < fv[lo]) {
            const double fe = evals < maxeval ? value(1, xe) : inf;
            if (fe < fr) { S[hi] = xe; fv[hi] = fe; } else { S[hi] = xr; fv[hi] = fr; }
        } else if (fr < fv[nhi] || (d == 1 && fr < fv[hi])) {
            S[hi] = xr; fv[hi] = fr;
        } else {
            const bool outside = fr < fv[hi];
            const Pt& xc = outside ? xco : xci;
            const double fc = evals < maxeval ? value(outside ? 2 : 3, xc) : inf;
            if (fc < std::min(fr, fv[hi])) { S[hi] = xc; fv[hi] = fc; }
            else {                                               // shrink towards the best vertex: d independent refinements
                std::vector<size_t> js;
                for (size_t j = 0; j <= d; ++j) if (j != lo) js.push_back(j);
                size_t m = std::min(js.size(), (size_t)std::max(0, maxeval - evals));
                for (size_t q = 0; q < m; ++q) for (size_t k = 0; k < d; ++k) S[js[q]][k] = S[lo][k] + 0.5 * (S[js[q]][k] - S[lo][k]);
                if (speculative && m > 0) {
                    std::vector<Pt> ys;
                    for (size_t q = 0; q < m; ++q) ys.push_back(S[js[q]]);
                    auto f = run(ys);
                    for (size_t q = 0; q < m; ++q) fv[js[q]] = consume(S[js[q]], f[q]);
                } else for (size_t q = 0; q < m; ++q) fv[js[q]] = consume(S[js[q]], run({S[js[q]]})[0]);
            }
        }
    }
    size_t best = 0;
    for (size_t k = 1; k <= d; ++k) if (fv[k] < fv[best]) best = k;
    for (size_t k = 0; k < d; ++k) x[freeIdx[k]] = S[best][k];
    if (launches) *launches = nlaunch;
    return fv[best];
}

}  // namespace

// ------------------------------------------------------------------ public API
namespace dsc_host {
const std::vector<LmRecord>& lastTrace() { return g_trace; }
void setSolver(double rtol, int maxIters) { g_pcg.rtol = rtol; g_pcg.max_iters = maxIters; }
const PhaseTimes& lastPhaseTimes() { return g_times; }

TriangulationResult triangulateMatches(KeyFrame& refKF, KeyFrame& currKF, const std::vector<int>& matches, const std::string& method,
                                       const std::string& location, int gate, float minCos, float depthLimit, bool checkReprojection) {
    size_t nRef = refKF.getKeyPoints().size();
    std::vector<int> refIdx;
    std::vector<float> uv1, uv2, d1, d2;
    bool depth = !refKF.getDepthMeasurements().empty() && !currKF.getDepthMeasurements().empty();
    for (size_t i = 0; i < nRef && i < matches.size(); ++i) {
        int j = matches[i];
        if (j < 0) continue;
        refIdx.push_back((int)i);
        cv::Point2f a = refKF.getKeyPoint(i).pt, b = currKF.getKeyPoint((size_t)j).pt;
        uv1.push_back(a.x); uv1.push_back(a.y); uv2.push_back(b.x); uv2.push_back(b.y);
        if (depth) { d1.push_back(refKF.getDepthMeasure(i)); d2.push_back(currKF.getDepthMeasure((size_t)j)); }
    }
    int n = (int)refIdx.size();
    dsc_pair pr = make_pair(refKF, currKF);
    dsc_tri_params tp{method_id(method), location_id(location), gate, minCos, depthLimit, checkReprojection ? 1 : 0};
    std::vector<float> X1(3 * (size_t)n), X2(3 * (size_t)n), cs(n);
    std::vector<uint8_t> valid(n);
    int nv = 0;
    ck(dsc_triangulate(ctx(), &pr, &tp, n, uv1.data(), uv2.data(), depth ? d1.data() : nullptr, depth ? d2.data() : nullptr, X1.data(),
                       X2.data(), valid.data(), cs.data(), &nv),
       "dsc_triangulate");
    TriangulationResult r;
    r.x3D_1.assign(nRef, Eigen::Vector3f()); r.x3D_2.assign(nRef, Eigen::Vector3f());
    r.valid.assign(nRef, 0); r.cosParallax.assign(nRef, 1.f);
    for (int k = 0; k < n; ++k) {
        int i = refIdx[k];
        r.x3D_1[i] = Eigen::Vector3f(X1[3 * k], X1[3 * k + 1], X1[3 * k + 2]);
        r.x3D_2[i] = Eigen::Vector3f(X2[3 * k], X2[3 * k + 1], X2[3 * k + 2]);
        r.valid[i] = valid[k]; r.cosParallax[i] = cs[k];
    }
    r.nValid = nv;
    return r;
}

int triangulateSimulatedMapPoints(Map& map, KeyFrame_ refKF, KeyFrame_ currKF, const std::string& method, const std::string& location, float minCos) {
    size_t n = std::min(refKF->getKeyPoints().size(), currKF->getKeyPoints().size());
    std::vector<int> matches(refKF->getKeyPoints().size(), -1);
    for (size_t i = 0; i < n; ++i) matches[i] = (int)i;                        // Mapping.cc:295-296: ordered pairs
    TriangulationResult r = triangulateMatches(*refKF, *currKF, matches, method, location, DSC_GATE_SIM, minCos);
    int created = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!r.valid[i]) continue;
        MapPoint_ a(new MapPoint(r.x3D_1[i])), b(new MapPoint(r.x3D_2[i]));    // Mapping.cc:329-339
        map.insertMapPoint(a); map.insertMapPoint(b);
        map.addObservation(refKF->getId(), a->getId(), i);
        map.addObservation(currKF->getId(), b->getId(), i);
        refKF->setMapPoint(i, a); currKF->setMapPoint(i, b);
        created += 2;
    }
    // KeyFrame::setInitialDepthScaleInSimulationImages (KeyFrame.cc:131-153), only if the scale is still 0
    if (!refKF->getEstimatedDepthScale()) { double s = 0; ck(dsc_depth_scale_init(ctx(), 1, &s), "dsc_depth_scale_init"); refKF->setEstimatedDepthScale(s); }
    if (!currKF->getEstimatedDepthScale()) { double s = 0; ck(dsc_depth_scale_init(ctx(), 2, &s), "dsc_depth_scale_init"); currKF->setEstimatedDepthScale(s); }
    return created;
}
int initializeMapFromMatches(Map& map, KeyFrame_ refKF, KeyFrame_ currKF, const std::vector<int>& matches, Settings& settings,
                             float* parallaxDegrees) {
    TriangulationResult r = triangulateMatches(*refKF, *currKF, matches, settings.getTrianMethod(), settings.getTrianLocation(), DSC_GATE_REAL,
                                               1.0f, settings.getDepthLimit(), settings.getCheckingSelection());
    std::vector<float> par;
    for (size_t i = 0; i < r.valid.size(); ++i) if (r.valid[i]) par.push_back(r.cosParallax[i]);
    if (par.empty()) return 0;
    std::sort(par.begin(), par.end());                                        // MonocularMapInitializer.cc:375-386
    float cosp = par[std::min<size_t>(50, par.size() - 1)];
    if (cosp < 0.f || cosp > 1.f) return 0;                                    // :379-382 "Parallax must be between 0 and 1"
    const float degrees50 = std::acos(cosp) * (float)(180.0 / M_PI);
    if (parallaxDegrees) *parallaxDegrees = degrees50;
    // the initialiser's acceptance test (:389; Mapping.cc:60 passes Triangulation.minCos as fMinParallax): a pair the
    // reference refuses to initialise from creates nothing here either
    if (!(par.size() >= 25 && degrees50 > settings.getMinCos())) return 0;
    Sophus::SE3f T1w = refKF->getPose(), T2w = currKF->getPose();
    auto usable = [&](size_t i, cv::Point2f& x1, cv::Point2f& x2) {
        if (!r.valid[i]) return false;
        x1 = refKF->getKeyPoint(i).pt; x2 = currKF->getKeyPoint((size_t)matches[i]).pt;
        double d1 = refKF->getDepthMeasure(x1.x, x1.y), d2 = currKF->getDepthMeasure(x2.x, x2.y);
        if (d1 <= 0.0 || d2 <= 0.0) return false;                             // Mapping.cc:194-200
        if (x1.x <= 0.1 || x1.x >= 1500 || x1.y <= 0.1 || x1.y >= 1500) return false;
        if (x2.x <= 0.1 || x2.x >= 1500 || x2.y <= 0.1 || x2.y >= 1500) return false;
        return true;
    };
    // The initial depth scales are means over the points whose parallax exceeds the setting (:211-254); with no such point
    // the reference divides by zero and refines with a NaN scale.  Refused here: nothing is created.
    {
        size_t n_scale = 0;
        cv::Point2f x1, x2;
        for (size_t i = 0; i < r.valid.size(); ++i)
            if (usable(i, x1, x2) && std::acos(r.cosParallax[i]) * (float)(180.0 / M_PI) > settings.getMinCos()) ++n_scale;
        if (n_scale == 0) return 0;
    }
    int created = 0;
    double scale1 = 0, scale2 = 0;
    float n_points = 0;
    for (size_t i = 0; i < r.valid.size(); ++i) {
        cv::Point2f x1, x2;
        if (!usable(i, x1, x2)) continue;
        MapPoint_ a(new MapPoint(r.x3D_1[i])), b(new MapPoint(r.x3D_2[i]));
        map.insertMapPoint(a); map.insertMapPoint(b);
        map.addObservation(refKF->getId(), a->getId(), i);                    // :205-209
        map.addObservation(currKF->getId(), b->getId(), (size_t)matches[i]);
        refKF->setMapPoint(i, a); currKF->setMapPoint(i, b);
        float degrees = std::acos(r.cosParallax[i]) * (float)(180.0 / M_PI);
        if (degrees > settings.getMinCos()) {                                 // :220-233 (sic: degrees against minCos)
            Eigen::Vector3f c1 = T1w * r.x3D_1[i], c2 = T2w * r.x3D_2[i];
            scale1 += refKF->getDepthMeasure(x1.x, x1.y, false) / c1[2];
            scale2 += currKF->getDepthMeasure(x2.x, x2.y, false) / c2[2];
            n_points++;
        }
        created += 2;
    }
    refKF->setEstimatedDepthScale(scale1 / n_points);                         // :250-254
    currKF->setEstimatedDepthScale(scale2 / n_points);
    return created;
}
}  // namespace dsc_host

#ifndef DSC_IN_REFERENCE_TREE
void KeyFrame::setInitialDepthScaleInSimulationImages() {
    throw std::logic_error("use dsc_host::triangulateSimulatedMapPoints: the initial depth scale is reduced on the GPU");
}
#endif

bool useTriangulationMethod(const Eigen::Vector3f& xn1, const Eigen::Vector3f& xn2, const Sophus::SE3f& T1w, const Sophus::SE3f& T2w,
                            Eigen::Vector3f& x3D_1, Eigen::Vector3f& x3D_2, std::string TrianMethod, std::string TrianLocation) {
    float T1[12], T2[12];
    dsc_host::pose34(T1w, T1);
    dsc_host::pose34(T2w, T2);
    const float a[3] = {xn1[0], xn1[1], xn1[2]}, b[3] = {xn2[0], xn2[1], xn2[2]};
    float X1[3], X2[3];
    ck(dsc_triangulate_rays(ctx(), T1, T2, method_id(TrianMethod), location_id(TrianLocation), 1, a, b, X1, X2), "dsc_triangulate_rays");
    x3D_1 = Eigen::Vector3f(X1[0], X1[1], X1[2]);
    x3D_2 = Eigen::Vector3f(X2[0], X2[1], X2[2]);
    return true;                                                               // Geometry.cc:229
}

void arapOptimization(Map* pMap, double repBalanceWeight, double globalBalanceWeight, double arapBalanceWeight, double alphaWeight,
                      double betaWeight, float DepthError, int nOptIterations, double* optimizationUpdate) {
    if (optimizationUpdate) *optimizationUpdate = 0;                            // :974-976
    dsc_weights w = weights(repBalanceWeight, globalBalanceWeight, arapBalanceWeight, alphaWeight, betaWeight, DepthError);
    // The reference puts every key-frame pair in one g2o graph; its executables only ever hold two key frames
    // (the main loops stop after the first mapped pair), so pairs are refined one after the other here.
    for_each_pair(pMap, [&](KeyFrame_ kf1, ID id1, KeyFrame_ kf2, ID id2) {
        PairProblem pp;
        double t0 = now_ms();
        if (!gather_pair(pMap, kf1, id1, kf2, id2, pp)) return;
        double t1 = now_ms();
        if (!upload_pair(pp)) return;
        dsc_synchronize(ctx());
        double t2 = now_ms();
        run_lm(w, nOptIterations);
        double t3 = now_ms();
        const long long npts = (long long)pp.mp1.size();
        write_back(pMap, pp, optimizationUpdate);
        double t4 = now_ms();
        g_times = dsc_host::PhaseTimes{t1 - t0, t2 - t1, t3 - t2, t4 - t3, npts};
    });
}

void calculatePixelsStandDev(std::shared_ptr<Map> map, PixelsError& pe) {
    for_each_pair(map.get(), [&](KeyFrame_ kf1, ID id1, KeyFrame_ kf2, ID id2) {
        // The statistic needs the map points and the key points only: when the device still holds this pair's
        // refinement (the usual call order: arapOptimization, then the statistic), it is read from there; otherwise
        // points and observations are uploaded without building the mesh.
        double s[2];
        if (resident_pair(map.get(), kf1, id1, kf2, id2)) {
            ck(dsc_pixel_sigma(ctx(), s), "dsc_pixel_sigma");
            pe.desvc1 = s[0]; pe.desvc2 = s[1]; pe.desv = 0.5 * (s[0] + s[1]);
            return;
        }
        PairProblem pp;
        if (!gather_pair(map.get(), kf1, id1, kf2, id2, pp, /*with_mesh=*/false)) return;
        upload_pair(pp, /*with_graph=*/false);
        ck(dsc_pixel_sigma(ctx(), s), "dsc_pixel_sigma");
        pe.desvc1 = s[0]; pe.desvc2 = s[1]; pe.desv = 0.5 * (s[0] + s[1]);
    });
}

void deformationOptimization(std::shared_ptr<Map> pMap, Settings& settings, std::shared_ptr<MapVisualizer>& mapVisualizer,
                             const std::vector<Eigen::Vector3f> /*originalPoints*/, const std::vector<Eigen::Vector3f> /*movedPoints*/) {
    float depthSigma = settings.getSimulatedDepthWeight() / 1000;               // :449
    double rep = settings.getOptRepWeight(), arap = settings.getOptArapWeight(), glob = settings.getOptGlobalWeight();
    double alpha = settings.getOptAlphaWeight(), beta = settings.getOptBetaWeight();
    std::string sel = settings.getOptSelection(), wsel = settings.getOptWeightsSelection();
    int nOptimizations = settings.getnOptimizations(), nOptIterations = settings.getnOptIterations();
    bool drawRays = settings.getDrawRaysSelection();
    size_t nMapPoints = pMap->getMapPoints().size();
    double optimizationUpdate = 100;
    for (int i = 1; i <= nOptimizations && optimizationUpdate >= (0.0001 * nMapPoints); i++) {   // :482
        if (sel == "open3DArap") {
            throw std::runtime_error("Optimization.selection open3DArap (Open3D DeformAsRigidAsPossible) is outside the accelerated path");
        } else if (sel == "twoOptimizations" && wsel == "nlopt") {
            // weight search: every objective evaluation re-runs the refinement from the uploaded state
            // (dsc_reset_state) instead of cloning the Map (nloptOptimization.cc:5-37)
            std::vector<double> x = {rep, glob, arap};
            std::vector<double> lb = {settings.getNloptRepLowerBound(), settings.getNloptGlobalLowerBound(), settings.getNloptArapLowerBound()};
            std::vector<double> ub = {settings.getNloptRepUpperBound(), settings.getNloptGlobalUpperBound(), settings.getNloptArapUpperBound()};
            bool uploaded = false;
            PairProblem pp;
            for_each_pair(pMap.get(), [&](KeyFrame_ kf1, ID id1, KeyFrame_ kf2, ID id2) {
                if (uploaded) return;
                PairProblem cand;
                if (!gather_pair(pMap.get(), kf1, id1, kf2, id2, cand)) return;
                if (!upload_pair(cand)) return;
                pp = std::move(cand);
                uploaded = true;
            });
            if (uploaded) {
                // Default (= DSC_WEIGHT_SEARCH=single): one refinement at a time on the main context -- on ONE GPU a refinement
                // already has the whole device (dense factorisation up to 600 correspondences, one-launch solves above), and it
                // measured fastest at every size tried.  DSC_WEIGHT_SEARCH=batch: the candidates of a step
                // together as replicas on the batched path (the arrangement for spare SMs / more GPUs; same walk, tested);
                // =sequential: one at a time on the batched path.
                const char* mode = std::getenv("DSC_WEIGHT_SEARCH");
                const bool on_batch = mode && (std::string(mode) == "batch" || std::string(mode) == "sequential");
                const bool speculative = on_batch && !(mode && std::string(mode) == "sequential");
                SearchBatch sb;
                if (on_batch) sb.upload(pp);
                std::vector<NmEval> log;
                int launches = 0;
                double minf = nelder_mead(x, lb, ub, settings.getNloptRelTolerance(), settings.getNloptAbsTolerance(),
                                          settings.getNloptnOptimizations(), speculative, [&](const std::vector<std::vector<double>>& ys) {
                                              if (on_batch) return sb.evaluate(ys, alpha, beta, depthSigma, nOptIterations);
                                              std::vector<double> f;
                                              for (auto& y : ys) {
                                                  ck(dsc_reset_state(ctx()), "dsc_reset_state");
                                                  run_lm(weights(y[0], y[1], y[2], alpha, beta, depthSigma), nOptIterations);
                                                  double sg[2];
                                                  ck(dsc_pixel_sigma(ctx(), sg), "dsc_pixel_sigma");
                                                  f.push_back(std::pow(std::log(sg[0]), 2) + std::pow(std::log(sg[1]), 2));
                                              }
                                              return f;
                                          }, &log, &launches);
                for (auto& e : log)                                            // outerObjective's own lines, in its order
                    std::cout << "Current x values. Reprojection: " << e.x[0] << ", Global T: " << e.x[1] << ", ARAP: " << e.x[2] << "\nerror: " << e.f << "\n";
                std::cout << "Weight search: " << log.size() << " objective evaluations in " << launches << " launches ("
                          << (speculative ? "candidates of a step refined together" : on_batch ? "one at a time, batched path" : "one at a time") << ")\n";
                std::cout << "\nWEIGHTS OPTIMIZED\nOptimized repBalanceWeight: " << x[0] << "\nOptimized globalBalanceWeight: " << x[1]
                          << "\nOptimized arapBalanceWeight: " << x[2] << "\nFinal minimized ABSOLUTE error: " << minf << std::endl;
            }
            arapOptimization(pMap.get(), x[0], x[1], x[2], alpha, beta, depthSigma, nOptIterations, &optimizationUpdate);   // :525
            rep = x[0]; glob = x[1]; arap = x[2];
        } else {
            // "g2oArap": the refinement with the configured weights (:566-569).
            // "twoOptimizations" with weightsSelection "eigen" (:532-563): the reference hands Eigen::LevenbergMarquardt a
            // functor that declares 2 residuals (EigenOptimization.h:31, Functor<double>(2, 2)) for the 3 unknowns
            // (rep, global, arap).  Eigen's LevenbergMarquardt::minimizeInit refuses m < n ("if (n <= 0 || m < n || ...)
            // return ImproperInputParameters", unsupported/Eigen/src/NonLinearOptimization/LevenbergMarquardt.h) before the
            // first function evaluation, so minimize() returns 0 with x untouched and the reference then runs its "final
            // optimization with optimized weights" on the CONFIGURED weights.  That is reproduced literally: no search,
            // the same messages, one refinement.
            if (sel == "twoOptimizations" && wsel != "nlopt") {
                std::cout << "Return code: 0" << std::endl;                      // LevenbergMarquardtSpace::ImproperInputParameters
                std::cout << "\nWEIGHTS OPTIMIZED\nOptimized repBalanceWeight: " << rep << "\nOptimized globalBalanceWeight: " << glob
                          << "\nOptimized arapBalanceWeight: " << arap << "\n\nFinal optimization with optimized weights:\n" << std::endl;
            }
            arapOptimization(pMap.get(), rep, glob, arap, alpha, beta, depthSigma, nOptIterations, &optimizationUpdate);
        }
        std::cout << "\nOptimization COMPLETED... " << i << " / " << nOptimizations << " iterations.\nOptimization change: "
                  << optimizationUpdate << std::endl;
        if (mapVisualizer) mapVisualizer->update(drawRays);
    }
    if (mapVisualizer) mapVisualizer->update(drawRays);
}

// ------------------------------------------------------------------ classic bundle adjustment (g2oBundleAdjustment.cc:38-444) on dsc_ba_*
namespace {
dsc_camera ba_camera(CameraModel* c) {
    dsc_camera o{};
    o.model = model_id(c);
    for (int i = 0; i < 8; ++i) o.params[i] = i < c->getNumberOfParameters() ? c->getParameter(i) : 0.f;
    return o;
}
const double kHuber2D = (double)(float)std::sqrt(5.99);                        // const float thHuber2D = sqrt(5.99)
struct BaHandle {
    dsc_ba* h = nullptr;
    BaHandle() {
        int dev = 0;
        if (const char* d = std::getenv("DSC_DEVICE")) dev = std::atoi(d);
        int st = dsc_ba_create(dev, &h);
        if (st != DSC_OK) throw std::runtime_error(std::string("dsc_ba_create: ") + dsc_status_string(st) + " (there is no CPU fallback)");
    }
    ~BaHandle() { if (h) dsc_ba_destroy(h); }
    void ck(int st, const char* what) { if (st != DSC_OK) throw std::runtime_error(std::string(what) + ": " + dsc_ba_last_error(h)); }
};
// vertices and edges of bundleAdjustment / localBundleAdjustment in the reference's creation order
struct BaGraph {
    std::vector<KeyFrame_> kfs;                 // poses: optimised key frames first, then the fixed ones
    std::vector<ID> kfIds;
    std::vector<uint8_t> fixed;
    std::vector<MapPoint_> mps;
    std::unordered_map<MapPoint*, int> mpIndex;
    std::vector<int32_t> obsPose, obsPoint;
    std::vector<float> uv, isg;
    std::vector<size_t> obsSlot;                // vMpInKf
    void addKeyFrame(KeyFrame_ kf, ID id, bool fix, const std::set<ID>* onlyPoints) {
        const int k = (int)kfs.size();
        kfs.push_back(kf); kfIds.push_back(id); fixed.push_back(fix ? 1 : 0);
        auto& vMPs = kf->getMapPoints();
        for (size_t i = 0; i < vMPs.size(); ++i) {
            MapPoint_ mp = vMPs[i];
            if (!mp) continue;
            if (onlyPoints && !onlyPoints->count(mp->getId())) continue;   // :381 fixed key frames: local points only
            auto it = mpIndex.find(mp.get());
            int j;
            if (it == mpIndex.end()) {
                if (onlyPoints) continue;                                   // (:383 asserts the point is already a vertex)
                j = (int)mps.size(); mps.push_back(mp); mpIndex[mp.get()] = j;
            } else j = it->second;
            cv::KeyPoint kp = kf->getKeyPoint(i);
            obsPose.push_back(k); obsPoint.push_back(j); obsSlot.push_back(i);
            uv.push_back(kp.pt.x); uv.push_back(kp.pt.y);
            isg.push_back(kf->getInvSigma2(kp.octave));
        }
    }
    void upload(BaHandle& b) {
        std::vector<double> p7(7 * kfs.size()), X(3 * mps.size());
        std::vector<dsc_camera> cams(kfs.size());
        for (size_t k = 0; k < kfs.size(); ++k) { se3_to_7(kfs[k]->getPose(), p7.data() + 7 * k); cams[k] = ba_camera(kfs[k]->getCalibration().get()); }
        for (size_t j = 0; j < mps.size(); ++j) { auto p = mps[j]->getWorldPosition(); for (int c = 0; c < 3; ++c) X[3 * j + c] = (double)p[c]; }
        b.ck(dsc_ba_upload(b.h, (int)kfs.size(), p7.data(), fixed.data(), cams.data(), (int)mps.size(), X.data(), 0, (long long)obsPose.size(),
                           obsPose.data(), obsPoint.data(), uv.data(), isg.data()), "dsc_ba_upload");
    }
    void writeBack(BaHandle& b) {                                               // :126-140, :429-443
        std::vector<double> p7(7 * kfs.size()), X(3 * mps.size());
        b.ck(dsc_ba_download(b.h, p7.data(), X.data()), "dsc_ba_download");
        for (size_t k = 0; k < kfs.size(); ++k) {
            if (fixed[k] && kfIds[k] != 0) continue;                           // fixed key frames of the local map are not written
            Sophus::SE3f T = se3_from_7(p7.data() + 7 * k);
            kfs[k]->setPose(T);
        }
        for (size_t j = 0; j < mps.size(); ++j) {
            Eigen::Vector3f p((float)X[3 * j], (float)X[3 * j + 1], (float)X[3 * j + 2]);
            mps[j]->setWorldPosition(p);
        }
    }
};
}  // namespace

void bundleAdjustment(Map* pMap) {
    BaGraph g;
    for (auto& kv : pMap->getKeyFrames()) g.addKeyFrame(kv.second, kv.first, kv.second->getId() == 0, nullptr);   // :63-118
    if (g.kfs.empty()) return;
    BaHandle b;
    g.upload(b);
    b.ck(dsc_ba_optimize(b.h, 20, kHuber2D, nullptr, nullptr), "dsc_ba_optimize");                              // :122-123
    g.writeBack(b);
}

void localBundleAdjustment(Map* pMap, ID currKeyFrameId) {
    std::set<ID> sLocalMapPoints, sLocalKeyFrames, sFixedKeyFrames;
    pMap->getLocalMapOfKeyFrame(currKeyFrameId, sLocalMapPoints, sLocalKeyFrames, sFixedKeyFrames);              // :252
    BaGraph g;
    for (ID id : sLocalKeyFrames) { KeyFrame_ kf = pMap->getKeyFrame(id); g.addKeyFrame(kf, id, kf->getId() == 0, nullptr); }      // :280-344
    for (ID id : sFixedKeyFrames) g.addKeyFrame(pMap->getKeyFrame(id), id, true, &sLocalMapPoints);                                // :347-398
    if (g.kfs.empty()) return;
    BaHandle b;
    g.upload(b);
    const size_t O = g.obsPose.size();
    std::vector<double> chiA(O), chiB(O);
    std::vector<uint8_t> posA(O), posB(O), active(O);
    b.ck(dsc_ba_optimize(b.h, 5, kHuber2D, nullptr, nullptr), "dsc_ba_optimize");                               // :401-402
    b.ck(dsc_ba_edge_chi2(b.h, chiA.data(), posA.data()), "dsc_ba_edge_chi2");
    for (size_t e = 0; e < O; ++e) active[e] = (chiA[e] > 5.991 || !posA[e]) ? 0 : 1;                             // :405-413 (+ every kernel off)
    b.ck(dsc_ba_set_levels(b.h, active.data()), "dsc_ba_set_levels");
    b.ck(dsc_ba_optimize(b.h, 10, 0.0, nullptr, nullptr), "dsc_ba_optimize");                                   // :415-416
    b.ck(dsc_ba_edge_chi2(b.h, chiB.data(), posB.data()), "dsc_ba_edge_chi2");
    for (size_t e = 0; e < O; ++e) {                                                                              // :419-428
        // a level-1 edge was not evaluated by the second run: its chi2() is the one it was classified with
        const double chi = active[e] ? chiB[e] : chiA[e];
        if (chi > 5.991 || !posB[e]) {
            const ID kfId = g.kfIds[(size_t)g.obsPose[e]], mpId = g.mps[(size_t)g.obsPoint[e]]->getId();
            pMap->getKeyFrame(kfId)->setMapPoint(g.obsSlot[e], nullptr);
            pMap->removeObservation(kfId, mpId);
            pMap->checkKeyFrame(kfId);
        }
    }
    g.writeBack(b);
}

int poseOnlyOptimization(Frame& currFrame) {
    auto& vMapPoints = currFrame.getMapPoints();
    std::vector<size_t> slot;
    std::vector<int32_t> obsPose, obsPoint;
    std::vector<float> uv, isg;
    std::vector<double> X;
    for (size_t i = 0; i < vMapPoints.size(); ++i) {                                                              // :168-197
        MapPoint_ mp = vMapPoints[i];
        if (!mp) continue;
        cv::KeyPoint kp = currFrame.getKeyPoint(i);
        auto p = mp->getWorldPosition();
        obsPose.push_back(0); obsPoint.push_back((int32_t)slot.size()); slot.push_back(i);
        uv.push_back(kp.pt.x); uv.push_back(kp.pt.y);
        isg.push_back(currFrame.getInvSigma2(kp.octave));
        for (int c = 0; c < 3; ++c) X.push_back((double)p[c]);
    }
    const size_t O = slot.size();
    double start7[7];
    se3_to_7(currFrame.getPose(), start7);
    dsc_camera cam = ba_camera(currFrame.getCalibration().get());
    uint8_t notFixed = 0;
    BaHandle b;
    b.ck(dsc_ba_upload(b.h, 1, start7, &notFixed, &cam, (int)O, X.data(), /*points_fixed=*/1, (long long)O, obsPose.data(), obsPoint.data(), uv.data(),
                       isg.data()), "dsc_ba_upload");
    // vInlier is indexed by key-point slot upstream (size = all slots, false where there is no map point)
    std::vector<bool> vInlier(vMapPoints.size(), false);
    for (size_t e = 0; e < O; ++e) vInlier[slot[e]] = true;
    std::vector<uint8_t> level0(O, 1);
    std::vector<double> stored(O), fresh(O);
    b.ck(dsc_ba_edge_chi2(b.h, stored.data(), nullptr), "dsc_ba_edge_chi2");
    double delta = kHuber2D;
    for (int round = 0; round < 4; ++round) {                                                                     // :199-228
        b.ck(dsc_ba_set_poses(b.h, start7), "dsc_ba_set_poses");                                               // every round restarts from fPose
        b.ck(dsc_ba_set_levels(b.h, level0.data()), "dsc_ba_set_levels");
        b.ck(dsc_ba_optimize(b.h, 10, delta, nullptr, nullptr), "dsc_ba_optimize");
        b.ck(dsc_ba_edge_chi2(b.h, fresh.data(), nullptr), "dsc_ba_edge_chi2");
        // `if(!vInlier[mpIndex]) e->computeError();` with mpIndex = the ROUND number (:209-210): edges at level 1 keep
        // the error of their last evaluation unless the flag of key point #round happens to be false
        const bool refreshAll = (size_t)round < vInlier.size() && !vInlier[(size_t)round];
        for (size_t e = 0; e < O; ++e) {
            if (level0[e] || refreshAll) stored[e] = fresh[e];
            const bool in = !(stored[e] > 5.991);
            vInlier[slot[e]] = in;
            level0[e] = in ? 1 : 0;
        }
        if (round == 2) delta = 0.0;                                                                              // :222-224 setRobustKernel(0)
    }
    int nGood = 0;
    for (size_t i = 0; i < vInlier.size(); ++i) {                                                                 // :231-239
        if (!vInlier[i]) currFrame.setMapPoint(i, nullptr);
        else nGood++;
    }
    double p7[7];
    b.ck(dsc_ba_download(b.h, p7, nullptr), "dsc_ba_download");
    Sophus::SE3f T = se3_from_7(p7);
    currFrame.setPose(T);
    return nGood;
}

// ------------------------------------------------------------------ C hooks for the tests (ctypes)
extern "C" int dsch_delaunay(int n, const double* xy, int* tri_out, int max_tri) {
    auto tri = dsc_host::Delaunay2D::triangulate(xy, n);
    int m = (int)tri.size();
    for (int k = 0; k < m && k < max_tri; ++k) { tri_out[3 * k] = tri[k][0]; tri_out[3 * k + 1] = tri[k][1]; tri_out[3 * k + 2] = tri[k][2]; }
    return m;
}

extern "C" int dsch_mesh_graph(int n, const double* xyz, int ntri, const int* tri, int* rowptr, int* col, double* w, int max_e, double* area) {
    std::vector<std::array<double, 3>> V(n);
    for (int i = 0; i < n; ++i) V[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    std::vector<std::array<int, 3>> T(ntri);
    for (int k = 0; k < ntri; ++k) T[k] = {tri[3 * k], tri[3 * k + 1], tri[3 * k + 2]};
    auto g = dsc_host::mesh_graph(V, T, 0.0);
    for (int i = 0; i <= n; ++i) rowptr[i] = g.rowptr[i];
    int E = (int)g.col.size();
    for (int e = 0; e < E && e < max_e; ++e) { col[e] = g.col[e]; w[e] = g.w[e]; }
    *area = g.area;
    return E;
}

// Nelder-Mead on analytic objectives, both ways of evaluating a step (tests/test_host_shim.py): log[k] = (x[dim], f) of the
// k-th evaluation the sequential algorithm makes.  fn 0: sum_k (log10 x_k - (k + 1))^2;  fn 1: Rosenbrock (dim 2).
extern "C" int dsch_nelder_mead(int fn, int dim, const double* x0, const double* lb, const double* ub, double xtol_rel, double xtol_abs,
                                int maxeval, int speculative, double* xout, double* fout, double* log, int maxlog, int* launches) {
    std::vector<double> x(x0, x0 + dim), l(lb, lb + dim), u(ub, ub + dim);
    auto f = [&](const std::vector<double>& y) {
        if (fn == 1) return 100.0 * std::pow(y[1] - y[0] * y[0], 2) + std::pow(1.0 - y[0], 2);
        double a = 0.0;
        for (int k = 0; k < dim; ++k) a += std::pow(std::log10(y[k]) - (k + 1), 2);
        return a;
    };
    std::vector<NmEval> used;
    int nl = 0;
    double best = nelder_mead(x, l, u, xtol_rel, xtol_abs, maxeval, speculative != 0, [&](const std::vector<std::vector<double>>& ys) {
        std::vector<double> v;
        for (auto& y : ys) v.push_back(f(y));
        return v;
    }, &used, &nl);
    for (int k = 0; k < dim; ++k) xout[k] = x[k];
    *fout = best;
    if (launches) *launches = nl;
    for (size_t e = 0; e < used.size() && (int)e < maxlog; ++e) {
        for (int k = 0; k < dim; ++k) log[e * (dim + 1) + k] = used[e].x[k];
        log[e * (dim + 1) + dim] = used[e].f;
    }
    return (int)used.size();
}

#ifndef DSC_IN_REFERENCE_TREE
// The classic bundle-adjustment entry points on a Map built from arrays (tests/test_gpu_ba.py): mode 0 bundleAdjustment, 1
// localBundleAdjustment(curr), 2 poseOnlyOptimization of a Frame with the pose and observations of pose 0.  pose34: [K][12]
// float [R|t]; X: [M][3] float; observations (pose, point, uv, octave).  Out: poses as 7 doubles of the written-back float
// poses, points, per observation whether its map point was taken out of its key frame / frame, mode 2: the inlier count.
extern "C" int dsch_ba_flow(int mode, int K, const float* pose34, const float* cam8, int M, const float* X, int O, const int* obs_pose,
                            const int* obs_point, const float* obs_uv, const int* obs_octave, int curr, float minCommonObs, double* pose7_out,
                            float* X_out, unsigned char* removed_out, int* n_good) {
    try {
        KeyFrame::restartIds(); MapPoint::restartIds();
        std::vector<float> cp(cam8, cam8 + 8);
        auto calib = std::make_shared<KannalaBrandt8>(cp);
        auto poseOf = [&](int k) {
            Eigen::Matrix3f R;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = pose34[12 * k + 4 * r + c];
            return Sophus::SE3f(R, Eigen::Vector3f(pose34[12 * k + 3], pose34[12 * k + 7], pose34[12 * k + 11]));
        };
        std::vector<MapPoint_> mps(M);
        for (int j = 0; j < M; ++j) { Eigen::Vector3f p(X[3 * j], X[3 * j + 1], X[3 * j + 2]); mps[j] = std::make_shared<MapPoint>(p); }
        std::vector<std::vector<int>> obsOf(K);
        for (int e = 0; e < O; ++e) obsOf[obs_pose[e]].push_back(e);
        auto keysOf = [&](int k) {
            std::vector<cv::KeyPoint> keys;
            for (int e : obsOf[k]) { cv::KeyPoint kp(cv::Point2f(obs_uv[2 * e], obs_uv[2 * e + 1]), 1.0f); kp.octave = obs_octave[e]; keys.push_back(kp); }
            return keys;
        };
        if (n_good) *n_good = 0;
        if (mode == 2) {
            Frame f(keysOf(0), poseOf(0), calib);
            for (size_t s = 0; s < obsOf[0].size(); ++s) f.setMapPoint(s, mps[obs_point[obsOf[0][s]]]);
            int ng = poseOnlyOptimization(f);
            if (n_good) *n_good = ng;
            se3_to_7(f.getPose(), pose7_out);
            for (size_t s = 0; s < obsOf[0].size(); ++s) removed_out[obsOf[0][s]] = f.getMapPoints()[s] ? 0 : 1;
            return 0;
        }
        Map map;
        map.setMinCommonObs(minCommonObs);
        std::vector<KeyFrame_> kfs(K);
        for (int j = 0; j < M; ++j) map.insertMapPoint(mps[j]);
        for (int k = 0; k < K; ++k) {
            kfs[k] = std::make_shared<KeyFrame>(keysOf(k), poseOf(k), calib);
            map.insertKeyFrame(kfs[k]);
            for (size_t s = 0; s < obsOf[k].size(); ++s) {
                MapPoint_ mp = mps[obs_point[obsOf[k][s]]];
                kfs[k]->setMapPoint(s, mp);
                map.addObservation(kfs[k]->getId(), mp->getId(), s);
            }
        }
        if (mode == 0) bundleAdjustment(&map); else localBundleAdjustment(&map, kfs[curr]->getId());
        for (int k = 0; k < K; ++k) {
            se3_to_7(kfs[k]->getPose(), pose7_out + 7 * k);
            for (size_t s = 0; s < obsOf[k].size(); ++s) removed_out[obsOf[k][s]] = kfs[k]->getMapPoints()[s] ? 0 : 1;
        }
        for (int j = 0; j < M; ++j) { auto p = mps[j]->getWorldPosition(); for (int c = 0; c < 3; ++c) X_out[3 * j + c] = p[c]; }
        return 0;
    } catch (const std::exception& ex) {
        std::cerr << "dsch_ba_flow: " << ex.what() << std::endl;
        return 1;
    }
}
#endif

#ifndef DSC_IN_REFERENCE_TREE
// Map::getLocalMapOfKeyFrame / addObservation / removeObservation of host/Map.h on a Map given as observation lists
// (tests/test_host_shim.py, CPU): obs = (key frame index, map point index); out_* are 0/1 flags per key frame / point.
extern "C" int dsch_local_map(int K, int M, int O, const int* obs_kf, const int* obs_mp, int n_remove, const int* remove_obs, int curr, float minCommonObs,
                              unsigned char* local_kf, unsigned char* fixed_kf, unsigned char* local_mp) {
    KeyFrame::restartIds(); MapPoint::restartIds();
    Map map;
    map.setMinCommonObs(minCommonObs);
    std::vector<float> cp = {500.f, 500.f, 320.f, 240.f, 0.f, 0.f, 0.f, 0.f};
    auto calib = std::make_shared<KannalaBrandt8>(cp);
    std::vector<KeyFrame_> kfs(K);
    std::vector<MapPoint_> mps(M);
    for (int j = 0; j < M; ++j) { Eigen::Vector3f p(0.f, 0.f, 1.f); mps[j] = std::make_shared<MapPoint>(p); map.insertMapPoint(mps[j]); }
    std::vector<int> slots(K, 0);
    for (int e = 0; e < O; ++e) slots[obs_kf[e]]++;
    for (int k = 0; k < K; ++k) {
        kfs[k] = std::make_shared<KeyFrame>(std::vector<cv::KeyPoint>((size_t)slots[k]), Sophus::SE3f(), calib);
        map.insertKeyFrame(kfs[k]);
    }
    std::vector<int> next(K, 0);
    for (int e = 0; e < O; ++e) map.addObservation(kfs[obs_kf[e]]->getId(), mps[obs_mp[e]]->getId(), (size_t)next[obs_kf[e]]++);
    for (int r = 0; r < n_remove; ++r) map.removeObservation(kfs[obs_kf[remove_obs[r]]]->getId(), mps[obs_mp[remove_obs[r]]]->getId());
    std::set<ID> lmp, lkf, fkf;
    map.getLocalMapOfKeyFrame(kfs[curr]->getId(), lmp, lkf, fkf);
    for (int k = 0; k < K; ++k) { local_kf[k] = lkf.count(kfs[k]->getId()) ? 1 : 0; fixed_kf[k] = fkf.count(kfs[k]->getId()) ? 1 : 0; }
    for (int j = 0; j < M; ++j) local_mp[j] = lmp.count(mps[j]->getId()) ? 1 : 0;
    return 0;
}
#endif
