// dsc_ba_api.cuh -- C ABI of the classic bundle-adjustment paths (include/dsc.h: dsc_ba_*), included by dsc_api.cu.
// Host side: set-up (observations sorted by point, the entry list of the Schur complement), the g2o Levenberg-Marquardt
// loop (same control flow as dsc_optimize), the fold of the per-chunk partials and the Cholesky factorisation of the
// reduced camera system (6 x free poses: a few dozen unknowns).  Kernels: dsc_ba.cuh.
#pragma once
#include "dsc_ba.cuh"

struct dsc_ba {
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int K = 0, M = 0, Kf = 0;
    long long O = 0, E = 0;
    bool points_fixed = false, have = false;
    std::vector<double> pose7;                 // current estimate [K][7] (q xyzw, t)
    std::vector<unsigned char> fixed;
    std::vector<int> free_col;                 // pose -> first column of its 6 unknowns in the reduced system, -1 if fixed
    std::vector<int> perm;                     // sorted observation -> caller's index
    std::vector<dsc::BaEntryChunk> chunks;
    std::vector<int> chunk_a, chunk_b;         // poses of the segment each chunk belongs to
    // device
    dsc::BaPose *d_pose = nullptr, *d_pose_t = nullptr;
    dsc::CamF* d_cam = nullptr;
    unsigned char *d_free = nullptr, *d_act = nullptr, *d_pos = nullptr;
    double4 *d_X = nullptr, *d_Xt = nullptr;
    int *d_ptr = nullptr, *d_opose = nullptr, *d_orig = nullptr, *d_slot = nullptr, *d_ea = nullptr, *d_eb = nullptr, *d_ept = nullptr;
    float2* d_uv = nullptr;
    float* d_isg = nullptr;
    double *d_Hll = nullptr, *d_bl = nullptr, *d_W = nullptr, *d_A = nullptr, *d_g = nullptr, *d_part = nullptr, *d_dP = nullptr, *d_chi = nullptr;
    dsc::BaEntryChunk* d_chunks = nullptr;
    size_t part_cap = 0;
    size_t capK = 0, capM = 0, capO = 0, capE = 0, capC = 0;   // grow-only capacities (no cudaMalloc / cudaFree on a warm handle)
    std::vector<double> h_part;
    std::vector<dsc::BaPose> h_pose[2];
    unsigned h_pose_next = 0;
    long long launches = 0;
};

namespace {
constexpr int kBaMaxFreePoses = 512;
int bafail(dsc_ba* b, int code, const std::string& what) {
    if (b) b->err = std::string(status_str(code)) + ": " + what;
    return code;
}
#define ACK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return bafail(ba, DSC_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(e_)); \
    } while (0)

void ba_pose_to_dev(const double* p7, dsc::BaPose& o) {
    dsc::quat_to_rot(p7, o.R);
    o.t[0] = p7[4]; o.t[1] = p7[5]; o.t[2] = p7[6];
}
int ba_grid(const dsc_ba* b, long long n) {
    return (int)std::max(1LL, std::min((n + dsc::kThreads - 1) / dsc::kThreads, (long long)b->sms * 8));
}
int ba_upload_poses(dsc_ba* ba, const std::vector<double>& p7, dsc::BaPose* dst) {
    // (a copy from pageable memory is staged before the call returns; the staging vector is a member and is only rewritten
    // by the next call, which the stream-ordered consumers of this one precede)
    std::vector<dsc::BaPose>& hp = ba->h_pose[ba->h_pose_next++ & 1];
    hp.resize((size_t)ba->K);
    for (int k = 0; k < ba->K; ++k) ba_pose_to_dev(p7.data() + 7 * (size_t)k, hp[k]);
    ACK(cudaMemcpyAsync(dst, hp.data(), sizeof(dsc::BaPose) * (size_t)ba->K, cudaMemcpyHostToDevice, ba->stream));
    return DSC_OK;
}
// in-place Cholesky solve of the dense SPD system A x = b (row-major n x n); false if A is not positive definite
bool ba_chol_solve(std::vector<double>& A, std::vector<double>& b, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        d = std::sqrt(d);
        A[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
    for (int i = 0; i < n; ++i) if (!std::isfinite(b[i])) return false;
    return true;
}
}  // namespace

extern "C" int dsc_ba_create(int device, dsc_ba** out) {
    if (!out) return DSC_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return DSC_ERR_NO_DEVICE; }
    if (device < 0 || device >= count) return DSC_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return DSC_ERR_CUDA;
    dsc_ba* ba = new dsc_ba();
    ba->device = device;
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, device);
    ba->sms = prop.multiProcessorCount;
    if (cudaStreamCreate(&ba->stream) != cudaSuccess || cudaEventCreate(&ba->ev0) != cudaSuccess || cudaEventCreate(&ba->ev1) != cudaSuccess) {
        delete ba;
        return DSC_ERR_CUDA;
    }
    *out = ba;
    return DSC_OK;
}

extern "C" void dsc_ba_destroy(dsc_ba* ba) {
    if (!ba) return;
    cudaSetDevice(ba->device);
    dev_free(ba->d_pose); dev_free(ba->d_pose_t); dev_free(ba->d_cam); dev_free(ba->d_free); dev_free(ba->d_act); dev_free(ba->d_pos);
    dev_free(ba->d_X); dev_free(ba->d_Xt); dev_free(ba->d_ptr); dev_free(ba->d_opose); dev_free(ba->d_orig); dev_free(ba->d_slot); dev_free(ba->d_ea);
    dev_free(ba->d_eb); dev_free(ba->d_ept); dev_free(ba->d_uv); dev_free(ba->d_isg); dev_free(ba->d_Hll); dev_free(ba->d_bl);
    dev_free(ba->d_W); dev_free(ba->d_A); dev_free(ba->d_g); dev_free(ba->d_part); dev_free(ba->d_dP); dev_free(ba->d_chi); dev_free(ba->d_chunks);
    if (ba->ev0) cudaEventDestroy(ba->ev0);
    if (ba->ev1) cudaEventDestroy(ba->ev1);
    if (ba->stream) cudaStreamDestroy(ba->stream);
    delete ba;
}

extern "C" const char* dsc_ba_last_error(const dsc_ba* ba) { return ba ? ba->err.c_str() : "null handle"; }

extern "C" int dsc_ba_upload(dsc_ba* ba, int n_poses, const double* poses7, const uint8_t* pose_fixed, const dsc_camera* cams, int n_points,
                             const double* X, int points_fixed, long long n_obs, const int32_t* obs_pose, const int32_t* obs_point,
                             const float* obs_uv, const float* obs_inv_sigma2) {
    if (!ba || n_poses < 1 || n_points < 0 || n_obs < 0 || !poses7 || !cams || (n_points > 0 && !X) || (n_obs > 0 && (!obs_pose || !obs_point || !obs_uv)))
        return bafail(ba, DSC_ERR_INVALID_ARG, "dsc_ba_upload");
    if (n_obs > 0x7fffffffLL) return bafail(ba, DSC_ERR_INVALID_ARG, "too many observations");
    ACK(cudaSetDevice(ba->device));
    ba->have = false;
    const int K = n_poses, M = n_points;
    const long long O = n_obs;
    for (long long o = 0; o < O; ++o)
        if (obs_pose[o] < 0 || obs_pose[o] >= K || obs_point[o] < 0 || obs_point[o] >= M) return bafail(ba, DSC_ERR_INVALID_ARG, "observation out of range");
    ba->K = K; ba->M = M; ba->O = O; ba->points_fixed = points_fixed != 0;
    ba->pose7.assign(poses7, poses7 + 7 * (size_t)K);
    for (int k = 0; k < K; ++k) {                                   // SE3Quat normalises its rotation
        double* q = ba->pose7.data() + 7 * (size_t)k;
        double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (!(nrm > 0.0) || !std::isfinite(nrm)) return bafail(ba, DSC_ERR_INVALID_ARG, "pose quaternion");
        double s = (q[3] < 0 ? -1.0 : 1.0) / nrm;
        for (int c = 0; c < 4; ++c) q[c] *= s;
    }
    ba->fixed.assign((size_t)K, 0);
    ba->free_col.assign((size_t)K, -1);
    ba->Kf = 0;
    for (int k = 0; k < K; ++k) {
        ba->fixed[k] = pose_fixed && pose_fixed[k] ? 1 : 0;
        if (!ba->fixed[k]) ba->free_col[k] = 6 * ba->Kf++;
    }
    // the reduced camera system is folded and factorised on the host (dense, 6 x free poses): fine for the local maps of the
    // reference (a handful of key frames), refused beyond kBaMaxFreePoses rather than silently taking minutes per trial
    if (ba->Kf > kBaMaxFreePoses)
        return bafail(ba, DSC_ERR_INVALID_ARG, "more than " + std::to_string(kBaMaxFreePoses) + " free poses: the host factorisation of the reduced system is not meant for that");
    // ---- observations sorted by (point, pose): CSR over the points
    std::vector<int> ptr((size_t)M + 1, 0);
    for (long long o = 0; o < O; ++o) ptr[(size_t)obs_point[o] + 1]++;
    for (int j = 0; j < M; ++j) ptr[(size_t)j + 1] += ptr[j];
    ba->perm.assign((size_t)O, 0);
    {
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        for (long long o = 0; o < O; ++o) ba->perm[(size_t)cur[obs_point[o]]++] = (int)o;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < M; ++j)
            if (ptr[(size_t)j + 1] - ptr[j] > 1)
                std::sort(ba->perm.begin() + ptr[j], ba->perm.begin() + ptr[(size_t)j + 1],
                          [&](int a, int b) { return obs_pose[a] != obs_pose[b] ? obs_pose[a] < obs_pose[b] : a < b; });
    }
    std::vector<int> opose((size_t)O);
    std::vector<float2> uv((size_t)O);
    std::vector<float> isg((size_t)O);
#pragma omp parallel for schedule(static)
    for (long long s = 0; s < O; ++s) {
        const int o = ba->perm[(size_t)s];
        opose[(size_t)s] = obs_pose[o];
        uv[(size_t)s] = make_float2(obs_uv[2 * (size_t)o], obs_uv[2 * (size_t)o + 1]);
        isg[(size_t)s] = obs_inv_sigma2 ? obs_inv_sigma2[o] : 1.0f;
    }
    // ---- storage slots of the per-observation blocks, rank-major: all first observations in point order, then all second ones, ...
    std::vector<int> slot((size_t)O);
    {
        int maxdeg = 0;
        for (int j = 0; j < M; ++j) maxdeg = std::max(maxdeg, ptr[(size_t)j + 1] - ptr[j]);
        std::vector<long long> rank_count((size_t)maxdeg + 1, 0);
        for (int j = 0; j < M; ++j) rank_count[(size_t)(ptr[(size_t)j + 1] - ptr[j])]++;       // points of each degree
        std::vector<long long> rank_base((size_t)maxdeg + 1, 0);                              // first slot of rank s
        long long alive = M - rank_count[0], at = 0;
        for (int sdx = 0; sdx < maxdeg; ++sdx) { rank_base[sdx] = at; at += alive; alive -= rank_count[(size_t)sdx + 1]; }
        std::vector<long long> next(rank_base.begin(), rank_base.end());
        for (int j = 0; j < M; ++j)
            for (int sdx = 0; sdx < ptr[(size_t)j + 1] - ptr[j]; ++sdx) slot[(size_t)ptr[j] + sdx] = (int)next[sdx]++;
    }
    // ---- entries of the reduced system: pairs (a <= b) of observations of one point from FREE poses, sorted by (a, b);
    // counted per point, placed by a prefix sum (threads), then a stable counting sort by pair
    struct Ent { int a, b, oa, ob, pt; };
    std::vector<size_t> eoff((size_t)M + 1, 0);
    int twice = 0;
#pragma omp parallel for schedule(static) reduction(| : twice)
    for (int j = 0; j < M; ++j) {
        size_t f = 0;
        for (int s = ptr[j]; s < ptr[(size_t)j + 1]; ++s) {
            if (s > ptr[j] && opose[(size_t)s] == opose[(size_t)s - 1]) twice |= 1;
            if (!ba->fixed[opose[(size_t)s]]) ++f;
        }
        eoff[(size_t)j + 1] = ba->points_fixed ? f : f * (f + 1) / 2;
    }
    if (twice) return bafail(ba, DSC_ERR_INVALID_ARG, "a point is observed twice from the same pose");
    for (int j = 0; j < M; ++j) eoff[(size_t)j + 1] += eoff[j];
    if (eoff[M] > (size_t)0x7fffffff) return bafail(ba, DSC_ERR_INVALID_ARG, "too many observation pairs for 32-bit entry indices");
    std::vector<Ent> ents(eoff[M]);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; ++j) {
        size_t at = eoff[j];
        for (int s = ptr[j]; s < ptr[(size_t)j + 1]; ++s) {
            if (ba->fixed[opose[(size_t)s]]) continue;
            for (int t = s; t < ptr[(size_t)j + 1]; ++t) {
                if (ba->fixed[opose[(size_t)t]]) continue;
                if (t > s && ba->points_fixed) continue;            // no coupling between poses without free points
                ents[at++] = Ent{opose[(size_t)s], opose[(size_t)t], s, t, j};
            }
        }
    }
    {   // stable counting sort by (a, b): the entries of a segment stay in point order
        std::vector<size_t> start((size_t)K * K + 1, 0);
        for (const Ent& e : ents) start[(size_t)e.a * K + e.b + 1]++;
        for (size_t q = 0; q < (size_t)K * K; ++q) start[q + 1] += start[q];
        std::vector<Ent> sorted(ents.size());
        for (const Ent& e : ents) sorted[start[(size_t)e.a * K + e.b]++] = e;
        ents.swap(sorted);
    }
    ba->E = (long long)ents.size();
    std::vector<int> ea(ents.size()), eb(ents.size()), ept(ents.size());
    ba->chunks.clear(); ba->chunk_a.clear(); ba->chunk_b.clear();
    for (size_t e = 0; e < ents.size();) {
        size_t f = e;
        while (f < ents.size() && ents[f].a == ents[e].a && ents[f].b == ents[e].b) ++f;
        for (size_t c = e; c < f; c += dsc::kBaChunk) {
            ba->chunks.push_back({(int)c, (int)std::min(f, c + (size_t)dsc::kBaChunk), ents[e].a == ents[e].b ? 1 : 0, 0});
            ba->chunk_a.push_back(ents[e].a); ba->chunk_b.push_back(ents[e].b);
        }
        e = f;
    }
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)ents.size(); ++e) { ea[(size_t)e] = slot[(size_t)ents[(size_t)e].oa]; eb[(size_t)e] = slot[(size_t)ents[(size_t)e].ob]; ept[(size_t)e] = ents[(size_t)e].pt; }
    // ---- device buffers
    const size_t Os = (size_t)std::max<long long>(O, 1), Ms = (size_t)std::max(M, 1), Es = std::max<size_t>(ents.size(), 1);
    const size_t nch = std::max<size_t>(ba->chunks.size(), 1);
    if ((size_t)K > ba->capK) {
        ACK(dev_alloc(ba->d_pose, (size_t)K)); ACK(dev_alloc(ba->d_pose_t, (size_t)K)); ACK(dev_alloc(ba->d_cam, (size_t)K));
        ACK(dev_alloc(ba->d_free, (size_t)K)); ACK(dev_alloc(ba->d_dP, 6 * (size_t)K));
        ba->capK = (size_t)K;
    }
    if (Ms > ba->capM) {
        const size_t c = Ms + Ms / 8;
        ACK(dev_alloc(ba->d_X, c)); ACK(dev_alloc(ba->d_Xt, c)); ACK(dev_alloc(ba->d_ptr, c + 1)); ACK(dev_alloc(ba->d_Hll, 6 * c)); ACK(dev_alloc(ba->d_bl, 3 * c));
        ba->capM = c;
    }
    if (Os > ba->capO) {
        const size_t c = Os + Os / 8;
        ACK(dev_alloc(ba->d_act, c)); ACK(dev_alloc(ba->d_pos, c)); ACK(dev_alloc(ba->d_chi, c)); ACK(dev_alloc(ba->d_opose, c)); ACK(dev_alloc(ba->d_orig, c)); ACK(dev_alloc(ba->d_slot, c));
        ACK(dev_alloc(ba->d_uv, c)); ACK(dev_alloc(ba->d_isg, c)); ACK(dev_alloc(ba->d_W, 18 * c)); ACK(dev_alloc(ba->d_A, 21 * c)); ACK(dev_alloc(ba->d_g, 6 * c));
        ba->capO = c;
    }
    if (Es > ba->capE) {
        const size_t c = Es + Es / 8;
        ACK(dev_alloc(ba->d_ea, c)); ACK(dev_alloc(ba->d_eb, c)); ACK(dev_alloc(ba->d_ept, c));
        ba->capE = c;
    }
    if (nch > ba->capC) { ACK(dev_alloc(ba->d_chunks, nch + nch / 8)); ba->capC = nch + nch / 8; }
    {
        const size_t want = std::max<size_t>(nch * dsc::kBaDiag, (size_t)ba->sms * 8 * 2 + 16);
        if (want > ba->part_cap) { ACK(dev_alloc(ba->d_part, want + want / 8)); ba->part_cap = want + want / 8; }
    }
    ba->h_part.assign(ba->part_cap, 0.0);
    std::vector<dsc::CamF> hc((size_t)K);
    std::vector<unsigned char> hf((size_t)K);
    for (int k = 0; k < K; ++k) {
        hc[k].model = cams[k].model;
        for (int c = 0; c < 8; ++c) hc[k].p[c] = cams[k].params[c];
        hf[k] = ba->fixed[k] ? 0 : 1;
    }
    std::vector<double4> hx((size_t)M);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; ++j) hx[j] = make_double4(X[3 * (size_t)j], X[3 * (size_t)j + 1], X[3 * (size_t)j + 2], 0.0);
    ACK(cudaMemcpyAsync(ba->d_cam, hc.data(), sizeof(dsc::CamF) * (size_t)K, cudaMemcpyHostToDevice, ba->stream));
    ACK(cudaMemcpyAsync(ba->d_free, hf.data(), (size_t)K, cudaMemcpyHostToDevice, ba->stream));
    if (M) ACK(cudaMemcpyAsync(ba->d_X, hx.data(), sizeof(double4) * (size_t)M, cudaMemcpyHostToDevice, ba->stream));
    ACK(cudaMemcpyAsync(ba->d_ptr, ptr.data(), sizeof(int) * ((size_t)M + 1), cudaMemcpyHostToDevice, ba->stream));
    if (O) {
        ACK(cudaMemcpyAsync(ba->d_opose, opose.data(), sizeof(int) * (size_t)O, cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_orig, ba->perm.data(), sizeof(int) * (size_t)O, cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_slot, slot.data(), sizeof(int) * (size_t)O, cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_uv, uv.data(), sizeof(float2) * (size_t)O, cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_isg, isg.data(), sizeof(float) * (size_t)O, cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemsetAsync(ba->d_act, 1, (size_t)O, ba->stream));
    }
    if (!ents.empty()) {
        ACK(cudaMemcpyAsync(ba->d_ea, ea.data(), sizeof(int) * ents.size(), cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_eb, eb.data(), sizeof(int) * ents.size(), cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_ept, ept.data(), sizeof(int) * ents.size(), cudaMemcpyHostToDevice, ba->stream));
        ACK(cudaMemcpyAsync(ba->d_chunks, ba->chunks.data(), sizeof(dsc::BaEntryChunk) * ba->chunks.size(), cudaMemcpyHostToDevice, ba->stream));
    }
    ACK(cudaStreamSynchronize(ba->stream));
    int rc = ba_upload_poses(ba, ba->pose7, ba->d_pose);
    if (rc) return rc;
    ba->have = true;
    return DSC_OK;
}

extern "C" int dsc_ba_set_poses(dsc_ba* ba, const double* poses7) {
    if (!ba || !poses7) return DSC_ERR_INVALID_ARG;
    if (!ba->have) return bafail(ba, DSC_ERR_STATE, "dsc_ba_set_poses before dsc_ba_upload");
    ACK(cudaSetDevice(ba->device));
    ba->pose7.assign(poses7, poses7 + 7 * (size_t)ba->K);
    for (int k = 0; k < ba->K; ++k) {
        double* q = ba->pose7.data() + 7 * (size_t)k;
        double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        if (!(nrm > 0.0) || !std::isfinite(nrm)) return bafail(ba, DSC_ERR_INVALID_ARG, "pose quaternion");
        double s = (q[3] < 0 ? -1.0 : 1.0) / nrm;
        for (int c = 0; c < 4; ++c) q[c] *= s;
    }
    return ba_upload_poses(ba, ba->pose7, ba->d_pose);
}

extern "C" int dsc_ba_set_levels(dsc_ba* ba, const uint8_t* active) {
    if (!ba) return DSC_ERR_INVALID_ARG;
    if (!ba->have) return bafail(ba, DSC_ERR_STATE, "dsc_ba_set_levels before dsc_ba_upload");
    ACK(cudaSetDevice(ba->device));
    if (ba->O == 0) return DSC_OK;
    if (!active) { ACK(cudaMemsetAsync(ba->d_act, 1, (size_t)ba->O, ba->stream)); ACK(cudaStreamSynchronize(ba->stream)); return DSC_OK; }
    std::vector<unsigned char> h((size_t)ba->O);
    for (long long s = 0; s < ba->O; ++s) h[(size_t)s] = active[ba->perm[(size_t)s]] ? 1 : 0;
    ACK(cudaMemcpyAsync(ba->d_act, h.data(), (size_t)ba->O, cudaMemcpyHostToDevice, ba->stream));
    ACK(cudaStreamSynchronize(ba->stream));
    return DSC_OK;
}

namespace {
// activeRobustChi2 of (X, poses)
int ba_cost(dsc_ba* ba, const double4* X, const dsc::BaPose* poses, double delta, double* chi2) {
    const int nb = ba_grid(ba, ba->M);
    dsc::ba_cost_kernel<<<nb, dsc::kThreads, 0, ba->stream>>>(ba->M, ba->d_ptr, ba->d_opose, ba->d_uv, ba->d_isg, ba->d_act, X, poses, ba->d_cam, delta, ba->d_part);
    ba->launches++;
    ACK(cudaGetLastError());
    ACK(cudaMemcpyAsync(ba->h_part.data(), ba->d_part, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, ba->stream));
    ACK(cudaStreamSynchronize(ba->stream));
    *chi2 = host_sum(ba->h_part.data(), nb);
    return DSC_OK;
}
}  // namespace

extern "C" int dsc_ba_optimize(dsc_ba* ba, int n_iters, double huber_delta, dsc_iter_record* records, dsc_opt_stats* stats) {
    if (!ba || n_iters < 0) return DSC_ERR_INVALID_ARG;
    if (!ba->have) return bafail(ba, DSC_ERR_STATE, "dsc_ba_optimize before dsc_ba_upload");
    ACK(cudaSetDevice(ba->device));
    const int K = ba->K, M = ba->M, n = 6 * ba->Kf;
    const double delta = huber_delta > 0.0 ? huber_delta : 0.0;
    const int pf = ba->points_fixed ? 1 : 0;
    dsc_opt_stats st{};
    const long long launches0 = ba->launches;
    ACK(cudaEventRecord(ba->ev0, ba->stream));
    double lambda = 0.0, ni = 2.0;
    const int nbm = ba_grid(ba, M), nch = (int)ba->chunks.size();
    std::vector<double> Hpp((size_t)n * n), bp((size_t)n), S((size_t)n * n), rhs((size_t)n), WHW((size_t)n * n), WHb((size_t)n);
    for (int it = 0; it < n_iters; ++it) {
        // ---- linearise at the current estimate
        dsc::ba_linearize_kernel<<<nbm, dsc::kThreads, 0, ba->stream>>>(M, ba->d_ptr, ba->d_opose, ba->d_uv, ba->d_isg, ba->d_act, ba->d_slot, ba->d_X, ba->d_pose,
                                                                     ba->d_cam, ba->d_free, delta, pf, (size_t)std::max<long long>(ba->O, 1), ba->d_Hll, ba->d_bl, ba->d_W, ba->d_A, ba->d_g, ba->d_part);
        ba->launches++;
        ACK(cudaGetLastError());
        ACK(cudaMemcpyAsync(ba->h_part.data(), ba->d_part, sizeof(double) * 2 * (size_t)nbm, cudaMemcpyDeviceToHost, ba->stream));
        ACK(cudaStreamSynchronize(ba->stream));
        double current = host_sum(ba->h_part.data(), nbm, 2, 0), maxdiag = 0.0;
        for (int b = 0; b < nbm; ++b) maxdiag = std::max(maxdiag, ba->h_part[2 * (size_t)b + 1]);
        if (!std::isfinite(current)) return bafail(ba, DSC_ERR_NONFINITE, "cost is not finite");
        int trials = 0, accepted = 0;
        const double chi_before = current;
        double rho = 0.0, lambda_start = lambda;
        bool have_hpp = false;
        for (;;) {
            // ---- reduced camera system at this lambda: fold the chunk partials in order
            bool ok = true;
            if (nch > 0) {
                dsc::ba_schur_kernel<<<nch, dsc::kThreads, 0, ba->stream>>>(ba->d_chunks, ba->d_ea, ba->d_eb, ba->d_ept, ba->d_Hll, ba->d_bl, ba->d_W, ba->d_A, ba->d_g,
                                                                         lambda, pf ? 0 : 1, (size_t)std::max<long long>(ba->O, 1), ba->d_part);
                ba->launches++;
                ACK(cudaGetLastError());
                ACK(cudaMemcpyAsync(ba->h_part.data(), ba->d_part, sizeof(double) * (size_t)nch * dsc::kBaDiag, cudaMemcpyDeviceToHost, ba->stream));
                ACK(cudaStreamSynchronize(ba->stream));
            }
            std::fill(Hpp.begin(), Hpp.end(), 0.0); std::fill(bp.begin(), bp.end(), 0.0);
            std::fill(WHW.begin(), WHW.end(), 0.0); std::fill(WHb.begin(), WHb.end(), 0.0);
            for (int c = 0; c < nch; ++c) {
                const double* p = ba->h_part.data() + (size_t)c * dsc::kBaDiag;
                const int ca = ba->free_col[ba->chunk_a[c]], cb = ba->free_col[ba->chunk_b[c]];
                if (ba->chunks[c].diag) {
                    int q = 0;
                    for (int r = 0; r < 6; ++r)
                        for (int cc = r; cc < 6; ++cc, ++q) {
                            Hpp[(size_t)(ca + r) * n + ca + cc] += p[q]; WHW[(size_t)(ca + r) * n + ca + cc] += p[27 + q];
                            if (cc != r) { Hpp[(size_t)(ca + cc) * n + ca + r] += p[q]; WHW[(size_t)(ca + cc) * n + ca + r] += p[27 + q]; }
                        }
                    for (int r = 0; r < 6; ++r) { bp[ca + r] += p[21 + r]; WHb[ca + r] += p[48 + r]; }
                } else {
                    for (int r = 0; r < 6; ++r)
                        for (int cc = 0; cc < 6; ++cc) { WHW[(size_t)(ca + r) * n + cb + cc] += p[r * 6 + cc]; WHW[(size_t)(cb + cc) * n + ca + r] += p[r * 6 + cc]; }
                }
            }
            if (it == 0 && !have_hpp) {                           // computeLambdaInit: tau * largest diagonal entry of H
                for (int r = 0; r < n; ++r) maxdiag = std::max(maxdiag, std::fabs(Hpp[(size_t)r * n + r]));
                lambda = 1e-5 * maxdiag;
                ni = 2.0;
                lambda_start = lambda;
                have_hpp = true;
                continue;                                          // (the pass above ran with lambda = 0: only its A / g sums were needed)
            }
            have_hpp = true;
            for (int r = 0; r < n; ++r) {
                for (int c = 0; c < n; ++c) S[(size_t)r * n + c] = Hpp[(size_t)r * n + c] - WHW[(size_t)r * n + c];
                S[(size_t)r * n + r] += lambda;
                rhs[r] = bp[r] - WHb[r];
            }
            std::vector<double> dp = rhs;
            if (n > 0) { std::vector<double> Sc = S; ok = ba_chol_solve(Sc, dp, n); }
            double temp = std::numeric_limits<double>::max(), scale = 1e-3;
            std::vector<double> trial7 = ba->pose7;
            if (ok) {
                std::vector<double> dP(6 * (size_t)K, 0.0);
                double sp = 0.0;
                for (int k = 0; k < K; ++k) {
                    const int c0 = ba->free_col[k];
                    if (c0 < 0) continue;
                    for (int q = 0; q < 6; ++q) { dP[6 * (size_t)k + q] = dp[c0 + q]; sp += dp[c0 + q] * (lambda * dp[c0 + q] + bp[c0 + q]); }
                    dsc::se3_oplus(ba->pose7.data() + 7 * (size_t)k, dp.data() + c0, trial7.data() + 7 * (size_t)k);
                }
                ACK(cudaMemcpyAsync(ba->d_dP, dP.data(), sizeof(double) * dP.size(), cudaMemcpyHostToDevice, ba->stream));
                int rc = ba_upload_poses(ba, trial7, ba->d_pose_t);
                if (rc) return rc;
                // back-substitution of the points and the trial's cost in one pass
                dsc::ba_backsub_kernel<<<nbm, dsc::kThreads, 0, ba->stream>>>(M, ba->d_ptr, ba->d_opose, ba->d_slot, ba->d_X, ba->d_Hll, ba->d_bl, ba->d_W, ba->d_dP, lambda, pf,
                                                                           (size_t)std::max<long long>(ba->O, 1), ba->d_Xt, ba->d_uv, ba->d_isg, ba->d_act, ba->d_pose_t,
                                                                           ba->d_cam, delta, ba->d_part);
                ba->launches++;
                ACK(cudaGetLastError());
                ACK(cudaMemcpyAsync(ba->h_part.data(), ba->d_part, sizeof(double) * 2 * (size_t)nbm, cudaMemcpyDeviceToHost, ba->stream));
                ACK(cudaStreamSynchronize(ba->stream));
                scale = sp + host_sum(ba->h_part.data(), nbm, 2, 0) + 1e-3;
                temp = host_sum(ba->h_part.data(), nbm, 2, 1);
                if (!std::isfinite(temp)) temp = std::numeric_limits<double>::max();
            }
            rho = (current - temp) / scale;
            if (rho > 0 && std::isfinite(temp) && temp < std::numeric_limits<double>::max()) {
                double alpha = 1.0 - std::pow(2.0 * rho - 1.0, 3);
                alpha = std::min(alpha, 2.0 / 3.0);
                lambda *= std::max(1.0 / 3.0, alpha);
                ni = 2.0;
                current = temp;
                ba->pose7 = trial7;
                std::swap(ba->d_X, ba->d_Xt);
                std::swap(ba->d_pose, ba->d_pose_t);
                accepted = 1;
            } else {
                lambda *= ni;
                ni *= 2.0;
            }
            ++trials;
            if (!(rho < 0 && trials < 10)) break;
        }
        st.iterations = it + 1;
        st.total_trials += trials;
        if (records) {
            dsc_iter_record& r = records[it];
            r.chi2_before = chi_before; r.lambda = lambda_start; r.trials = trials; r.accepted = accepted; r.pcg_iters = 0; r.chi2_after = current;
        }
        if (trials == 10 || rho == 0) { st.terminated = 1; break; }
    }
    ACK(cudaEventRecord(ba->ev1, ba->stream));
    double fin = 0.0;
    { int rc = ba_cost(ba, ba->d_X, ba->d_pose, delta, &fin); if (rc) return rc; }
    float ms = 0.f;
    ACK(cudaEventElapsedTime(&ms, ba->ev0, ba->ev1));
    st.final_chi2 = fin; st.device_ms = (double)ms; st.kernel_launches = (int)(ba->launches - launches0);
    if (stats) *stats = st;
    return DSC_OK;
}

extern "C" int dsc_ba_edge_chi2(dsc_ba* ba, double* chi2, uint8_t* depth_positive) {
    if (!ba) return DSC_ERR_INVALID_ARG;
    if (!ba->have) return bafail(ba, DSC_ERR_STATE, "dsc_ba_edge_chi2 before dsc_ba_upload");
    ACK(cudaSetDevice(ba->device));
    if (ba->O == 0) return DSC_OK;
    dsc::ba_edge_kernel<<<ba_grid(ba, ba->M), dsc::kThreads, 0, ba->stream>>>(ba->M, ba->d_ptr, ba->d_opose, ba->d_uv, ba->d_isg, ba->d_orig, ba->d_X, ba->d_pose,
                                                                         ba->d_cam, ba->d_chi, ba->d_pos);
    ba->launches++;
    ACK(cudaGetLastError());
    if (chi2) ACK(cudaMemcpyAsync(chi2, ba->d_chi, sizeof(double) * (size_t)ba->O, cudaMemcpyDeviceToHost, ba->stream));
    if (depth_positive) ACK(cudaMemcpyAsync(depth_positive, ba->d_pos, (size_t)ba->O, cudaMemcpyDeviceToHost, ba->stream));
    ACK(cudaStreamSynchronize(ba->stream));
    return DSC_OK;
}

extern "C" int dsc_ba_download(dsc_ba* ba, double* poses7, double* X) {
    if (!ba) return DSC_ERR_INVALID_ARG;
    if (!ba->have) return bafail(ba, DSC_ERR_STATE, "dsc_ba_download before dsc_ba_upload");
    ACK(cudaSetDevice(ba->device));
    if (poses7) std::memcpy(poses7, ba->pose7.data(), sizeof(double) * 7 * (size_t)ba->K);
    if (X && ba->M) {
        std::vector<double4> hx((size_t)ba->M);
        ACK(cudaMemcpyAsync(hx.data(), ba->d_X, sizeof(double4) * (size_t)ba->M, cudaMemcpyDeviceToHost, ba->stream));
        ACK(cudaStreamSynchronize(ba->stream));
        for (int j = 0; j < ba->M; ++j) { X[3 * (size_t)j] = hx[j].x; X[3 * (size_t)j + 1] = hx[j].y; X[3 * (size_t)j + 2] = hx[j].z; }
    }
    return DSC_OK;
}

extern "C" int dsc_ba_launch_count(const dsc_ba* ba, long long* count) {
    if (!ba || !count) return DSC_ERR_INVALID_ARG;
    *count = ba->launches;
    return DSC_OK;
}
