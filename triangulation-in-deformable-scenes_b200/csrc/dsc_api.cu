// dsc_api.cu -- context, C ABI (include/dsc.h) and the Levenberg-Marquardt host loop.
//
// The LM control flow restates g2o::OptimizationAlgorithmLevenberg::solve as configured by the
// reference at Modules/Optimization/g2oBundleAdjustment.cc:619-628,959-962 (upstream g2o
// core/optimization_algorithm_levenberg.cpp; the library itself is not vendored by the reference):
// lambda_0 = 1e-5 max|diag H|, rho = (chi - chi_new) / (dx.(lambda dx + b) + 1e-3), good step
// lambda *= max(1/3, min(2/3, 1 - (2 rho - 1)^3)), bad step lambda *= ni, ni *= 2, <= 10 trials.
// All state stays in HBM; per trial the host reads back two scalars.
#include "../../include/dsc.h"
#include "dsc_kernels.cuh"
#include "dsc_kernels_ell.cuh"
#include "dsc_knn.cuh"
#include "dsc_delaunay.cuh"
#include "dsc_graph.cuh"
#include "dsc_dense.cuh"
#include "dsc_small.cuh"
#include "dsc_batch.cuh"
#include "dsc_shard_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <parallel/algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace dsc;

#define DSC_VERSION 200

struct dsc_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;
    cudaEvent_t evo[4] = {nullptr, nullptr, nullptr, nullptr};      // dsc_optimize: begin, end, phase a, phase b
    std::string err;
    long long launches = 0;

    // ---- triangulation stage
    int tn = 0, tcap = 0;
    bool t_has_depth = false;
    PairDev tpair{};
    float2 *t_uv1 = nullptr, *t_uv2 = nullptr;
    float *t_d1 = nullptr, *t_d2 = nullptr, *t_X1 = nullptr, *t_X2 = nullptr, *t_cos = nullptr;
    uint8_t* t_valid = nullptr;
    bool t_done = false;

    // ---- refinement problem
    int n = 0, cap = 0;
    long long E = 0;
    bool have_problem = false, have_graph = false, have_rot = false;
    PairDev pair{};
    double area = 1.0;
    long long ntri = 0;
    Globals g0{};                         // uploaded globals
    std::vector<int> perm;                // internal index -> caller index
    // pinned bounce buffer of the host -> device copies of PAGEABLE caller memory (two halves, see h2d)
    unsigned char* bounce[2] = {nullptr, nullptr};
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    unsigned bounce_next = 0;
    float* d_box = nullptr;                       // [4 + 4 * kMaxBlocks] bounding box of KF1's (x, y) + block partials
    float *X1f = nullptr, *X2f = nullptr;         // staging (caller order)
    int* d_perm = nullptr;
    double *P = nullptr, *Ptrial = nullptr, *P0 = nullptr, *Q = nullptr;
    float4* uv = nullptr;
    double2* dm = nullptr;
    float2* isg = nullptr;
    double* Je = nullptr;                 // per directed edge {u, m, g}
    // raw (caller order) device copies: the internal order is produced on the device (dsc_graph.cuh)
    float2 *r_uv1 = nullptr, *r_uv2 = nullptr; double *r_d1 = nullptr, *r_d2 = nullptr; float *r_isg1 = nullptr, *r_isg2 = nullptr;
    bool r_has_isg1 = false, r_has_isg2 = false;
    int *g_rp0 = nullptr, *g_col0 = nullptr, *g_inv = nullptr, *g_width = nullptr, *g_sums = nullptr;
    double* g_w0 = nullptr;
    unsigned long long *g_key0 = nullptr, *g_key1 = nullptr;
    void* g_tmp = nullptr;
    size_t g_ecap = 0, g_ncap = 0, g_tmpcap = 0;
    int* spmv_part = nullptr;                   // [spmv_units + 1] slice ranges = work units of cg_spmv_kernel
    int spmv_units = 0, spmv_cap = 0;
    int *ecol = nullptr, *sliceptr = nullptr;   // sliced ELL of the gather kernels (see dsc_set_graph)
    double* ewgt = nullptr;                     // ELL edge weights
    long long nblk = 0, blkcap = 0;
    int slcap = 0;
    double *b = nullptr, *D = nullptr, *U = nullptr, *Minv = nullptr;
    float* MinvS = nullptr;                      // float copy of the preconditioner blocks the fp64 per-iteration kernels stream (DSC_MINV_F64=1: double)
    bool minv_float = true;
    double* vec[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x r z w p s
    double* small = nullptr;              // 6 x 8 global vectors + Ginv(64)
    Globals *Gcur = nullptr, *Gtrial = nullptr;
    LinGlobal* lin = nullptr;
    CgControl* ctl = nullptr;
    int* errflag = nullptr;
    double* part = nullptr;               // partial-sum scratch
    double *gpart[2] = {nullptr, nullptr}, *dpart = nullptr, *bpart = nullptr;
    double* h_pinned = nullptr;           // pinned host scratch
    // pinned, persistent host staging of the graph / observation uploads (grown on demand)
    int* hs_sp = nullptr;
    size_t hc_sp = 0;
    // ---- Delaunay graph builder (device CSR of the last dsc_delaunay_build / dsc_set_graph_delaunay)
    int dl_n = 0; long long dl_E = 0, dl_ntri = 0; double dl_area = 0.0;
    int *dl_rowptr = nullptr, *dl_col = nullptr; double* dl_w = nullptr;
    float* dl_X = nullptr; size_t dl_xcap = 0;
    size_t dl_cap = 0;                               // points the three arrays are sized for (a planar mesh has < 6 n directed edges)
    long long dl_uncertified = 0;
    // grow-only device workspace of the graph builders (carved per call: no cudaMalloc / cudaFree on a warm context)
    unsigned char* ws = nullptr; size_t ws_cap = 0, ws_off = 0;
    // ---- kNN graph builder (device CSR of the last dsc_knn_build)
    int knn_n = 0; long long knn_E = 0;
    int *knn_rowptr = nullptr, *knn_col = nullptr;
    size_t knn_cap_n = 0, knn_cap_e = 0;
    dsc_pcg_params pcg{1e-10, 4000, 32};
    struct IterGraph { cudaGraphExec_t exec = nullptr; const double* P = nullptr; WeightsDev W{}; int precision = 0; } graphs[2];   // per state buffer
    bool use_graphs = true;
    int small_cluster = 0;                           // CTAs of the one-launch PCG of small problems (0 = not available)
    int small_max_rows = kSmallMaxRows;              // largest problem that takes that path (DSC_SMALL_MAX_ROWS overrides)
    int grid_max_rows = 0;                           // above small_max_rows and up to this: the same solve by a cooperative launch over all SMs (0 = off)
    int grid_blocks = 0;
    int solver = DSC_SOLVER_AUTO;                    // dense Cholesky for small problems, PCG above (dsc_set_solver)
    double *dnH = nullptr, *dnA = nullptr, *dn_rhs = nullptr, *dn_sol = nullptr, *dn_l11 = nullptr;
    int dn_cap = 0;
    int precision = DSC_PRECISION_F64;               // storage of the solver data (dsc_set_precision)
    // fp32 mode: float copies of what the PCG streams + its six vectors; the double buffers keep serving the residual
    // of the iterative refinement (X = vec[0], peeked solution = vec[1], operator output = vec[3])
    float *JeF = nullptr, *UF = nullptr, *MinvF = nullptr, *vecF[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int f32_cap = 0; long long f32_blkcap = 0;
    bool f32_fresh = false;                          // JeF's padding slots are all-zero for the current graph
    struct Refine { int pass_k = 0, total = 0, passes = 0; double gamma0 = 0.0, rho = 1.0; bool have_X = false; } rf;
    // ---- point-sharded pair (dsc_shard_init / dsc_shard_attach, dsc_shard.cuh): this context is ONE rank of a frame pair
    bool sharded = false, sh_attached = false;
    ShardDev sh{};                                   // device-visible descriptor (peer-mapped buffers of every rank)
    unsigned char* sh_arena = nullptr;               // local arena: Pbuf[2] | zbuf[2] | mailboxes | flags (exported by IPC)
    void* sh_peer[kMaxShards] = {};                  // peers' arenas as mapped here
    int sh_cap = 0;                                  // correspondences the arena was sized for
    int* sh_part = nullptr; int sh_units = 0, sh_partcap = 0;    // work units of cg_spmv over the rank's own slices
    unsigned char* sh_mask = nullptr;                // [cap] export mask
    double* sh_out = nullptr;                        // [kMboxDoubles] result of the last shard_allreduce_kernel
    int* sh_err = nullptr;
    long long sh_halo_rows = 0;                      // own rows that at least one peer holds as halo rows
    int early_levels = 0;                            // early rejection of clearly bad LM trials (off by default)
    double early_rtol[4] = {0, 0, 0, 0}, early_margin[4] = {0, 0, 0, 0};
};

namespace {

const char* status_str(int s) {
    switch (s) {
        case DSC_OK: return "ok";
        case DSC_ERR_INVALID_ARG: return "invalid argument";
        case DSC_ERR_CUDA: return "CUDA error";
        case DSC_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case DSC_ERR_STATE: return "call order violated";
        case DSC_ERR_NONFINITE: return "non-finite cost";
        case DSC_ERR_PCG_BREAKDOWN: return "PCG breakdown";
        case DSC_ERR_GRAPH: return "neighbour graph invalid (must be symmetric, in range, no self loops)";
        case DSC_ERR_ALLOC: return "allocation failed";
        case DSC_ERR_SHARD: return "point-sharded pair: a peer rank did not answer (or the sharding set-up is incomplete)";
    }
    return "unknown";
}

int fail(dsc_ctx* c, int code, const std::string& what) {
    if (c) c->err = std::string(status_str(code)) + ": " + what;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, DSC_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
cudaError_t dev_alloc(T*& p, size_t count) {
    if (p) { cudaFree(p); p = nullptr; }
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}
template <typename T>
void dev_free(T*& p) { if (p) { cudaFree(p); p = nullptr; } }
inline size_t ws_round(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
template <typename T>
cudaError_t pin_reserve(T*& p, size_t& cap, size_t count) {       // pinned host buffer with at least `count` elements
    if (count <= cap) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    size_t want = count + count / 8 + 64;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
}

// Host -> device copy of CALLER memory on the context's stream.  Pinned or registered memory (dsc_pin_host,
// cudaHostRegister, cudaMallocHost) is read by the copy engine where it lies: no staging, no host loop.  Pageable memory
// goes through the context's pinned bounce buffer in chunks; the parallel memcpy of chunk c + 1 overlaps the DMA of
// chunk c.  The caller's memory must stay valid until the stream has been synchronised (every upload entry point does).
constexpr size_t kBounceBytes = (size_t)16 << 20;
int h2d(dsc_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return DSC_OK;
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return DSC_OK;
    }
    for (int h = 0; h < 2; ++h)
        if (!ctx->bounce[h]) {
            CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->bounce[h]), kBounceBytes));
            CK(cudaEventCreateWithFlags(&ctx->bounce_ev[h], cudaEventDisableTiming));
        }
    const unsigned char* s8 = static_cast<const unsigned char*>(src);
    unsigned char* d8 = static_cast<unsigned char*>(dst);
    for (size_t off = 0; off < bytes; off += kBounceBytes) {
        const size_t len = std::min(kBounceBytes, bytes - off);
        const int h = (int)(ctx->bounce_next++ & 1u);
        CK(cudaEventSynchronize(ctx->bounce_ev[h]));               // the DMA that last read this half has finished
        unsigned char* bb = ctx->bounce[h];
        const long long pieces = (long long)((len + 65535) / 65536);
#pragma omp parallel for schedule(static)
        for (long long q = 0; q < pieces; ++q) {
            const size_t o = (size_t)q * 65536, l = std::min((size_t)65536, len - o);
            std::memcpy(bb + o, s8 + off + o, l);
        }
        CK(cudaMemcpyAsync(d8 + off, bb, len, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->bounce_ev[h], ctx->stream));
    }
    return DSC_OK;
}

int grid_threads(const dsc_ctx* c, long long n) {           // thread-per-item kernels
    long long nb = (n + kThreads - 1) / kThreads;
    long long cap = (long long)c->sms * 8;
    return (int)std::max(1LL, std::min(nb, cap));
}
int grid_spmv(const dsc_ctx* c, long long n) {              // persistent blocks over work units of >= 4 slices (see dsc_set_graph)
    long long nb = ((n + 31) / 32 + 3) / 4;
    long long cap = (long long)c->sms * 2;
    return (int)std::max(1LL, std::min(nb, cap));
}
int grid_tiles(const dsc_ctx* c, long long n, int per_sm) {  // one block per tile of kSortGroup rows, persistent
    long long nb = (n + kSortGroup - 1) / kSortGroup;
    return (int)std::max(1LL, std::min(nb, (long long)c->sms * per_sm));
}
void fill_pair(const dsc_pair* in, PairDev& o) {
    o.cam1.model = in->cam1.model; o.cam2.model = in->cam2.model;
    for (int k = 0; k < 8; ++k) { o.cam1.p[k] = in->cam1.params[k]; o.cam2.p[k] = in->cam2.params[k]; }
    for (int cam = 0; cam < 2; ++cam) {
        const float* T = cam == 0 ? in->T1w : in->T2w;
        PoseF& pf = cam == 0 ? o.T1f : o.T2f;
        double* Rd = cam == 0 ? o.R1 : o.R2;
        double* td = cam == 0 ? o.t1 : o.t2;
        double R64[9];
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc) { pf.R[r * 3 + cc] = T[r * 4 + cc]; R64[r * 3 + cc] = (double)T[r * 4 + cc]; }
            pf.t[r] = T[r * 4 + 3];
            td[r] = (double)T[r * 4 + 3];
        }
        // g2o::SE3Quat(kfPose.unit_quaternion().cast<double>(), ...): quaternion, normalised, back to a matrix
        double q[4];
        rot_to_quat(R64, q);
        quat_to_rot(q, Rd);
    }
}

// Work units of cg_spmv_kernel over the slices [s0, s1) (s0 a tile boundary): full rounds of 16-slice tiles, then the
// partial last round cut into one equal-work unit per block (a unit may be empty: the kernel skips it).
constexpr double kRowCost = 32.0 * 256.0 / 2432.0;             // the 32 rows of a slice in units of one ELL block
std::vector<int> spmv_units_of(const int* sp, int s0, int s1, int nbs) {
    std::vector<int> hpart;
    const int tsl = kSortGroup / 32;
    const int ntiles = (s1 - s0 + tsl - 1) / tsl;
    const int full = (ntiles / nbs) * nbs;                     // tiles in complete rounds
    for (int t = 0; t < full; ++t) hpart.push_back(s0 + t * tsl);
    int sl = std::min(s1, s0 + full * tsl);
    if (sl < s1) {
        const double total = (double)(sp[s1] - sp[sl]) + kRowCost * (s1 - sl);
        double acc = 0.0;
        for (int b = 0; b < nbs; ++b) {
            hpart.push_back(sl);
            const double target = total * (b + 1) / nbs;
            while (sl < s1) {
                const double c = (sp[sl + 1] - sp[sl]) + kRowCost;
                if (acc + 0.5 * c > target) break;
                acc += c; ++sl;
            }
        }
    }
    hpart.push_back(s1);
    return hpart;
}

void drop_graphs(dsc_ctx* c) {
    for (auto& g : c->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.P = nullptr; }
}

WeightsDev make_weights(const dsc_ctx* c, const dsc_weights* w) {
    WeightsDev o;
    o.rep = w->rep;
    o.arap_info = w->arap * (double)c->ntri * (double)c->ntri;
    double sd = (double)w->depth_sigma;
    o.depth_info = 1.0 / (sd * sd);
    o.inv_area = 1.0 / c->area;
    o.huber = (double)(float)std::sqrt(100.991);
    return o;
}

double host_sum(const double* p, int n, int stride = 1, int off = 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += p[(size_t)i * stride + off];
    return s;
}

}  // namespace

// ------------------------------------------------------------------ context
extern "C" int dsc_version(void) { return DSC_VERSION; }
extern "C" const char* dsc_status_string(int s) { return status_str(s); }
extern "C" const char* dsc_last_error(const dsc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int dsc_create(int device, dsc_ctx** out) {
    if (!out) return DSC_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DSC_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return DSC_ERR_INVALID_ARG;
    dsc_ctx* ctx = new dsc_ctx();
    ctx->device = device;
    auto bail = [&](int code) { dsc_destroy(ctx); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(DSC_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(DSC_ERR_CUDA);
    ctx->sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(DSC_ERR_CUDA);
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->evA); cudaEventCreate(&ctx->evB);
    for (auto& e : ctx->evo) if (cudaEventCreate(&e) != cudaSuccess) return bail(DSC_ERR_CUDA);
    if (cudaMalloc(&ctx->Gcur, sizeof(Globals)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->Gtrial, sizeof(Globals)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->lin, sizeof(LinGlobal)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->ctl, sizeof(CgControl)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->errflag, sizeof(int)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->small, sizeof(double) * (6 * 8 + 64 + 16)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->part, sizeof(double) * kMaxBlocks * kLinPart) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->gpart[0], sizeof(double) * kMaxBlocks) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->gpart[1], sizeof(double) * kMaxBlocks) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->dpart, sizeof(double) * kMaxBlocks) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->bpart, sizeof(double) * kMaxBlocks * 8) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMallocHost(&ctx->h_pinned, sizeof(double) * kMaxBlocks * kLinPart) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    if (cudaMalloc(&ctx->d_box, sizeof(float) * (4 + 4 * kMaxBlocks)) != cudaSuccess) return bail(DSC_ERR_ALLOC);
    cudaMemset(ctx->errflag, 0, sizeof(int));
    ctx->use_graphs = std::getenv("DSC_NO_GRAPHS") == nullptr;
    ctx->minv_float = std::getenv("DSC_MINV_F64") == nullptr;
    if (std::getenv("DSC_NO_CLUSTER_PCG") == nullptr) {
        // the one-launch PCG of small problems needs a cluster of 16 (non-portable) or 8 CTAs of 256 threads
        ctx->small_cluster = cudaFuncSetAttribute(pcg_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess ? kSmallCluster : 8;
        cudaGetLastError();
        // tuning knobs: cluster size (1..16) and the size limit of the one-launch path
        if (const char* cs = std::getenv("DSC_CLUSTER")) ctx->small_cluster = std::max(1, std::min(kSmallCluster, std::atoi(cs)));
        if (const char* mr = std::getenv("DSC_SMALL_MAX_ROWS")) ctx->small_max_rows = std::max(0, std::atoi(mr));
        // grid-wide variant: needs cooperative launch; one CTA per SM (254 registers x 256 threads)
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_grid_kernel, kThreads, 0) == cudaSuccess && per_sm > 0) {
            ctx->grid_blocks = ctx->sms * per_sm;
            ctx->grid_max_rows = kGridMaxRows;
            if (const char* gr = std::getenv("DSC_GRID_MAX_ROWS")) ctx->grid_max_rows = std::max(0, std::atoi(gr));
        }
        cudaGetLastError();
    }
    if (cudaFuncSetAttribute(rotations_ell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinBytes) != cudaSuccess ||
        cudaFuncSetAttribute(cost_ell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinBytes) != cudaSuccess ||
        cudaFuncSetAttribute(cg_spmv_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SpmvCfg<double>::kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(cg_spmv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SpmvCfg<float>::kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(cg_spmv_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SpmvCfg<double>::kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(linearize_ell_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinBytes) != cudaSuccess ||
        cudaFuncSetAttribute(dense_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * kDenseMaxM)) != cudaSuccess ||
        cudaFuncSetAttribute(linearize_ell_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinBytes) != cudaSuccess) return bail(DSC_ERR_CUDA);
    *out = ctx;
    return DSC_OK;
}

extern "C" void dsc_destroy(dsc_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    drop_graphs(ctx);
    if (ctx->sharded) {                                // P, Ptrial and z live in the exported arena; peers' arenas are unmapped
        ctx->P = nullptr; ctx->Ptrial = nullptr; ctx->vec[2] = nullptr;
        for (int q = 0; q < kMaxShards; ++q) if (ctx->sh_peer[q]) cudaIpcCloseMemHandle(ctx->sh_peer[q]);
        if (ctx->sh_arena) cudaFree(ctx->sh_arena);
        if (ctx->sh.sent) cudaFree(ctx->sh.sent);
        dev_free(ctx->sh_part); dev_free(ctx->sh_mask); dev_free(ctx->sh_out); dev_free(ctx->sh_err);
    }
    dev_free(ctx->t_uv1); dev_free(ctx->t_uv2); dev_free(ctx->t_d1); dev_free(ctx->t_d2);
    dev_free(ctx->t_X1); dev_free(ctx->t_X2); dev_free(ctx->t_cos); dev_free(ctx->t_valid);
    dev_free(ctx->X1f); dev_free(ctx->X2f); dev_free(ctx->d_perm);
    dev_free(ctx->dnH); dev_free(ctx->dnA); dev_free(ctx->dn_rhs); dev_free(ctx->dn_sol); dev_free(ctx->dn_l11);
    dev_free(ctx->r_uv1); dev_free(ctx->r_uv2); dev_free(ctx->r_d1); dev_free(ctx->r_d2); dev_free(ctx->r_isg1); dev_free(ctx->r_isg2);
    dev_free(ctx->d_box); dev_free(ctx->g_rp0); dev_free(ctx->g_col0); dev_free(ctx->g_inv);
    dev_free(ctx->g_width); dev_free(ctx->g_sums); dev_free(ctx->g_w0); dev_free(ctx->g_key0); dev_free(ctx->g_key1);
    if (ctx->g_tmp) { cudaFree(ctx->g_tmp); ctx->g_tmp = nullptr; }
    dev_free(ctx->P); dev_free(ctx->Ptrial); dev_free(ctx->P0); dev_free(ctx->Q);
    dev_free(ctx->uv); dev_free(ctx->dm); dev_free(ctx->isg);
    dev_free(ctx->Je); dev_free(ctx->ecol); dev_free(ctx->ewgt); dev_free(ctx->sliceptr); dev_free(ctx->spmv_part);
    dev_free(ctx->b); dev_free(ctx->D); dev_free(ctx->U); dev_free(ctx->Minv); dev_free(ctx->MinvS);
    for (auto& v : ctx->vec) dev_free(v);
    for (auto& v : ctx->vecF) dev_free(v);
    dev_free(ctx->JeF); dev_free(ctx->UF); dev_free(ctx->MinvF);
    dev_free(ctx->knn_rowptr); dev_free(ctx->knn_col);
    dev_free(ctx->dl_rowptr); dev_free(ctx->dl_col); dev_free(ctx->dl_w); dev_free(ctx->ws); dev_free(ctx->dl_X);
    dev_free(ctx->small); dev_free(ctx->Gcur); dev_free(ctx->Gtrial); dev_free(ctx->lin); dev_free(ctx->ctl);
    dev_free(ctx->errflag); dev_free(ctx->part); dev_free(ctx->gpart[0]); dev_free(ctx->gpart[1]);
    dev_free(ctx->dpart); dev_free(ctx->bpart);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (void* q : {(void*)ctx->hs_sp, (void*)ctx->bounce[0], (void*)ctx->bounce[1]}) if (q) cudaFreeHost(q);
    for (auto& e : ctx->bounce_ev) if (e) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->evA) cudaEventDestroy(ctx->evA);
    if (ctx->evB) cudaEventDestroy(ctx->evB);
    for (auto& e : ctx->evo) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Page-lock caller memory (cudaHostRegister) so that the upload / download entry points move it by DMA without a host
// copy.  Worth it for buffers that live across calls (a SLAM front end's key-point / map-point arrays, bench.py's
// inputs); registering costs about as much as one copy of the buffer.  ptr / bytes as for cudaHostRegister.
extern "C" int dsc_pin_host(const void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return DSC_ERR_INVALID_ARG;
    cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return DSC_OK; }
    if (e != cudaSuccess) { cudaGetLastError(); return DSC_ERR_CUDA; }
    return DSC_OK;
}
extern "C" int dsc_unpin_host(const void* ptr) {
    if (!ptr) return DSC_ERR_INVALID_ARG;
    cudaError_t e = cudaHostUnregister(const_cast<void*>(ptr));
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorHostMemoryNotRegistered ? DSC_OK : DSC_ERR_CUDA; }
    return DSC_OK;
}

extern "C" int dsc_synchronize(dsc_ctx* ctx) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}
extern "C" int dsc_timer_start(dsc_ctx* ctx) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return DSC_OK;
}
extern "C" int dsc_timer_stop(dsc_ctx* ctx, double* ms) {
    if (!ctx || !ms) return DSC_ERR_INVALID_ARG;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = (double)f;
    return DSC_OK;
}
extern "C" int dsc_launch_count(const dsc_ctx* ctx, long long* count) {
    if (!ctx || !count) return DSC_ERR_INVALID_ARG;
    *count = ctx->launches;
    return DSC_OK;
}

// ------------------------------------------------------------------ K1 triangulation
extern "C" int dsc_tri_upload(dsc_ctx* ctx, const dsc_pair* pair, int n, const float* uv1, const float* uv2,
                              const float* depth1, const float* depth2) {
    if (!ctx || !pair || n < 0 || (n > 0 && (!uv1 || !uv2))) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_tri_upload");
    CK(cudaSetDevice(ctx->device));
    if (n > ctx->tcap) {
        CK(dev_alloc(ctx->t_uv1, (size_t)n)); CK(dev_alloc(ctx->t_uv2, (size_t)n));
        CK(dev_alloc(ctx->t_d1, (size_t)n)); CK(dev_alloc(ctx->t_d2, (size_t)n));
        CK(dev_alloc(ctx->t_X1, (size_t)3 * n)); CK(dev_alloc(ctx->t_X2, (size_t)3 * n));
        CK(dev_alloc(ctx->t_cos, (size_t)n)); CK(dev_alloc(ctx->t_valid, (size_t)n));
        ctx->tcap = n;
    }
    ctx->tn = n;
    fill_pair(pair, ctx->tpair);
    ctx->t_has_depth = depth1 && depth2;
    ctx->t_done = false;
    if (n == 0) return DSC_OK;
    CK(cudaMemcpyAsync(ctx->t_uv1, uv1, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->t_uv2, uv2, sizeof(float2) * n, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->t_has_depth) {
        CK(cudaMemcpyAsync(ctx->t_d1, depth1, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->t_d2, depth2, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    }
    return DSC_OK;
}

extern "C" int dsc_tri_run(dsc_ctx* ctx, const dsc_tri_params* prm) {
    if (!ctx || !prm) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_tri_run");
    if (prm->method == DSC_TRI_DEPTH && !ctx->t_has_depth) return fail(ctx, DSC_ERR_INVALID_ARG, "DepthMeasurement needs depths");
    if (prm->method < 0 || prm->method > 3 || prm->location < 0 || prm->location > 2 || prm->gate < 0 || prm->gate > 2)
        return fail(ctx, DSC_ERR_INVALID_ARG, "triangulation parameters");
    CK(cudaSetDevice(ctx->device));
    ctx->t_done = true;
    if (ctx->tn == 0) return DSC_OK;
    TriParams tp{prm->method, prm->location, prm->gate, prm->min_cos, prm->depth_limit, prm->check_reproj};
    triangulate_kernel<<<grid_threads(ctx, ctx->tn), kThreads, 0, ctx->stream>>>(
        ctx->tn, ctx->t_uv1, ctx->t_uv2, ctx->t_d1, ctx->t_d2, ctx->tpair, tp, ctx->t_X1, ctx->t_X2, ctx->t_valid, ctx->t_cos);
    ctx->launches++;
    CK(cudaGetLastError());
    return DSC_OK;
}

extern "C" int dsc_tri_download(dsc_ctx* ctx, float* X1, float* X2, uint8_t* valid, float* cos_parallax, int* n_valid) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (!ctx->t_done) return fail(ctx, DSC_ERR_STATE, "dsc_tri_download before dsc_tri_run");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->tn;
    std::vector<uint8_t> hv;
    if (n > 0) {
        if (X1) CK(cudaMemcpyAsync(X1, ctx->t_X1, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (X2) CK(cudaMemcpyAsync(X2, ctx->t_X2, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (cos_parallax) CK(cudaMemcpyAsync(cos_parallax, ctx->t_cos, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        uint8_t* dst = valid;
        if (!dst && n_valid) { hv.resize(n); dst = hv.data(); }
        if (dst) CK(cudaMemcpyAsync(dst, ctx->t_valid, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (n_valid) { int c = 0; for (int i = 0; i < n; ++i) c += dst[i] ? 1 : 0; *n_valid = c; }
    } else if (n_valid) *n_valid = 0;
    return DSC_OK;
}

extern "C" int dsc_triangulate(dsc_ctx* ctx, const dsc_pair* pair, const dsc_tri_params* prm, int n,
                               const float* uv1, const float* uv2, const float* depth1, const float* depth2,
                               float* X1, float* X2, uint8_t* valid, float* cos_parallax, int* n_valid) {
    int s = dsc_tri_upload(ctx, pair, n, uv1, uv2, depth1, depth2);
    if (s) return s;
    s = dsc_tri_run(ctx, prm);
    if (s) return s;
    return dsc_tri_download(ctx, X1, X2, valid, cos_parallax, n_valid);
}

extern "C" int dsc_triangulate_rays(dsc_ctx* ctx, const float* T1w, const float* T2w, int method, int location, int n,
                                    const float* xn1, const float* xn2, float* X1, float* X2) {
    if (!ctx || !T1w || !T2w || n < 0 || (n > 0 && (!xn1 || !xn2 || !X1 || !X2))) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_triangulate_rays");
    if (method < 0 || method > 3 || location < 0 || location > 2) return fail(ctx, DSC_ERR_INVALID_ARG, "triangulation parameters");
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return DSC_OK;
    dsc_pair pr{};
    std::memcpy(pr.T1w, T1w, sizeof(pr.T1w));
    std::memcpy(pr.T2w, T2w, sizeof(pr.T2w));
    PairDev pd{};
    fill_pair(&pr, pd);
    if (n > ctx->tcap) {
        CK(dev_alloc(ctx->t_uv1, (size_t)n)); CK(dev_alloc(ctx->t_uv2, (size_t)n));
        CK(dev_alloc(ctx->t_d1, (size_t)n)); CK(dev_alloc(ctx->t_d2, (size_t)n));
        CK(dev_alloc(ctx->t_X1, (size_t)3 * n)); CK(dev_alloc(ctx->t_X2, (size_t)3 * n));
        CK(dev_alloc(ctx->t_cos, (size_t)n)); CK(dev_alloc(ctx->t_valid, (size_t)n));
        ctx->tcap = n;
    }
    ctx->tn = 0; ctx->t_done = false;                      // the output buffers of the pixel stage are reused: invalidate it
    float *d_in1 = nullptr, *d_in2 = nullptr;              // stream-ordered staging of the rays
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_in1), sizeof(float) * 3 * (size_t)n, ctx->stream));
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_in2), sizeof(float) * 3 * (size_t)n, ctx->stream));
    CK(cudaMemcpyAsync(d_in1, xn1, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_in2, xn2, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    triangulate_rays_kernel<<<grid_threads(ctx, n), kThreads, 0, ctx->stream>>>(n, d_in1, d_in2, pd, method, location, ctx->t_X1, ctx->t_X2);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(X1, ctx->t_X1, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(X2, ctx->t_X2, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaFreeAsync(d_in1, ctx->stream));
    CK(cudaFreeAsync(d_in2, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

extern "C" int dsc_depth_scale_init(dsc_ctx* ctx, int which, double* scale) {
    if (!ctx || !scale || (which != 1 && which != 2)) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_depth_scale_init");
    if (!ctx->t_done || !ctx->t_has_depth) return fail(ctx, DSC_ERR_STATE, "needs a triangulation run with depths");
    CK(cudaSetDevice(ctx->device));
    int nb = grid_threads(ctx, ctx->tn);
    if (ctx->tn == 0) { *scale = std::numeric_limits<double>::quiet_NaN(); return DSC_OK; }
    depth_scale_kernel<<<nb, kThreads, 0, ctx->stream>>>(ctx->tn, which == 1 ? ctx->t_X1 : ctx->t_X2,
                                                        which == 1 ? ctx->t_d1 : ctx->t_d2, ctx->t_valid,
                                                        which == 1 ? ctx->tpair.T1f : ctx->tpair.T2f, ctx->part);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double s = host_sum(ctx->h_pinned, nb, 2, 0), c = host_sum(ctx->h_pinned, nb, 2, 1);
    *scale = s / (double)(float)c;
    return DSC_OK;
}

// ------------------------------------------------------------------ refinement problem
// device state in the internal order from the raw (caller order) device copies; perm = ctx->d_perm or identity
static int build_state(dsc_ctx* ctx) {
    int n = ctx->n;
    if (n == 0) return DSC_OK;
    const int* dp = ctx->perm.empty() ? nullptr : ctx->d_perm;
    permute_obs_kernel<<<grid_threads(ctx, n), kThreads, 0, ctx->stream>>>(n, dp, ctx->r_uv1, ctx->r_uv2, ctx->r_d1, ctx->r_d2,
                                                                          ctx->r_has_isg1 ? ctx->r_isg1 : nullptr, ctx->r_has_isg2 ? ctx->r_isg2 : nullptr,
                                                                          ctx->uv, ctx->dm, ctx->isg);
    init_state_kernel<<<grid_threads(ctx, n), kThreads, 0, ctx->stream>>>(n, ctx->X1f, ctx->X2f, dp, ctx->P0);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->P, ctx->P0, sizeof(double) * 8 * n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->Gcur, &ctx->g0, sizeof(Globals), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

// caller arrays -> raw device copies (caller order, copied as they are: h2d) -> internal order (device kernels)
static int upload_state(dsc_ctx* ctx, const float* X1, const float* X2, const float* uv1, const float* uv2, const double* d1,
                        const double* d2, const float* isg1, const float* isg2) {
    const size_t n = (size_t)ctx->n;
    if (n == 0) return DSC_OK;
    int rc;
    if ((rc = h2d(ctx, ctx->r_uv1, uv1, sizeof(float2) * n))) return rc;
    if ((rc = h2d(ctx, ctx->r_uv2, uv2, sizeof(float2) * n))) return rc;
    if ((rc = h2d(ctx, ctx->r_d1, d1, sizeof(double) * n))) return rc;
    if ((rc = h2d(ctx, ctx->r_d2, d2, sizeof(double) * n))) return rc;
    ctx->r_has_isg1 = isg1 != nullptr; ctx->r_has_isg2 = isg2 != nullptr;
    if (isg1 && (rc = h2d(ctx, ctx->r_isg1, isg1, sizeof(float) * n))) return rc;
    if (isg2 && (rc = h2d(ctx, ctx->r_isg2, isg2, sizeof(float) * n))) return rc;
    if ((rc = h2d(ctx, ctx->X1f, X1, sizeof(float) * 3 * n))) return rc;
    if ((rc = h2d(ctx, ctx->X2f, X2, sizeof(float) * 3 * n))) return rc;
    return build_state(ctx);
}

extern "C" int dsc_problem_upload(dsc_ctx* ctx, const dsc_pair* pair, int n,
                                  const float* X1, const float* X2, const float* uv1, const float* uv2,
                                  const double* depth1, const double* depth2,
                                  const float* inv_sigma2_1, const float* inv_sigma2_2,
                                  double scale1, double scale2, const double* Tg7) {
    if (!ctx || !pair || n < 0 || (n > 0 && (!X1 || !X2 || !uv1 || !uv2 || !depth1 || !depth2)))
        return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_problem_upload");
    CK(cudaSetDevice(ctx->device));
    if (ctx->sharded && n > ctx->sh_cap) return fail(ctx, DSC_ERR_INVALID_ARG, "more correspondences than dsc_shard_init was sized for");
    if (n > ctx->cap) {
        size_t N = (size_t)n;
        CK(dev_alloc(ctx->X1f, 3 * N)); CK(dev_alloc(ctx->X2f, 3 * N)); CK(dev_alloc(ctx->d_perm, N));
        if (!ctx->sharded) { CK(dev_alloc(ctx->P, 8 * N)); CK(dev_alloc(ctx->Ptrial, 8 * N)); }      // (sharded: in the exported arena)
        CK(dev_alloc(ctx->P0, 8 * N)); CK(dev_alloc(ctx->Q, 4 * N));
        CK(dev_alloc(ctx->uv, N)); CK(dev_alloc(ctx->dm, N)); CK(dev_alloc(ctx->isg, N));
        CK(dev_alloc(ctx->r_uv1, N)); CK(dev_alloc(ctx->r_uv2, N)); CK(dev_alloc(ctx->r_d1, N)); CK(dev_alloc(ctx->r_d2, N));
        CK(dev_alloc(ctx->r_isg1, N)); CK(dev_alloc(ctx->r_isg2, N));
        CK(dev_alloc(ctx->b, 6 * N)); CK(dev_alloc(ctx->D, 21 * 32 * ((N + 31) / 32))); CK(dev_alloc(ctx->U, (size_t)kURec * 32 * ((N + 31) / 32))); CK(dev_alloc(ctx->Minv, 21 * 32 * ((N + 31) / 32))); CK(dev_alloc(ctx->MinvS, 21 * 32 * ((N + 31) / 32)));
        for (int k = 0; k < 6; ++k) if (!(ctx->sharded && k == 2)) CK(dev_alloc(ctx->vec[k], 6 * N));
        ctx->cap = n;
    }
    ctx->n = n;
    fill_pair(pair, ctx->pair);
    Globals g{};
    if (Tg7) {
        double nq = std::sqrt(Tg7[0] * Tg7[0] + Tg7[1] * Tg7[1] + Tg7[2] * Tg7[2] + Tg7[3] * Tg7[3]);
        if (!(nq > 0.0)) return fail(ctx, DSC_ERR_INVALID_ARG, "T_global quaternion has zero norm");
        double s = (Tg7[3] < 0 ? -1.0 : 1.0) / nq;
        for (int k = 0; k < 4; ++k) g.Tg[k] = Tg7[k] * s;
        for (int k = 4; k < 7; ++k) g.Tg[k] = Tg7[k];
    } else { g.Tg[3] = 1.0; }
    g.s1 = scale1; g.s2 = scale2;
    quat_to_rot(g.Tg, g.Rg);
    ctx->g0 = g;
    ctx->perm.clear();
    ctx->have_problem = true; ctx->have_graph = false; ctx->have_rot = false;
    drop_graphs(ctx);
    return upload_state(ctx, X1, X2, uv1, uv2, depth1, depth2, inv_sigma2_1, inv_sigma2_2);
}

// ------------------------------------------------------------------ point-sharded pair: set-up (dsc_shard.cuh)
// Row partition: contiguous tile ranges of (nearly) equal work -- ELL blocks + rows, the cost model of the operator.
extern "C" int dsc_shard_partition(const int32_t* sliceptr, int nslices, int world, int32_t* row_begin) {
    if (!sliceptr || !row_begin || nslices < 0 || world < 1 || world > kMaxShards) return DSC_ERR_INVALID_ARG;
    const int tsl = kSortGroup / 32;
    const int ntiles = (nslices + tsl - 1) / tsl;
    auto cost_to = [&](int tile) {                             // work of tiles [0, tile)
        const int sl = std::min(nslices, tile * tsl);
        return (double)sliceptr[sl] + kRowCost * sl;
    };
    const double total = cost_to(ntiles);
    row_begin[0] = 0;
    int t = 0;
    for (int r = 1; r < world; ++r) {
        const double target = total * r / world;
        while (t < ntiles && cost_to(t + 1) - 0.5 * (cost_to(t + 1) - cost_to(t)) <= target) ++t;
        row_begin[r] = std::min(t * kSortGroup, nslices * 32);
    }
    row_begin[world] = nslices * 32;                           // (clamped to n by the caller)
    for (int r = 1; r <= world; ++r) row_begin[r] = std::max(row_begin[r], row_begin[r - 1]);
    return DSC_OK;
}

static size_t shard_align(size_t x) { return (x + 255) & ~(size_t)255; }
struct ShardLayout { size_t P[2], z[2], mbox, flags, total; };
static ShardLayout shard_layout(int cap, int world) {
    ShardLayout L{};
    size_t off = 0;
    const size_t N = (size_t)cap;
    for (int k = 0; k < 2; ++k) { L.P[k] = off; off = shard_align(off + sizeof(double) * 8 * N); }
    for (int k = 0; k < 2; ++k) { L.z[k] = off; off = shard_align(off + sizeof(double) * 6 * N); }
    L.mbox = off; off = shard_align(off + sizeof(double) * SF_COUNT * 2 * world * kMboxDoubles);
    L.flags = off; off = shard_align(off + sizeof(unsigned long long) * SF_COUNT * world);
    L.total = off;
    return L;
}

extern "C" int dsc_shard_init(dsc_ctx* ctx, int rank, int world, int max_points, void* handle_out) {
    if (!ctx || !handle_out || world < 1 || world > kMaxShards || rank < 0 || rank >= world || max_points < 1)
        return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_shard_init");
    if (ctx->sharded || ctx->have_problem) return fail(ctx, DSC_ERR_STATE, "dsc_shard_init: call it once, before dsc_problem_upload");
    static_assert(sizeof(cudaIpcMemHandle_t) == DSC_SHARD_HANDLE_BYTES, "handle size");
    CK(cudaSetDevice(ctx->device));
    const ShardLayout L = shard_layout(max_points, world);
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->sh_arena), L.total));
    CK(cudaMemset(ctx->sh_arena, 0, L.total));
    unsigned long long* sent = nullptr; unsigned int* ticket = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&sent), sizeof(unsigned long long) * SF_COUNT + sizeof(unsigned int) * SF_COUNT));
    CK(cudaMemset(sent, 0, sizeof(unsigned long long) * SF_COUNT + sizeof(unsigned int) * SF_COUNT));
    ticket = reinterpret_cast<unsigned int*>(sent + SF_COUNT);
    CK(dev_alloc(ctx->sh_mask, (size_t)max_points)); CK(dev_alloc(ctx->sh_out, (size_t)kMboxDoubles)); CK(dev_alloc(ctx->sh_err, (size_t)1));
    CK(cudaMemset(ctx->sh_mask, 0, (size_t)max_points));
    CK(cudaMemset(ctx->sh_err, 0, sizeof(int)));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->sh_arena));
    std::memcpy(handle_out, &h, sizeof(h));
    dev_free(ctx->P); dev_free(ctx->Ptrial); dev_free(ctx->vec[2]);
    ctx->P = reinterpret_cast<double*>(ctx->sh_arena + L.P[0]);
    ctx->Ptrial = reinterpret_cast<double*>(ctx->sh_arena + L.P[1]);
    ctx->vec[2] = reinterpret_cast<double*>(ctx->sh_arena + L.z[0]);
    ctx->cap = 0;                                              // every other buffer is (re)allocated by the next upload
    ShardDev& S = ctx->sh;
    S = ShardDev{};
    S.rank = rank; S.world = world;
    S.sent = sent; S.ticket = ticket; S.exportmask = ctx->sh_mask; S.error = ctx->sh_err;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    const char* lim = std::getenv("DSC_SHARD_TIMEOUT_S");
    S.spin_limit = (long long)((lim ? std::atof(lim) : 5.0) * 1e3 * (double)prop.clockRate);       // clockRate is in kHz
    ctx->sh_cap = max_points;
    ctx->sharded = true; ctx->sh_attached = false;
    return DSC_OK;
}

extern "C" int dsc_shard_attach(dsc_ctx* ctx, const void* handles) {
    if (!ctx || !handles) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_shard_attach");
    if (!ctx->sharded || ctx->sh_attached) return fail(ctx, DSC_ERR_STATE, "dsc_shard_attach: after dsc_shard_init, once");
    CK(cudaSetDevice(ctx->device));
    ShardDev& S = ctx->sh;
    const ShardLayout L = shard_layout(ctx->sh_cap, S.world);
    for (int q = 0; q < S.world; ++q) {
        unsigned char* base = ctx->sh_arena;
        if (q != S.rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)q * DSC_SHARD_HANDLE_BYTES, sizeof(h));
            void* mapped = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(ctx, DSC_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(q) + ") -> " + cudaGetErrorString(e));
            ctx->sh_peer[q] = mapped;
            base = static_cast<unsigned char*>(mapped);
        }
        for (int k = 0; k < 2; ++k) {
            S.Pbuf[k][q] = reinterpret_cast<double*>(base + L.P[k]);
            S.zbuf[k][q] = reinterpret_cast<double*>(base + L.z[k]);
        }
        S.mbox[q] = reinterpret_cast<double*>(base + L.mbox);
        S.flags[q] = reinterpret_cast<unsigned long long*>(base + L.flags);
    }
    ctx->sh_attached = true;
    return DSC_OK;
}

extern "C" int dsc_shard_info(const dsc_ctx* ctx, int* rank, int* world, int* row_begin, int* row_end, long long* halo_rows) {
    if (!ctx || !ctx->sharded) return DSC_ERR_INVALID_ARG;
    if (rank) *rank = ctx->sh.rank;
    if (world) *world = ctx->sh.world;
    if (row_begin) *row_begin = ctx->sh.row_begin[ctx->sh.rank];
    if (row_end) *row_end = ctx->sh.row_begin[ctx->sh.rank + 1];
    if (halo_rows) *halo_rows = ctx->sh_halo_rows;
    return DSC_OK;
}

// after the sliced ELL exists: the row partition, the rank's work units of the operator, the export mask
static int shard_setup_graph(dsc_ctx* ctx, const int* sp, int nslices) {
    if (!ctx->sh_attached) return fail(ctx, DSC_ERR_SHARD, "dsc_shard_attach has not been called");
    ShardDev& S = ctx->sh;
    int rb[kMaxShards + 1];
    dsc_shard_partition(sp, nslices, S.world, rb);
    for (int r = 0; r <= S.world; ++r) S.row_begin[r] = std::min(rb[r], ctx->n);
    const int r0 = S.row_begin[S.rank], r1 = S.row_begin[S.rank + 1];
    const int s0 = r0 / 32, s1 = (r1 + 31) / 32;
    std::vector<int> hpart = spmv_units_of(sp, s0, s1, grid_spmv(ctx, std::max(1, r1 - r0)));
    ctx->sh_units = (int)hpart.size() - 1;
    if ((int)hpart.size() > ctx->sh_partcap) { CK(dev_alloc(ctx->sh_part, hpart.size())); ctx->sh_partcap = (int)hpart.size(); }
    CK(cudaMemcpyAsync(ctx->sh_part, hpart.data(), sizeof(int) * hpart.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->sh_mask, 0, (size_t)ctx->n, ctx->stream));
    if (r1 > r0) {
        shard_exportmask_kernel<<<grid_threads(ctx, r1 - r0), 256, 0, ctx->stream>>>(S, ctx->n, ctx->sliceptr, ctx->ecol, ctx->sh_mask);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    std::vector<unsigned char> hm((size_t)ctx->n);
    if (ctx->n) CK(cudaMemcpyAsync(hm.data(), ctx->sh_mask, (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    long long halo = 0;
    for (unsigned char m : hm) halo += m ? 1 : 0;
    ctx->sh_halo_rows = halo;
    // the trial buffer of every rank starts as a copy of the state: rows nobody pushes (not in any halo) stay valid
    if (ctx->n) CK(cudaMemcpyAsync(ctx->Ptrial, ctx->P, sizeof(double) * 8 * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

static int reserve_graph(dsc_ctx* ctx, int n, long long E) {
    if ((size_t)n + 1 > ctx->g_ncap) {
        size_t N = (size_t)n + 1, NS = ((size_t)n + 31) / 32 + 2;
        CK(dev_alloc(ctx->g_rp0, N)); CK(dev_alloc(ctx->g_inv, N)); CK(dev_alloc(ctx->g_key0, N)); CK(dev_alloc(ctx->g_key1, N));
        CK(dev_alloc(ctx->g_width, NS)); CK(dev_alloc(ctx->g_sums, NS / kScanBlock + 2));
        ctx->g_ncap = N;
    }
    if ((size_t)E > ctx->g_ecap) { CK(dev_alloc(ctx->g_col0, (size_t)E)); CK(dev_alloc(ctx->g_w0, (size_t)E)); ctx->g_ecap = (size_t)E; }
    return DSC_OK;
}

// the rest of dsc_set_graph once the raw CSR (caller numbering) is on the device in g_rp0 / g_col0 / g_w0
static int set_graph_tail(dsc_ctx* ctx, int n, long long E, double area, long long n_triangles, int reorder, bool validate, bool timing) {
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dsc_set_graph] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };

    // symmetric, in range, no self loops, no duplicates, symmetric weights (the reference's mesh adjacency always is)
    if (validate && n > 0) {
        CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
        validate_graph_kernel<<<grid_threads(ctx, n), kThreads, 0, ctx->stream>>>(n, ctx->g_rp0, ctx->g_col0, ctx->g_w0, ctx->errflag);
        ctx->launches++;
        int* hp = reinterpret_cast<int*>(ctx->h_pinned + 5 * kMaxBlocks);
        CK(cudaMemcpyAsync(hp, ctx->errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        static const char* const what[] = {"", "graph is not symmetric", "edge weights are not symmetric", "duplicate edge",
                                           "column index out of range or self loop"};
        if (*hp) return fail(ctx, DSC_ERR_GRAPH, what[std::min(*hp, 4)]);
        lap("validate");
    }
    // ---- internal numbering on the device: Morton order of KF1's world (x, y) -- the plane the reference triangulates
    // in -- then, for the sliced ELL, a stable sort by degree (descending) inside every group of kSortGroup rows
    const int nbv = grid_threads(ctx, std::max(n, 1));
    const bool permuted = reorder && n > 1;
    if (permuted) {
        bbox_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->X1f, ctx->d_box + 4);
        bbox_final_kernel<<<1, 32, 0, ctx->stream>>>(nbv, ctx->d_box + 4, ctx->d_box);
        morton_key_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->X1f, ctx->d_box, ctx->g_key0);
        ctx->launches += 2;
        size_t need = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, need, ctx->g_key0, ctx->g_key1, n, 0, 64, ctx->stream));
        if (need > ctx->g_tmpcap) {
            if (ctx->g_tmp) cudaFree(ctx->g_tmp);
            ctx->g_tmp = nullptr; ctx->g_tmpcap = 0;
            CK(cudaMalloc(&ctx->g_tmp, need));
            ctx->g_tmpcap = need;
        }
        CK(cub::DeviceRadixSort::SortKeys(ctx->g_tmp, need, ctx->g_key0, ctx->g_key1, n, 0, 64, ctx->stream));
        perm_from_key_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->g_key1, ctx->d_perm);
        degree_sort_kernel<<<(n + kSortGroup - 1) / kSortGroup, kSortGroup, 0, ctx->stream>>>(n, ctx->g_rp0, ctx->d_perm);
        inverse_perm_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->d_perm, ctx->g_inv);
        ctx->launches += 4;
    }
    const int* dperm = permuted ? ctx->d_perm : nullptr;
    const int* dinv = permuted ? ctx->g_inv : nullptr;
    // ---- sliced ELL of the PCG operator and the gather kernels: slice = 32 consecutive rows (one warp), width = longest
    // row of the slice; column k of slice s is "block" sliceptr[s] + k: 32 column indices, 32 weights and 9 x 32 Jacobian
    // doubles (Je).  Padding entries point at the row itself and keep an all-zero Jacobian record, so they add exactly 0.
    int nslices = (n + 31) / 32;
    if (nslices + 1 > ctx->slcap) { CK(dev_alloc(ctx->sliceptr, (size_t)nslices + 1)); ctx->slcap = nslices + 1; }
    {
        const int m = nslices + 1, nb = (m + kScanBlock - 1) / kScanBlock;
        slice_width_kernel<<<std::max(1, std::min(nbv, (nslices + 8) / 8)), kThreads, 0, ctx->stream>>>(n, nslices, dperm, ctx->g_rp0, ctx->g_width);
        scan_block_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(m, ctx->g_width, ctx->sliceptr, ctx->g_sums);
        scan_sums_kernel<<<1, kScanBlock, 0, ctx->stream>>>(nb, ctx->g_sums);
        scan_add_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(m, ctx->sliceptr, ctx->g_sums);
        ctx->launches += 4;
    }
    CK(cudaGetLastError());
    CK(pin_reserve(ctx->hs_sp, ctx->hc_sp, (size_t)nslices + 1));
    int* sp = ctx->hs_sp;
    CK(cudaMemcpyAsync(sp, ctx->sliceptr, sizeof(int) * ((size_t)nslices + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (permuted) {
        ctx->perm.resize(n);
        CK(cudaMemcpyAsync(ctx->perm.data(), ctx->d_perm, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        ctx->perm.clear();
    }
    CK(cudaStreamSynchronize(ctx->stream));
    lap("numbering");
    size_t nblk = (size_t)sp[nslices];
    // Work units of the PCG operator (cg_spmv_kernel): full rounds of 16-slice tiles, then the partial last round cut
    // into one equal-work unit per block; a slice costs its ELL blocks (2432 B each) + its 32 rows (256 B each).
    int nbs = grid_spmv(ctx, n);
    std::vector<int> hpart = spmv_units_of(sp, 0, nslices, nbs);
    ctx->spmv_units = (int)hpart.size() - 1;
    if ((long long)nblk > ctx->blkcap) {
        CK(dev_alloc(ctx->ecol, nblk * 32)); CK(dev_alloc(ctx->ewgt, nblk * 32)); CK(dev_alloc(ctx->Je, nblk * 288));
        ctx->blkcap = (long long)nblk;
    }
    if ((int)hpart.size() > ctx->spmv_cap) { CK(dev_alloc(ctx->spmv_part, hpart.size())); ctx->spmv_cap = (int)hpart.size(); }
    ctx->E = E; ctx->nblk = (long long)nblk;
    ctx->area = area; ctx->ntri = n_triangles;
    CK(cudaMemcpyAsync(ctx->spmv_part, hpart.data(), sizeof(int) * hpart.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0 && nblk > 0) {
        // Slot order inside a row: neighbours inside the row's own tile (the shared-memory window of the gather kernels)
        // first, halo neighbours last, each part ascending: the halo gathers concentrate in the last columns of a slice.
        ell_fill_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dperm, dinv, ctx->g_rp0, ctx->g_col0, ctx->g_w0, ctx->sliceptr, ctx->ecol, ctx->ewgt);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(ctx->Je, 0, sizeof(double) * nblk * 288, ctx->stream));
    }
    ctx->have_graph = true; ctx->have_rot = false;
    ctx->f32_fresh = false;
    drop_graphs(ctx);
    int urc = build_state(ctx);                        // (synchronises the stream: hpart may go out of scope)
    lap("ell + state");
    if (urc == DSC_OK && ctx->sharded) urc = shard_setup_graph(ctx, sp, nslices);
    return urc;
}


extern "C" int dsc_set_graph(dsc_ctx* ctx, int n, const int32_t* rowptr, const int32_t* col, const double* w,
                             double area, long long n_triangles, int reorder) {
    if (!ctx || !rowptr || n < 0) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_graph");
    if (!ctx->have_problem || n != ctx->n) return fail(ctx, DSC_ERR_STATE, "dsc_set_graph: upload a problem of the same size first");
    const bool validate = !(reorder & 2);
    reorder &= 1;
    const bool timing = std::getenv("DSC_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dsc_set_graph] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    if (!(area > 0.0) || n_triangles < 0) return fail(ctx, DSC_ERR_INVALID_ARG, "area must be > 0, n_triangles >= 0");
    CK(cudaSetDevice(ctx->device));
    long long E = n > 0 ? rowptr[n] : 0;
    if (rowptr[0] != 0 || E < 0 || (E > 0 && (!col || !w))) return fail(ctx, DSC_ERR_GRAPH, "rowptr");
    for (int i = 0; i < n; ++i) if (rowptr[i + 1] < rowptr[i]) return fail(ctx, DSC_ERR_GRAPH, "rowptr not monotone");
    // ---- raw CSR to the device as it is (h2d: straight from pinned / registered caller memory, else through the bounce buffer)
    { int rrc = reserve_graph(ctx, n, E); if (rrc) return rrc; }
    {
        int rc;
        if ((rc = h2d(ctx, ctx->g_rp0, rowptr, sizeof(int) * ((size_t)n + 1)))) return rc;
        if (E > 0 && (rc = h2d(ctx, ctx->g_col0, col, sizeof(int) * (size_t)E))) return rc;
        if (E > 0 && (rc = h2d(ctx, ctx->g_w0, w, sizeof(double) * (size_t)E))) return rc;
    }
    lap("csr staging");
    return set_graph_tail(ctx, n, E, area, n_triangles, reorder, validate, timing);
}

// Grow-only device workspace of the graph builders: one allocation, carved per call (ws_begin, ws_take...).  A warm context
// builds a graph without a single cudaMalloc / cudaFree (each costs 0.1 - 10 ms and serialises the device).
static cudaError_t ws_begin(dsc_ctx* ctx, size_t bytes) {
    ctx->ws_off = 0;
    if (bytes <= ctx->ws_cap) return cudaSuccess;
    if (ctx->ws) { cudaFree(ctx->ws); ctx->ws = nullptr; ctx->ws_cap = 0; }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->ws), want);
    if (e == cudaSuccess) ctx->ws_cap = want;
    return e;
}
template <typename T>
static T* ws_take(dsc_ctx* ctx, size_t count) {
    T* p = reinterpret_cast<T*>(ctx->ws + ctx->ws_off);
    ctx->ws_off += ws_round(count * sizeof(T));
    return p;
}

// ------------------------------------------------------------------ Delaunay graph on the GPU (dsc_delaunay.cuh)
// dX: device float [n][3]; leaves the CSR (caller numbering, rows ascending) in ctx->dl_*.
static int delaunay_build_device(dsc_ctx* ctx, int n, const float* dX, double min_weight) {
    ctx->dl_n = n; ctx->dl_E = 0; ctx->dl_ntri = 0; ctx->dl_area = 0.0; ctx->dl_uncertified = 0;
    if (!ctx->dl_rowptr || (size_t)n > ctx->dl_cap) {          // grow-only: row pointers + the < 6 n directed edges of a planar mesh
        const size_t cap = (size_t)n + (size_t)n / 8 + 16;
        ctx->dl_cap = 0;
        CK(dev_alloc(ctx->dl_rowptr, cap + 2));
        CK(dev_alloc(ctx->dl_col, 6 * cap));
        CK(dev_alloc(ctx->dl_w, 6 * cap));
        ctx->dl_cap = cap;
    }
    CK(cudaMemsetAsync(ctx->dl_rowptr, 0, sizeof(int) * ((size_t)n + 2), ctx->stream));
    if (n < 3) { CK(cudaStreamSynchronize(ctx->stream)); return DSC_OK; }
    // bounding box of (x, y) (two-stage device reduction, folded on the host) -> grid and the three ghost vertices
    const int nbv = grid_threads(ctx, n);
    bbox_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, ctx->d_box + 4);
    ctx->launches++;
    float* hb = reinterpret_cast<float*>(ctx->h_pinned);
    CK(cudaMemcpyAsync(hb, ctx->d_box + 4, sizeof(float) * 4 * nbv, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int b = 0; b < nbv; ++b) {
        x0 = std::min(x0, (double)hb[4 * b]); x1 = std::max(x1, (double)hb[4 * b + 1]);
        y0 = std::min(y0, (double)hb[4 * b + 2]); y1 = std::max(y1, (double)hb[4 * b + 3]);
    }
    const double wdt = x1 - x0, hgt = y1 - y0, dm = std::max(wdt, hgt);
    if (!(dm > 0.0) || !std::isfinite(dm)) return fail(ctx, DSC_ERR_INVALID_ARG, "Delaunay: the points have no extent in (x, y) or are not finite");
    // the search grid covers mean +- 4 sigma of the points (clipped to the bounding box): a handful of far outliers --
    // badly triangulated matches -- must not decide the cell size.  Points outside fall into the border cells (the ring
    // walk's distance bound stays valid for them; a query point outside is simply left to the second pass).
    double gx0 = x0, gx1 = x1, gy0 = y0, gy1 = y1;
    {
        const double mcx = 0.5 * (x0 + x1), mcy = 0.5 * (y0 + y1);
        delaunay_moments_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, mcx, mcy, ctx->part);
        ctx->launches++;
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * 4 * nbv, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        double mo[4];
        for (int k = 0; k < 4; ++k) mo[k] = host_sum(ctx->h_pinned, nbv, 4, k);
        const double mx = mo[0] / n, my = mo[1] / n;
        const double sx = std::sqrt(std::max(0.0, mo[2] / n - mx * mx)), sy = std::sqrt(std::max(0.0, mo[3] / n - my * my));
        if (sx > 0.0 && std::isfinite(sx)) { gx0 = std::max(x0, mcx + mx - 4.0 * sx); gx1 = std::min(x1, mcx + mx + 4.0 * sx); }
        if (sy > 0.0 && std::isfinite(sy)) { gy0 = std::max(y0, mcy + my - 4.0 * sy); gy1 = std::min(y1, mcy + my + 4.0 * sy); }
    }
    KnnGrid g{gx0, gy0, 1.0, 1, 1};
    {
        const double gw = gx1 - gx0, ghh = gy1 - gy0;
        double cell = std::sqrt(std::max(gw, 1e-300) * std::max(ghh, 1e-300) / std::max(1.0, n / 2.0));
        cell = std::max(cell, std::max(gw, ghh) / 4096.0);
        g.inv_cell = 1.0 / cell;
        g.nx = std::max(1, (int)std::ceil(gw / cell) + 1);
        g.ny = std::max(1, (int)std::ceil(ghh / cell) + 1);
    }
    const int ncells = g.nx * g.ny, m = std::max(ncells, n) + 1;
    // every work array comes out of the context's grow-only workspace
    const size_t nsums = (size_t)(m / kScanBlock + 2);
    {
        size_t need = 0;
        for (size_t c : {(size_t)n, (size_t)m, (size_t)m, (size_t)m, (size_t)n, nsums, (size_t)n * kDlMaxV, (size_t)n, (size_t)n, (size_t)n + 1, (size_t)n + 1,
                         (size_t)n, (size_t)n, (size_t)n + 1})
            need += ws_round(c * sizeof(int));
        need += ws_round((size_t)n);
        need += ws_round(sizeof(int) * (size_t)kDlWsSecond) + ws_round(sizeof(int) * (size_t)kDlWsSecond * kDlMaxExtra);
        cudaError_t e = ws_begin(ctx, need);
        if (e != cudaSuccess) return fail(ctx, DSC_ERR_ALLOC, cudaGetErrorString(e));
    }
    int *cell = ws_take<int>(ctx, n), *cnt = ws_take<int>(ctx, m), *start = ws_take<int>(ctx, m), *cursor = ws_take<int>(ctx, m),
        *order = ws_take<int>(ctx, n), *sums = ws_take<int>(ctx, nsums), *star = ws_take<int>(ctx, (size_t)n * kDlMaxV), *deg = ws_take<int>(ctx, n),
        *flag = ws_take<int>(ctx, n), *isu = ws_take<int>(ctx, (size_t)n + 1), *pos = ws_take<int>(ctx, (size_t)n + 1), *slot = ws_take<int>(ctx, n),
        *nextra = ws_take<int>(ctx, n), *rowdeg = ws_take<int>(ctx, (size_t)n + 1);
    unsigned char* dup = ws_take<unsigned char>(ctx, n);
    // second pass only: room for kDlWsSecond uncertified cells (the rim: ~2 sqrt(n)) is part of the workspace, more is allocated
    int *list = ws_take<int>(ctx, kDlWsSecond), *extra = ws_take<int>(ctx, (size_t)kDlWsSecond * kDlMaxExtra), *list_own = nullptr, *extra_own = nullptr;
    auto release = [&]() { dev_free(list_own); dev_free(extra_own); };
    const bool timing = std::getenv("DSC_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx->stream);
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[dsc_delaunay] %-18s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    auto scan = [&](int count, const int* in, int* out) {      // exclusive scan of count ints
        int nb = (count + kScanBlock - 1) / kScanBlock;
        scan_block_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(count, in, out, sums);
        scan_sums_kernel<<<1, kScanBlock, 0, ctx->stream>>>(nb, sums);
        scan_add_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(count, out, sums);
        ctx->launches += 3;
    };
    cudaError_t e = cudaSuccess;
    auto bail = [&](int code, const char* what) { release(); return fail(ctx, code, what); };
    CK(cudaMemsetAsync(cnt, 0, sizeof(int) * m, ctx->stream));
    CK(cudaMemsetAsync(cursor, 0, sizeof(int) * m, ctx->stream));
    knn_cell_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, g, cell, cnt);
    scan(ncells, cnt, start);
    knn_scatter_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, cell, start, cursor, order);
    delaunay_duplicates_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, g, start, order, ncells, dup);
    ctx->launches++;
    lap("grid");
    const double box = std::ldexp(dm, 40);                     // "the whole plane": 2^40 x the extent of the point cloud
    // ---- pass one: every cell from the points around it
    delaunay_cells_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, dX, g, start, order, ncells, box, dup, nullptr, 0, nullptr, nullptr, star, deg, flag);
    lap("cells, pass one");
    // ---- the uncertified points, in index order
    delaunay_flag_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, flag, isu);
    CK(cudaMemsetAsync(isu + n, 0, sizeof(int), ctx->stream));
    scan(n + 1, isu, pos);
    ctx->launches += 4;
    int nu = 0;
    CK(cudaMemcpyAsync(&nu, pos + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (cudaGetLastError() != cudaSuccess) return bail(DSC_ERR_CUDA, "Delaunay pass one");
    ctx->dl_uncertified = nu;
    if (nu > 0) {
        // ---- pass two: finish them from the certified points that list them and from each other
        if (nu > kDlWsSecond) {
            if ((e = dev_alloc(list_own, (size_t)nu)) || (e = dev_alloc(extra_own, (size_t)nu * kDlMaxExtra))) { release(); return fail(ctx, DSC_ERR_ALLOC, cudaGetErrorString(e)); }
            list = list_own; extra = extra_own;
        }
        CK(cudaMemsetAsync(nextra, 0, sizeof(int) * n, ctx->stream));
        delaunay_compact_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, isu, pos, list, slot);
        delaunay_extra_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, star, deg, flag, slot, extra, nextra);
        delaunay_sort_extra_kernel<<<grid_threads(ctx, nu), kThreads, 0, ctx->stream>>>(nu, list, extra, nextra);
        delaunay_cells_kernel<<<(nu + 127) / 128, 128, 0, ctx->stream>>>(n, dX, g, start, order, ncells, box, dup, list, nu, extra, nextra, star, deg, flag);
        lap("cells, pass two");
        delaunay_flag_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, flag, isu);
        scan(n + 1, isu, pos);
        ctx->launches += 5;
        int bad = 0;
        CK(cudaMemcpyAsync(&bad, pos + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (bad > 0) return bail(DSC_ERR_GRAPH, "Delaunay: a cell has more than 32 vertices or more than 96 certified neighbours (degenerate input: use the host mesh)");
    }
    // ---- rows: degrees + triangle count + area, scan, fill (ascending, cot weights)
    delaunay_edges_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, star, deg, min_weight, 0, nullptr, rowdeg, nullptr, nullptr, ctx->part);
    CK(cudaMemsetAsync(rowdeg + n, 0, sizeof(int), ctx->stream));
    scan(n + 1, rowdeg, ctx->dl_rowptr);
    ctx->launches++;
    int E = 0;
    CK(cudaMemcpyAsync(&E, ctx->dl_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * 2 * nbv, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dl_ntri = (long long)host_sum(ctx->h_pinned, nbv, 2, 0);
    ctx->dl_area = host_sum(ctx->h_pinned, nbv, 2, 1);
    if ((size_t)E > 6 * ctx->dl_cap) { release(); return fail(ctx, DSC_ERR_GRAPH, "Delaunay: more than 6 n directed edges (not a planar mesh)"); }
    delaunay_edges_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, dX, star, deg, min_weight, 1, ctx->dl_rowptr, nullptr, ctx->dl_col, ctx->dl_w, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    lap("rows");
    release();
    ctx->dl_E = E;
    return DSC_OK;
}

extern "C" int dsc_delaunay_build(dsc_ctx* ctx, int n, const float* X, double min_weight, long long* n_edges, long long* n_triangles, double* area,
                                  long long* n_second_pass) {
    if (!ctx || n < 0 || (n > 0 && !X)) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_delaunay_build");
    CK(cudaSetDevice(ctx->device));
    if (n > 0 && (size_t)n > ctx->dl_xcap) {                    // grow-only staging of the caller's points
        ctx->dl_xcap = 0;
        CK(dev_alloc(ctx->dl_X, 3 * ((size_t)n + (size_t)n / 8 + 16)));
        ctx->dl_xcap = (size_t)n + (size_t)n / 8 + 16;
    }
    if (n > 0) { int rc = h2d(ctx, ctx->dl_X, X, sizeof(float) * 3 * (size_t)n); if (rc) return rc; }
    int rc = delaunay_build_device(ctx, n, ctx->dl_X, min_weight);
    if (rc) return rc;
    if (n_edges) *n_edges = ctx->dl_E;
    if (n_triangles) *n_triangles = ctx->dl_ntri;
    if (area) *area = ctx->dl_area;
    if (n_second_pass) *n_second_pass = ctx->dl_uncertified;
    return DSC_OK;
}

extern "C" int dsc_delaunay_download(dsc_ctx* ctx, int32_t* rowptr, int32_t* col, double* w) {
    if (!ctx || !rowptr) return DSC_ERR_INVALID_ARG;
    if (!ctx->dl_rowptr) return fail(ctx, DSC_ERR_STATE, "dsc_delaunay_download before dsc_delaunay_build");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(rowptr, ctx->dl_rowptr, sizeof(int) * ((size_t)ctx->dl_n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (col && ctx->dl_E) CK(cudaMemcpyAsync(col, ctx->dl_col, sizeof(int) * (size_t)ctx->dl_E, cudaMemcpyDeviceToHost, ctx->stream));
    if (w && ctx->dl_E) CK(cudaMemcpyAsync(w, ctx->dl_w, sizeof(double) * (size_t)ctx->dl_E, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

// The reference's graph for the uploaded problem, built where the problem lies: Delaunay of the KF1 map points' world (x, y)
// (extractPositions + ComputeDelaunayTriangulation3D), cot weights, area, triangle count -> the same set-up as dsc_set_graph.
extern "C" int dsc_set_graph_delaunay(dsc_ctx* ctx, double min_weight, int reorder, double* area, long long* n_triangles, long long* n_edges) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_problem) return fail(ctx, DSC_ERR_STATE, "dsc_set_graph_delaunay: upload a problem first");
    CK(cudaSetDevice(ctx->device));
    const bool timing = std::getenv("DSC_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    const int n = ctx->n;
    int rc = delaunay_build_device(ctx, n, ctx->X1f, min_weight);
    if (rc) return rc;
    if (!(ctx->dl_area > 0.0)) return fail(ctx, DSC_ERR_GRAPH, "Delaunay: the mesh has no area (fewer than 3 points, or all collinear)");
    if ((rc = reserve_graph(ctx, n, ctx->dl_E))) return rc;
    CK(cudaMemcpyAsync(ctx->g_rp0, ctx->dl_rowptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->dl_E) {
        CK(cudaMemcpyAsync(ctx->g_col0, ctx->dl_col, sizeof(int) * (size_t)ctx->dl_E, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->g_w0, ctx->dl_w, sizeof(double) * (size_t)ctx->dl_E, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (timing) {
        CK(cudaStreamSynchronize(ctx->stream));
        std::fprintf(stderr, "[dsc_set_graph_delaunay] mesh %8.2f ms (%lld second-pass cells)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ctx->dl_uncertified);
    }
    if (area) *area = ctx->dl_area;
    if (n_triangles) *n_triangles = ctx->dl_ntri;
    if (n_edges) *n_edges = ctx->dl_E;
    return set_graph_tail(ctx, n, ctx->dl_E, ctx->dl_area, ctx->dl_ntri, reorder & 1, !(reorder & 2), timing);
}

extern "C" int dsc_compute_rotations(dsc_ctx* ctx) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_graph) return fail(ctx, DSC_ERR_STATE, "dsc_compute_rotations before dsc_set_graph");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n > 0) {
        rotations_ell_kernel<<<grid_tiles(ctx, ctx->n, DSC_ROT_BLOCKS), kEllThreads, kWinBytes, ctx->stream>>>(ctx->n, ctx->P, ctx->sliceptr, ctx->ecol, ctx->ewgt, ctx->Q);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    ctx->have_rot = true;
    return DSC_OK;
}

extern "C" int dsc_get_rotations(dsc_ctx* ctx, double* quat) {
    if (!ctx || !quat) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_rot) return fail(ctx, DSC_ERR_STATE, "no rotations yet");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    std::vector<double> tmp(4 * (size_t)n);
    if (n) CK(cudaMemcpyAsync(tmp.data(), ctx->Q, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        int d = ctx->perm.empty() ? i : ctx->perm[i];
        for (int k = 0; k < 4; ++k) quat[4 * (size_t)d + k] = tmp[4 * (size_t)i + k];
    }
    return DSC_OK;
}

extern "C" int dsc_set_rotations(dsc_ctx* ctx, const double* quat) {
    if (!ctx || !quat) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_graph) return fail(ctx, DSC_ERR_STATE, "dsc_set_rotations before dsc_set_graph");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    std::vector<double> tmp(4 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int s = ctx->perm.empty() ? i : ctx->perm[i];
        for (int k = 0; k < 4; ++k) tmp[4 * (size_t)i + k] = quat[4 * (size_t)s + k];
    }
    if (n) CK(cudaMemcpyAsync(ctx->Q, tmp.data(), sizeof(double) * 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_rot = true;
    return DSC_OK;
}

extern "C" int dsc_reset_state(dsc_ctx* ctx) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_problem) return fail(ctx, DSC_ERR_STATE, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n) CK(cudaMemcpyAsync(ctx->P, ctx->P0, sizeof(double) * 8 * ctx->n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->Gcur, &ctx->g0, sizeof(Globals), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

extern "C" int dsc_set_pcg(dsc_ctx* ctx, const dsc_pcg_params* prm) {
    if (!ctx || !prm || !(prm->rtol > 0.0) || prm->max_iters < 1 || prm->check_every < 1) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_pcg");
    ctx->pcg = *prm;
    return DSC_OK;
}

extern "C" int dsc_set_solver(dsc_ctx* ctx, int solver) {
    if (!ctx || solver < DSC_SOLVER_AUTO || solver > DSC_SOLVER_DENSE) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_solver");
    ctx->solver = solver;
    return DSC_OK;
}

extern "C" int dsc_set_precision(dsc_ctx* ctx, int precision) {
    if (!ctx || (precision != DSC_PRECISION_F64 && precision != DSC_PRECISION_F32)) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_precision");
    if (precision != ctx->precision) drop_graphs(ctx);
    ctx->precision = precision;
    return DSC_OK;
}

extern "C" int dsc_set_early_reject(dsc_ctx* ctx, int n_levels, const double* rtol_loose, const double* rho_margin) {
    if (!ctx || n_levels < 0 || n_levels > 4 || (n_levels > 0 && (!rtol_loose || !rho_margin)))
        return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_early_reject");
    for (int l = 0; l < n_levels; ++l) {
        if (!(rtol_loose[l] > 0.0) || !(rho_margin[l] >= 0.0) || (l > 0 && !(rtol_loose[l] < rtol_loose[l - 1])))
            return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_set_early_reject: tolerances must be positive and strictly decreasing, margins >= 0");
    }
    ctx->early_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) { ctx->early_rtol[l] = rtol_loose[l]; ctx->early_margin[l] = rho_margin[l]; }
    return DSC_OK;
}

static int ready(dsc_ctx* ctx, const dsc_weights* w) {
    if (!ctx || !w) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_problem || !ctx->have_graph || !ctx->have_rot)
        return fail(ctx, DSC_ERR_STATE, "need dsc_problem_upload, dsc_set_graph and rotations first");
    if (!(w->depth_sigma > 0.f) || !std::isfinite(w->depth_sigma))
        return fail(ctx, DSC_ERR_INVALID_ARG, "depth_sigma must be finite and > 0 (the reference divides by it, g2oBundleAdjustment.cc:824)");
    if (ctx->sharded) {
        if (!ctx->sh_attached) return fail(ctx, DSC_ERR_SHARD, "dsc_shard_attach has not been called");
        if (ctx->precision != DSC_PRECISION_F64) return fail(ctx, DSC_ERR_INVALID_ARG, "a sharded context runs the fp64 PCG path only");
        if (ctx->solver == DSC_SOLVER_DENSE) return fail(ctx, DSC_ERR_INVALID_ARG, "a sharded context runs the PCG path only");
    }
    return DSC_OK;
}

// ---- point-sharded pair: helpers of the solve path
static int shard_tile0(const dsc_ctx* ctx) { return ctx->sh.row_begin[ctx->sh.rank] / kSortGroup; }
static int shard_tile1(const dsc_ctx* ctx) { return (ctx->sh.row_begin[ctx->sh.rank + 1] + kSortGroup - 1) / kSortGroup; }
static int shard_rows(const dsc_ctx* ctx) { return ctx->sh.row_begin[ctx->sh.rank + 1] - ctx->sh.row_begin[ctx->sh.rank]; }
static int shard_pidx(const dsc_ctx* ctx, const double* P) { return P == ctx->sh.Pbuf[0][ctx->sh.rank] ? 0 : 1; }
// sum (entry maxidx: maximum) over the blocks of this rank and over the ranks -> ctx->sh_out[count]
static void shard_allreduce(dsc_ctx* ctx, int cls, const double* part, int nb, int stride, int count, int maxidx = -1) {
    shard_allreduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->sh, cls, part, nb, stride, count, maxidx, ctx->sh_out);
    ctx->launches++;
}
// DSC_ERR_SHARD if a wait on a peer has timed out (call after a stream synchronisation point is acceptable)
static int shard_check(dsc_ctx* ctx) {
    if (!ctx->sharded) return DSC_OK;
    int* hp = reinterpret_cast<int*>(ctx->h_pinned + 7 * kMaxBlocks);
    CK(cudaMemcpyAsync(hp, ctx->sh_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*hp) return fail(ctx, DSC_ERR_SHARD, "a peer rank did not answer within the time limit (DSC_SHARD_TIMEOUT_S)");
    return DSC_OK;
}

static int eval_cost(dsc_ctx* ctx, const WeightsDev& W, const double* P, const Globals* G, double* chi2, double* parts) {
    int nb = grid_tiles(ctx, ctx->n, DSC_COST_BLOCKS);
    if (ctx->sharded) {                                  // the rank's tiles, then the totals over the ranks (one "block" of 3)
        nb = std::max(1, std::min(shard_tile1(ctx) - shard_tile0(ctx), ctx->sms * DSC_COST_BLOCKS));
        cost_ell_kernel<<<nb, kEllThreads, kWinBytes, ctx->stream>>>(ctx->n, P, ctx->Q, ctx->uv, ctx->dm, ctx->isg, ctx->sliceptr, ctx->ecol,
                                                                   ctx->ewgt, G, ctx->pair, W, ctx->part, shard_tile0(ctx), shard_tile1(ctx));
        shard_allreduce(ctx, SF_C, ctx->part, nb, 3, 3);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->sh_out, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
        nb = 1;
    } else {
        cost_ell_kernel<<<nb, kEllThreads, kWinBytes, ctx->stream>>>(ctx->n, P, ctx->Q, ctx->uv, ctx->dm, ctx->isg, ctx->sliceptr, ctx->ecol,
                                                                   ctx->ewgt, G, ctx->pair, W, ctx->part, 0, -1);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * 3 * nb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    double p0 = host_sum(ctx->h_pinned, nb, 3, 0), p1 = host_sum(ctx->h_pinned, nb, 3, 1), p2 = host_sum(ctx->h_pinned, nb, 3, 2);
    if (parts) { parts[0] = p0; parts[1] = p1; parts[2] = p2; }
    *chi2 = p0 + p1 + p2;
    return DSC_OK;
}

extern "C" int dsc_cost(dsc_ctx* ctx, const dsc_weights* w, double* chi2, double* parts) {
    int s = ready(ctx, w);
    if (s) return s;
    if (!chi2) return DSC_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->n == 0) { *chi2 = 0.0; if (parts) parts[0] = parts[1] = parts[2] = 0.0; return DSC_OK; }
    int rc = eval_cost(ctx, make_weights(ctx, w), ctx->P, ctx->Gcur, chi2, parts);
    return rc ? rc : shard_check(ctx);
}

// The PCG's data in the two precisions: T = double -> Je, U, Minv, vec[]; T = float -> JeF, UF, MinvF, vecF[] (fp32 mode).
template <typename T> struct Sel;
template <> struct Sel<double> {
    static double* Je(dsc_ctx* c) { return c->Je; } static double* U(dsc_ctx* c) { return c->U; }
    static double* Minv(dsc_ctx* c) { return c->Minv; } static double* vec(dsc_ctx* c, int k) { return c->vec[k]; }
};
template <> struct Sel<float> {
    static float* Je(dsc_ctx* c) { return c->JeF; } static float* U(dsc_ctx* c) { return c->UF; }
    static float* Minv(dsc_ctx* c) { return c->MinvF; } static float* vec(dsc_ctx* c, int k) { return c->vecF[k]; }
};
template <typename T = double>
static CgVecsT<T> make_vecs(dsc_ctx* ctx) {
    CgVecsT<T> v;
    v.x = Sel<T>::vec(ctx, 0); v.r = Sel<T>::vec(ctx, 1); v.z = Sel<T>::vec(ctx, 2); v.w = Sel<T>::vec(ctx, 3); v.p = Sel<T>::vec(ctx, 4); v.s = Sel<T>::vec(ctx, 5);
    v.xg = ctx->small; v.rg = ctx->small + 8; v.zg = ctx->small + 16; v.wg = ctx->small + 24; v.pg = ctx->small + 32; v.sg = ctx->small + 40;
    return v;
}
static bool f32(const dsc_ctx* ctx) { return ctx->precision == DSC_PRECISION_F32; }
static double* refine_Xg(dsc_ctx* ctx) { return ctx->small + 112; }        // globals of the accumulated solution X (vec[0])
static double* refine_Xg_peek(dsc_ctx* ctx) { return ctx->small + 120; }   // globals of the peeked solution (vec[1])
// float buffers of the fp32 mode, (re)allocated for the uploaded problem; JeF's padding records are zeroed once per graph
static int ensure_f32(dsc_ctx* ctx) {
    if (!f32(ctx)) return DSC_OK;
    if (ctx->cap > ctx->f32_cap) {
        const size_t N = (size_t)ctx->cap, S32 = 32 * ((N + 31) / 32);
        CK(dev_alloc(ctx->UF, (size_t)kURec * S32)); CK(dev_alloc(ctx->MinvF, 21 * S32));
        for (auto& v : ctx->vecF) CK(dev_alloc(v, 6 * N + 8));
        ctx->f32_cap = ctx->cap;
    }
    if (ctx->blkcap > ctx->f32_blkcap) {
        CK(dev_alloc(ctx->JeF, (size_t)ctx->blkcap * 288));
        ctx->f32_blkcap = ctx->blkcap;
        ctx->f32_fresh = false;
    }
    if (!ctx->f32_fresh && ctx->nblk > 0) {
        CK(cudaMemsetAsync(ctx->JeF, 0, sizeof(float) * (size_t)ctx->nblk * 288, ctx->stream));
        drop_graphs(ctx);
    }
    ctx->f32_fresh = true;
    return DSC_OK;
}

static void launch_linearize(dsc_ctx* ctx, const WeightsDev& W, int nb) {
    const int t0 = ctx->sharded ? shard_tile0(ctx) : 0, t1 = ctx->sharded ? shard_tile1(ctx) : -1;
    if (f32(ctx))
        linearize_ell_kernel<true><<<nb, kLinThreads, kWinBytes, ctx->stream>>>(ctx->n, ctx->P, ctx->Q, ctx->uv, ctx->dm, ctx->isg, ctx->sliceptr, ctx->ecol,
                                                                              ctx->ewgt, ctx->Gcur, ctx->pair, W, ctx->b, ctx->D, ctx->U, ctx->Je, ctx->part, ctx->UF, ctx->JeF, t0, t1);
    else
        linearize_ell_kernel<false><<<nb, kLinThreads, kWinBytes, ctx->stream>>>(ctx->n, ctx->P, ctx->Q, ctx->uv, ctx->dm, ctx->isg, ctx->sliceptr, ctx->ecol,
                                                                               ctx->ewgt, ctx->Gcur, ctx->pair, W, ctx->b, ctx->D, ctx->U, ctx->Je, ctx->part, nullptr, nullptr, t0, t1);
}
static int run_linearize(dsc_ctx* ctx, const WeightsDev& W, LinGlobal* hlin) {
    int nb = grid_tiles(ctx, ctx->n, 1);
    { int erc = ensure_f32(ctx); if (erc) return erc; }
    if (ctx->sharded) {                                  // the rank's tiles; totals over the ranks = one "block" of partials
        nb = std::max(1, std::min(shard_tile1(ctx) - shard_tile0(ctx), ctx->sms));
        launch_linearize(ctx, W, nb);
        shard_allreduce(ctx, SF_L, ctx->part, nb, kLinPart, kLinPart, 3);
        finalize_linearize_kernel<<<1, kThreads, 0, ctx->stream>>>(1, ctx->sh_out, ctx->lin);
    } else {
        launch_linearize(ctx, W, nb);
        finalize_linearize_kernel<<<1, kThreads, 0, ctx->stream>>>(nb, ctx->part, ctx->lin);
    }
    ctx->launches += 2;
    CK(cudaGetLastError());
    LinGlobal* hp = reinterpret_cast<LinGlobal*>(ctx->h_pinned + 6 * kMaxBlocks);
    CK(cudaMemcpyAsync(hp, ctx->lin, sizeof(LinGlobal), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *hlin = *hp;
    return DSC_OK;
}

// PCG solve of (H + lambda I) dx = b; returns iterations, status DSC_OK / DSC_ERR_PCG_BREAKDOWN
// PCG solve of (H + lambda I) dx = b in two entry points so that a solve can be paused at a loose tolerance,
// inspected (trial cost) and resumed to the tight one: begin = preconditioner + r0/z0 + first operator
// application; resume = iterate until sqrt(r.z / r0.z0) <= rtol, breakdown or max_iters.
static bool grid_pcg(const dsc_ctx* ctx) { return ctx->grid_blocks > 0 && ctx->n > ctx->small_max_rows && ctx->n <= ctx->grid_max_rows; }
static bool small_active(const dsc_ctx* ctx) {
    return !f32(ctx) && !ctx->sharded && ((ctx->small_cluster > 0 && ctx->n <= ctx->small_max_rows) || grid_pcg(ctx));
}

// one launch = the whole solve (or its continuation after a pause) by a single thread-block cluster (dsc_small.cuh)
static int small_launch(dsc_ctx* ctx, const WeightsDev& W, int fresh) {
    CgVecs v = make_vecs(ctx);
    double* Ginv = ctx->small + 48;
    if (grid_pcg(ctx)) {                                  // the whole device, grid barriers (pcg_grid_kernel)
        int n = ctx->n, mi = ctx->pcg.max_iters;
        const double *P = ctx->P, *Je = ctx->Je, *U = ctx->U, *b = ctx->b, *D = ctx->D;
        const int *sp = ctx->sliceptr, *ec = ctx->ecol;
        const Globals* G = ctx->Gcur;
        const LinGlobal* lin = ctx->lin;
        PairDev pr = ctx->pair;
        WeightsDev Wc = W;
        void* args[] = {&n, &fresh, &mi, &P, &Je, &U, &sp, &ec, &G, &pr, &Wc, &b, &D, &lin, &ctx->Minv, &Ginv, &ctx->errflag, &v,
                        &ctx->gpart[0], &ctx->gpart[1], &ctx->dpart, &ctx->bpart, &ctx->ctl};
        const int blocks = std::min(ctx->grid_blocks, std::max(1, (n + 31) / 32));
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)pcg_grid_kernel, dim3(blocks), dim3(kThreads), args, 0, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, DSC_ERR_CUDA, std::string("pcg_grid_kernel -> ") + cudaGetErrorString(e));
        ctx->launches++;
        return DSC_OK;
    }
    for (;;) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctx->small_cluster); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = ctx->small_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, pcg_cluster_kernel, ctx->n, fresh, ctx->pcg.max_iters, (const double*)ctx->P, (const double*)ctx->Je,
                                           (const double*)ctx->U, (const int*)ctx->sliceptr, (const int*)ctx->ecol, (const Globals*)ctx->Gcur, ctx->pair, W,
                                           (const double*)ctx->b, (const double*)ctx->D, (const LinGlobal*)ctx->lin, ctx->Minv, Ginv, ctx->errflag, v,
                                           ctx->gpart[0], ctx->gpart[1], ctx->dpart, ctx->bpart, ctx->ctl);
        if (e == cudaSuccess) break;
        cudaGetLastError();
        if (fresh && ctx->small_cluster > 8) { ctx->small_cluster = 8; continue; }      // 16 CTAs not schedulable here: use the portable size
        return fail(ctx, DSC_ERR_CUDA, std::string("pcg_cluster_kernel -> ") + cudaGetErrorString(e));
    }
    ctx->launches++;
    return DSC_OK;
}

// the three PCG kernels in the context's precision (T = storage type of Je, U, Minv and the vectors)
template <typename T>
static void launch_init(dsc_ctx* ctx, double lambda) {
    if (sizeof(T) == 8 && ctx->minv_float) {
        cg_init_kernel<double, float><<<grid_threads(ctx, (long long)ctx->n), kThreads, 0, ctx->stream>>>(ctx->n, ctx->b, ctx->D, lambda, ctx->lin, ctx->MinvS, ctx->small + 48,
                                                                                                       ctx->errflag, make_vecs<double>(ctx), ctx->gpart[0], ctx->ctl);
        return;
    }
    cg_init_kernel<T><<<grid_threads(ctx, (long long)ctx->n), kThreads, 0, ctx->stream>>>(ctx->n, ctx->b, ctx->D, lambda, ctx->lin, Sel<T>::Minv(ctx), ctx->small + 48,
                                                                                       ctx->errflag, make_vecs<T>(ctx), ctx->gpart[0], ctx->ctl);
}
// w = (H + lambda I) z in precision T; z / zg / w default to the PCG's own vectors
template <typename T>
static void launch_spmv(dsc_ctx* ctx, const WeightsDev& W, double lambda, const CgControl* ctl, const T* z = nullptr, const double* zg = nullptr, T* w = nullptr) {
    CgVecsT<T> v = make_vecs<T>(ctx);
    cg_spmv_kernel<T><<<grid_spmv(ctx, ctx->n), kThreads, SpmvCfg<T>::kSmem, ctx->stream>>>(ctx->n, ctx->P, Sel<T>::Je(ctx), Sel<T>::U(ctx), ctx->sliceptr, ctx->ecol, ctx->spmv_part,
                                                                                         ctx->spmv_units, ctx->Gcur, ctx->pair, W, lambda, z ? z : v.z, zg ? zg : v.zg,
                                                                                         w ? w : v.w, ctx->dpart, ctx->bpart, ctx->lin, ctl, ShardDev{});
}
template <typename T>
static void launch_update(dsc_ctx* ctx, int par, int first, double lambda, int gin, int gout, double rtol2) {
    if (sizeof(T) == 8 && ctx->minv_float) {
        cg_update_kernel<double, float><<<grid_threads(ctx, (long long)ctx->n), kThreads, 0, ctx->stream>>>(ctx->n, par, first, ctx->MinvS, ctx->small + 48, ctx->lin, lambda,
                                                                                                         make_vecs<double>(ctx), ctx->gpart[gin], ctx->gpart[gout], ctx->dpart,
                                                                                                         ctx->bpart, grid_spmv(ctx, ctx->n), ctx->ctl, rtol2);
        return;
    }
    cg_update_kernel<T><<<grid_threads(ctx, (long long)ctx->n), kThreads, 0, ctx->stream>>>(ctx->n, par, first, Sel<T>::Minv(ctx), ctx->small + 48, ctx->lin, lambda, make_vecs<T>(ctx),
                                                                                         ctx->gpart[gin], ctx->gpart[gout], ctx->dpart, ctx->bpart, grid_spmv(ctx, ctx->n),
                                                                                         ctx->ctl, rtol2);
}

// point-sharded pair: the same three launches over the rank's rows; zpar = parity of the z buffer the launch reads
static void shard_launch_spmv(dsc_ctx* ctx, const WeightsDev& W, double lambda, int zpar) {
    CgVecs v = make_vecs(ctx);
    const int nbs = grid_spmv(ctx, std::max(1, shard_rows(ctx)));
    cg_spmv_kernel<double, true><<<nbs, kThreads, SpmvCfg<double>::kSmem, ctx->stream>>>(ctx->n, ctx->P, ctx->Je, ctx->U, ctx->sliceptr, ctx->ecol, ctx->sh_part, ctx->sh_units,
                                                                                      ctx->Gcur, ctx->pair, W, lambda, ctx->sh.zbuf[zpar][ctx->sh.rank], v.zg, v.w,
                                                                                      ctx->dpart, ctx->bpart, ctx->lin, ctx->ctl, ctx->sh);
}
static void shard_launch_update(dsc_ctx* ctx, int k, int first) {
    shard_cg_update_kernel<<<grid_threads(ctx, std::max(1, shard_rows(ctx))), kThreads, 0, ctx->stream>>>(ctx->sh, ctx->n, k & 1, first, ctx->Minv, ctx->small + 48, ctx->lin,
                                                                                                      make_vecs(ctx), k & 1, ctx->gpart[(k + 1) & 1], ctx->ctl);
}

static int pcg_begin(dsc_ctx* ctx, const WeightsDev& W, double lambda) {
    if (small_active(ctx)) {                           // the cluster kernel starts the solve itself
        CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
        ctl_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, lambda, 1, 0.0, 0, 0);
        ctx->launches++;
        CK(cudaGetLastError());
        return DSC_OK;
    }
    CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
    ctl_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, lambda, 1, 0.0, 0, 0);
    if (ctx->sharded) {
        shard_cg_init_kernel<<<grid_threads(ctx, std::max(1, shard_rows(ctx))), kThreads, 0, ctx->stream>>>(ctx->sh, ctx->n, ctx->b, ctx->D, lambda, ctx->lin, ctx->Minv, ctx->small + 48,
                                                                                                        ctx->errflag, make_vecs(ctx), 0, ctx->gpart[0], ctx->ctl);
        shard_launch_spmv(ctx, W, lambda, 0);
    } else if (f32(ctx)) { launch_init<float>(ctx, lambda); launch_spmv<float>(ctx, W, lambda, ctx->ctl); ctx->rf = dsc_ctx::Refine{}; }
    else { launch_init<double>(ctx, lambda); launch_spmv<double>(ctx, W, lambda, ctx->ctl); }
    ctx->launches += 3;
    CK(cudaGetLastError());
    return DSC_OK;
}

// kGraphIters PCG iterations (first = 0, parity starting even) captured once per state buffer and replayed:
// lambda and the tolerance live in CgControl, every other argument is fixed for the uploaded problem.
constexpr int kGraphIters = 16;
constexpr int kPcgUnconverged = 1;       // internal status of pcg_resume: iteration limit reached without convergence
constexpr int kEarlyWorthIters = 16;     // early-reject pauses are skipped while full solves take no more than this
static int iteration_graph(dsc_ctx* ctx, const WeightsDev& W, cudaGraphExec_t* out) {
    dsc_ctx::IterGraph* slot = nullptr;
    for (auto& g : ctx->graphs) if (g.exec && g.P == ctx->P && g.precision == ctx->precision && std::memcmp(&g.W, &W, sizeof(WeightsDev)) == 0) { *out = g.exec; return DSC_OK; }
    for (auto& g : ctx->graphs) if (!g.exec || g.P == ctx->P) { slot = &g; break; }
    if (!slot) slot = &ctx->graphs[0];
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    for (int k = 2; k < 2 + kGraphIters; ++k) {
        if (ctx->sharded) { shard_launch_update(ctx, k, 0); shard_launch_spmv(ctx, W, 0.0, (k + 1) & 1); }
        else if (f32(ctx)) { launch_update<float>(ctx, k & 1, 0, 0.0, k & 1, (k + 1) & 1, 0.0); launch_spmv<float>(ctx, W, 0.0, ctx->ctl); }
        else { launch_update<double>(ctx, k & 1, 0, 0.0, k & 1, (k + 1) & 1, 0.0); launch_spmv<double>(ctx, W, 0.0, ctx->ctl); }
    }
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess) return fail(ctx, DSC_ERR_CUDA, std::string("graph capture -> ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { slot->exec = nullptr; return fail(ctx, DSC_ERR_CUDA, std::string("graph instantiate -> ") + cudaGetErrorString(e)); }
    slot->P = ctx->P; slot->W = W; slot->precision = ctx->precision;
    *out = slot->exec;
    return DSC_OK;
}

// The iteration loop of both precisions: updates k .. (pass-local in the fp32 mode) until the control block reports
// convergence to rtol (relative to the gamma0 of the running pass), breakdown, or `limit` updates.
static int pcg_iterate(dsc_ctx* ctx, const WeightsDev& W, double lambda, double rtol, int* k_io, int limit, CgControl* out) {
    const double rtol2 = rtol * rtol;
    int k = *k_io;
    // resuming after a pause (k > 0): the converged latch belongs to the looser tolerance
    ctl_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, 0.0, 0, rtol2, 1, k > 0 ? 1 : 0);
    CgControl hc{};
    int poll = std::min(k == 0 ? 18 : 16, ctx->pcg.check_every);   // first poll early (well-damped solves need ~10 iterations): 2 direct + one graph
    while (k < limit) {
        int chunk = std::min(poll, limit - k);
        poll = std::min(2 * poll, ctx->pcg.check_every);
        for (int c = 0; c < chunk;) {
            if (ctx->use_graphs && k >= 2 && (k & 1) == 0 && chunk - c >= kGraphIters) {
                cudaGraphExec_t exec = nullptr;
                int grc = iteration_graph(ctx, W, &exec);
                if (grc) return grc;
                CK(cudaGraphLaunch(exec, ctx->stream));
                k += kGraphIters; c += kGraphIters;
                ctx->launches += 2 * kGraphIters;
                continue;
            }
            if (ctx->sharded) { shard_launch_update(ctx, k, k == 0 ? 1 : 0); shard_launch_spmv(ctx, W, lambda, (k + 1) & 1); }
            else if (f32(ctx)) { launch_update<float>(ctx, k & 1, k == 0 ? 1 : 0, lambda, k & 1, (k + 1) & 1, rtol2); launch_spmv<float>(ctx, W, lambda, ctx->ctl); }
            else { launch_update<double>(ctx, k & 1, k == 0 ? 1 : 0, lambda, k & 1, (k + 1) & 1, rtol2); launch_spmv<double>(ctx, W, lambda, ctx->ctl); }
            ctx->launches += 2;
            ++k; ++c;
        }
        CK(cudaGetLastError());
        CgControl* hp = reinterpret_cast<CgControl*>(ctx->h_pinned + 5 * kMaxBlocks);      // pinned: no staging, no device-wide lock
        CK(cudaMemcpyAsync(hp, ctx->ctl, sizeof(CgControl), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        hc = *hp;
        k = hc.iters;       // updates actually executed: kernels launched after convergence were no-ops
        if (hc.converged || hc.breakdown) break;
    }
    *k_io = hc.iters;
    *out = hc;
    if (hc.breakdown) return DSC_ERR_PCG_BREAKDOWN;
    if (!hc.converged) return kPcgUnconverged;        // iteration limit reached: the caller treats the step as a failed solve
    return DSC_OK;
}

// fp32 mode: mixed-precision iterative refinement.  The float PCG runs in passes; a pass that is asked for more than
// three digits is followed by a VERIFICATION: X += x_f (double), r = b - (H + lambda I) X with the double operator, and
// the next pass solves for the correction of that true residual.  The float recursion loses the true residual after a
// few digits (its attainable accuracy is ~ kappa * 6e-8), the refinement restores it: the returned step satisfies the
// same tolerance on the TRUE preconditioned residual that the fp64 path reaches on its recursive one.
constexpr double kRefineInnerFloor = 1e-6;    // digits asked of the first float pass (its recursion follows the true residual that far)
constexpr double kRefineLaterFloor = 1e-4;    // ... of a later pass: the float rounding of the operator data limits its true gain to 1e-3 .. 1e-4
constexpr double kRefineTrust = 1e-4;         // a pass that needs no more than this is not verified
constexpr int kRefineMaxPasses = 16;
static int pcg_resume_f32(dsc_ctx* ctx, const WeightsDev& W, double lambda, double tol, int* k_io) {
    dsc_ctx::Refine& rf = ctx->rf;
    const int n = ctx->n, nbv = grid_threads(ctx, (long long)n), nbs = grid_spmv(ctx, n);
    CgVecsT<float> vf = make_vecs<float>(ctx);
    double* X = ctx->vec[0];
    double* Wd = ctx->vec[3];
    for (;;) {
        const double need = tol / rf.rho;                      // reduction of the true residual this call still owes
        if (need >= 1.0) { *k_io = rf.total + rf.pass_k; return DSC_OK; }
        const bool trust = need >= kRefineTrust || tol >= 5e-5;   // (the loose levels of the early rejection only need the sign of rho)
        const double inner = trust ? 0.5 * need : std::max(need, rf.passes == 0 ? kRefineInnerFloor : kRefineLaterFloor);
        int k = rf.pass_k;
        if (k == 0 && rf.passes > 0) {                          // first operator application of a restarted pass
            launch_spmv<float>(ctx, W, lambda, ctx->ctl);
            ctx->launches++;
        }
        CgControl hc{};
        const int rc = pcg_iterate(ctx, W, lambda, inner, &k, ctx->pcg.max_iters - rf.total, &hc);
        rf.pass_k = k;
        *k_io = rf.total + rf.pass_k;
        if (rf.passes == 0 && k > 0) rf.gamma0 = hc.gamma0;     // preconditioned norm of the right-hand side
        if (rc != DSC_OK) return rc;
        if (trust) return DSC_OK;
        // ---- verification: X += x_f, true residual, restart
        refine_accumulate_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, rf.have_X ? X : nullptr, vf.x, X, rf.have_X ? refine_Xg(ctx) : nullptr, vf.xg,
                                                                   refine_Xg(ctx), rf.pass_k > 0 ? 1 : 0);
        rf.have_X = true; rf.total += rf.pass_k; rf.pass_k = 0; rf.passes++;
        launch_spmv<double>(ctx, W, lambda, nullptr, X, refine_Xg(ctx), Wd);
        cg_restart_kernel<<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->b, Wd, ctx->MinvF, ctx->small + 48, ctx->lin, refine_Xg(ctx), ctx->bpart, nbs, vf, ctx->gpart[0], ctx->ctl);
        refine_gamma_kernel<<<1, kThreads, 0, ctx->stream>>>(nbv, ctx->gpart[0], ctx->ctl);
        ctx->launches += 4;
        CK(cudaGetLastError());
        CgControl* hp = reinterpret_cast<CgControl*>(ctx->h_pinned + 5 * kMaxBlocks);
        CK(cudaMemcpyAsync(hp, ctx->ctl, sizeof(CgControl), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const double rho_new = std::sqrt(hp->gamma_true / rf.gamma0);
        if (std::getenv("DSC_REFINE_LOG")) std::fprintf(stderr, "[dsc refine] pass %d its %d rho %.3e -> %.3e (tol %.1e)\n", rf.passes, rf.total, rf.rho, rho_new, tol);
        if (!std::isfinite(rho_new)) return DSC_ERR_PCG_BREAKDOWN;
        if (rho_new <= tol) { rf.rho = rho_new; return DSC_OK; }
        if (rf.passes >= kRefineMaxPasses || rho_new > 0.5 * rf.rho) return kPcgUnconverged;   // the refinement stalls: a failed solve
        rf.rho = rho_new;
    }
}

static int pcg_resume(dsc_ctx* ctx, const WeightsDev& W, double lambda, double rtol, int* k_io) {
    if (f32(ctx)) return pcg_resume_f32(ctx, W, lambda, rtol, k_io);
    CgControl hc{};
    if (small_active(ctx)) {
        // resuming after a pause (k > 0): the converged latch belongs to the looser tolerance
        ctl_set_kernel<<<1, 1, 0, ctx->stream>>>(ctx->ctl, 0.0, 0, rtol * rtol, 1, *k_io > 0 ? 1 : 0);
        int src = small_launch(ctx, W, *k_io == 0 ? 1 : 0);
        if (src) return src;
        CgControl* hp = reinterpret_cast<CgControl*>(ctx->h_pinned + 5 * kMaxBlocks);
        CK(cudaMemcpyAsync(hp, ctx->ctl, sizeof(CgControl), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        hc = *hp;
        *k_io = hc.iters;
        return hc.breakdown ? DSC_ERR_PCG_BREAKDOWN : (hc.converged ? DSC_OK : kPcgUnconverged);
    }
    return pcg_iterate(ctx, W, lambda, rtol, k_io, ctx->pcg.max_iters, &hc);
}

// ---- dense direct solve of small problems (dsc_dense.cuh)
static bool dense_active(const dsc_ctx* ctx) {
    if (f32(ctx) || ctx->sharded) return false;            // the fp32 mode and the point-sharded pair are modes of the PCG path
    return ctx->solver == DSC_SOLVER_DENSE || (ctx->solver == DSC_SOLVER_AUTO && ctx->n <= DSC_DENSE_AUTO_MAX);
}
// once per LM iteration: H (dense, lower) and the right-hand side from the linearisation
static int dense_prepare(dsc_ctx* ctx, const WeightsDev& W) {
    const int n = ctx->n, m = 6 * n + 8;
    if (m > ctx->dn_cap) {
        CK(dev_alloc(ctx->dnH, (size_t)m * m)); CK(dev_alloc(ctx->dnA, (size_t)m * m));
        CK(dev_alloc(ctx->dn_rhs, (size_t)m)); CK(dev_alloc(ctx->dn_sol, (size_t)m)); CK(dev_alloc(ctx->dn_l11, (size_t)kDenseNB * kDenseNB));
        ctx->dn_cap = m;
    }
    CK(cudaMemsetAsync(ctx->dnH, 0, sizeof(double) * (size_t)m * m, ctx->stream));
    dense_assemble_kernel<<<grid_threads(ctx, n), kThreads, 0, ctx->stream>>>(n, m, ctx->P, ctx->D, ctx->U, ctx->Je, ctx->sliceptr, ctx->ecol,
                                                                            ctx->Gcur, ctx->pair, W, ctx->lin, ctx->dnH);
    dense_rhs_kernel<<<grid_threads(ctx, m), kThreads, 0, ctx->stream>>>(n, ctx->b, ctx->lin, ctx->dn_rhs);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return DSC_OK;
}
// (H + lambda I) dx = b by a blocked Cholesky; DSC_ERR_PCG_BREAKDOWN when the matrix is not positive definite
static int dense_solve(dsc_ctx* ctx, double lambda) {
    const int n = ctx->n, m = 6 * n + 8;
    CgVecs v = make_vecs(ctx);
    CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
    dense_shift_kernel<<<grid_threads(ctx, (long long)m * m), kThreads, 0, ctx->stream>>>(m, ctx->dnH, lambda, ctx->dnA);
    ctx->launches++;
    for (int k0 = 0; k0 < m; k0 += kDenseNB) {
        const int nb = std::min(kDenseNB, m - k0);
        const int rest = m - (k0 + nb);
        const int pblocks = std::max(1, (rest + kPanelThreads - 1) / kPanelThreads);
        dense_panel_kernel<<<pblocks, kPanelThreads, 0, ctx->stream>>>(m, k0, nb, ctx->dnA, ctx->dn_l11, ctx->errflag);
        ctx->launches++;
        if (rest > 0) {                                        // (pblocks > 1 implies rest > 0: the copy of L11 always happens)
            const int nt = (rest + kDenseNB - 1) / kDenseNB;
            dense_syrk_kernel<<<dim3(nt, nt), kDenseNB * 8, 0, ctx->stream>>>(m, k0, nb, ctx->dnA, ctx->dn_l11, pblocks > 1 ? 1 : 0);
            ctx->launches++;
        }
    }
    dense_solve_kernel<<<1, 1024, sizeof(double) * (size_t)m, ctx->stream>>>(m, ctx->dnA, ctx->dn_rhs, ctx->dn_sol);
    dense_scatter_kernel<<<grid_threads(ctx, m), kThreads, 0, ctx->stream>>>(n, ctx->dn_sol, v.x, v.xg);
    ctx->launches += 2;
    CK(cudaGetLastError());
    int* hp = reinterpret_cast<int*>(ctx->h_pinned + 5 * kMaxBlocks);
    CK(cudaMemcpyAsync(hp, ctx->errflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return *hp ? DSC_ERR_PCG_BREAKDOWN : DSC_OK;
}

// trial state x (+) dx, its robust chi2 and the rho denominator dx.(lambda dx + b) + 1e-3
static int eval_trial(dsc_ctx* ctx, const WeightsDev& W, double lambda, double* temp, double* scale) {
    int nbv = grid_threads(ctx, ctx->n);
    CgVecs v = make_vecs(ctx);
    if (ctx->sharded) {  // own rows + halo rows into the peers' trial buffers; the exchange publishes them and sums the scale
        nbv = grid_threads(ctx, std::max(1, shard_rows(ctx)));
        shard_apply_update_kernel<<<nbv, kThreads, 0, ctx->stream>>>(ctx->sh, ctx->n, ctx->P, v.x, v.xg, ctx->b, ctx->lin, lambda, ctx->Gcur,
                                                                    shard_pidx(ctx, ctx->Ptrial), ctx->Gtrial, ctx->part);
        shard_allreduce(ctx, SF_P, ctx->part, nbv, 1, 1);
        ctx->launches++;
        CK(cudaMemcpyAsync(ctx->h_pinned + 3 * kMaxBlocks, ctx->sh_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        int rc = eval_cost(ctx, W, ctx->Ptrial, ctx->Gtrial, temp, nullptr);
        if (rc) return rc;
        *scale = ctx->h_pinned[3 * kMaxBlocks] + 1e-3;
        return DSC_OK;
    }
    if (f32(ctx)) {     // the step is X (+ the running correction x_f): peek at it without disturbing the solve
        const dsc_ctx::Refine& rf = ctx->rf;
        refine_accumulate_kernel<<<nbv, kThreads, 0, ctx->stream>>>(ctx->n, rf.have_X ? ctx->vec[0] : nullptr, ctx->vecF[0], ctx->vec[1],
                                                                   rf.have_X ? refine_Xg(ctx) : nullptr, v.xg, refine_Xg_peek(ctx), rf.pass_k > 0 ? 1 : 0);
        apply_update_kernel<double><<<nbv, kThreads, 0, ctx->stream>>>(ctx->n, ctx->P, ctx->vec[1], refine_Xg_peek(ctx), ctx->b, ctx->lin, lambda, ctx->Gcur,
                                                                      ctx->Ptrial, ctx->Gtrial, ctx->part);
        ctx->launches++;
    } else
        apply_update_kernel<double><<<nbv, kThreads, 0, ctx->stream>>>(ctx->n, ctx->P, v.x, v.xg, ctx->b, ctx->lin, lambda, ctx->Gcur,
                                                                      ctx->Ptrial, ctx->Gtrial, ctx->part);
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->h_pinned + 3 * kMaxBlocks, ctx->part, sizeof(double) * nbv, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = eval_cost(ctx, W, ctx->Ptrial, ctx->Gtrial, temp, nullptr);
    if (rc) return rc;
    *scale = host_sum(ctx->h_pinned + 3 * kMaxBlocks, nbv) + 1e-3;
    return DSC_OK;
}

static double ev_ms(cudaEvent_t a, cudaEvent_t b2) { float f = 0.f; cudaEventElapsedTime(&f, a, b2); return (double)f; }

extern "C" int dsc_optimize(dsc_ctx* ctx, const dsc_weights* w, int n_iters, dsc_iter_record* records, dsc_opt_stats* stats) {
    int s = ready(ctx, w);
    if (s) return s;
    if (n_iters < 0) return fail(ctx, DSC_ERR_INVALID_ARG, "n_iters");
    CK(cudaSetDevice(ctx->device));
    dsc_opt_stats st{};
    long long launches0 = ctx->launches;
    if (ctx->n == 0) { if (stats) *stats = st; return DSC_OK; }
    WeightsDev W = make_weights(ctx, w);
    const bool dense = dense_active(ctx);
    if (dense && ctx->n > DSC_DENSE_MAX) return fail(ctx, DSC_ERR_INVALID_ARG, "dense solver: too many correspondences (DSC_DENSE_MAX)");
    cudaEvent_t e_begin = ctx->evo[0], e_end = ctx->evo[1], e_a = ctx->evo[2], e_b = ctx->evo[3];
    cudaEventRecord(e_begin, ctx->stream);
    double lambda = 0.0, ni = 2.0;
    double current = 0.0;
    bool expect_long = true;
    const bool early_log = std::getenv("DSC_EARLY_LOG") != nullptr;     // study aid: rho at every pause, predictor off
    int rc = DSC_OK;
    int nbv = grid_threads(ctx, ctx->n);
    CgVecs v = make_vecs(ctx);
    for (int it = 0; it < n_iters; ++it) {
        LinGlobal hl;
        cudaEventRecord(e_a, ctx->stream);
        rc = run_linearize(ctx, W, &hl);
        if (rc) break;
        cudaEventRecord(e_b, ctx->stream); cudaEventSynchronize(e_b);
        st.linearize_ms += ev_ms(e_a, e_b);
        current = hl.chi2[0] + hl.chi2[1] + hl.chi2[2];
        if (!std::isfinite(current)) { rc = fail(ctx, DSC_ERR_NONFINITE, "cost is not finite at linearisation"); break; }
        if (it == 0) { lambda = 1e-5 * hl.maxdiag; ni = 2.0; }          // computeLambdaInit, tau = 1e-5
        dsc_iter_record rec{};
        rec.chi2_before = current; rec.lambda = lambda;
        double rho = 0.0;
        int q = 0;
        bool accepted = false;
        if (dense) { rc = dense_prepare(ctx, W); if (rc) break; }
        do {
            int its = 0;
            cudaEventRecord(e_a, ctx->stream);
            int prc = dense ? dense_solve(ctx, lambda) : pcg_begin(ctx, W, lambda);
            cudaEventRecord(e_b, ctx->stream);
            if (dense) { cudaEventSynchronize(e_b); st.pcg_ms += ev_ms(e_a, e_b); }
            double temp = std::numeric_limits<double>::max(), scale = 1e-3;
            // Early rejection: the solve is paused at each loose tolerance, the trial is evaluated, and a step that
            // already fails the gain test by that level's margin is rejected without finishing the solve (only the
            // SIGN of rho matters for a rejected step: lambda *= ni either way).  Otherwise the same CG is resumed;
            // the last pass is always the tight tolerance, so accepted steps are computed exactly as without it.
            bool rejected_early = false;
            for (int level = 0; level <= ctx->early_levels && prc == DSC_OK && !rejected_early; ++level) {
                const bool last = level == ctx->early_levels;
                const double tol = last ? ctx->pcg.rtol : ctx->early_rtol[level];
                if (!last && (dense || !(tol > ctx->pcg.rtol) || (!expect_long && !early_log))) continue;
                if (!dense) {
                    cudaEventRecord(e_a, ctx->stream);
                    prc = pcg_resume(ctx, W, lambda, tol, &its);
                    cudaEventRecord(e_b, ctx->stream); cudaEventSynchronize(e_b);
                    st.pcg_ms += ev_ms(e_a, e_b);
                }
                if (prc != DSC_OK) break;
                cudaEventRecord(e_a, ctx->stream);
                rc = eval_trial(ctx, W, lambda, &temp, &scale);
                if (rc) break;
                cudaEventRecord(e_b, ctx->stream); cudaEventSynchronize(e_b);
                st.trial_ms += ev_ms(e_a, e_b);
                if (early_log) fprintf(stderr, "[dsc early] it %d trial %d level %d tol %.1e its %d rho %.6e\n", it, q, last ? -1 : level, tol, its, (current - temp) / scale);
                if (!last && std::isfinite(temp) && (current - temp) / scale < -ctx->early_margin[level]) {
                    rejected_early = true;
                    st.early_rejects++;
                }
            }
            if (rc) break;
            // A solve that hit the iteration limit is a failed solve, as a failed factorisation is for g2o
            // (OptimizationAlgorithmLevenberg: ok2 == false -> tempChi = max, the trial is rejected and lambda grows);
            // it is counted in dsc_opt_stats::pcg_unconverged so that it never passes silently.
            if (prc == kPcgUnconverged) { st.pcg_unconverged++; prc = DSC_ERR_PCG_BREAKDOWN; }
            if (prc != DSC_OK && prc != DSC_ERR_PCG_BREAKDOWN) { rc = prc; break; }
            if (prc == DSC_ERR_PCG_BREAKDOWN) { temp = std::numeric_limits<double>::max(); scale = 1e-3; }
            rec.pcg_iters += its; st.total_pcg_iters += its;
            // A pause costs about one PCG iteration (trial evaluation + a host round trip): worth it only while
            // solves are long.  Predictor: the previous solve that ran to the tight tolerance.
            if (!rejected_early && prc == DSC_OK) expect_long = its > kEarlyWorthIters;
            rho = (current - temp) / scale;
            if (rho > 0 && std::isfinite(temp)) {
                double alpha = 1.0 - std::pow(2.0 * rho - 1.0, 3);
                alpha = std::min(alpha, 2.0 / 3.0);
                lambda *= std::max(1.0 / 3.0, alpha);
                ni = 2.0;
                current = temp;
                std::swap(ctx->P, ctx->Ptrial);
                std::swap(ctx->Gcur, ctx->Gtrial);
                accepted = true;
            } else {
                lambda *= ni;
                ni *= 2.0;
            }
            ++q;
        } while (rho < 0 && q < 10);
        if (rc) break;
        rec.trials = q; rec.accepted = accepted ? 1 : 0; rec.chi2_after = current;
        st.total_trials += q; st.iterations = it + 1;
        if (records) records[it] = rec;
        if (q == 10 || rho == 0) { st.terminated = 1; break; }
    }
    cudaEventRecord(e_end, ctx->stream); cudaEventSynchronize(e_end);
    st.device_ms = ev_ms(e_begin, e_end);
    st.final_chi2 = current;
    st.kernel_launches = (int)(ctx->launches - launches0);
    if (stats) *stats = st;
    if (rc == DSC_OK) rc = shard_check(ctx);
    return rc;
}

// point-sharded pair: every rank's own rows of the state into every rank's copy, so that what follows sees the whole pair
static int shard_sync_state(dsc_ctx* ctx) {
    if (!ctx->sharded || !ctx->sh_attached || !ctx->have_graph || ctx->n == 0) return DSC_OK;
    shard_allgather_state_kernel<<<grid_threads(ctx, std::max(1, shard_rows(ctx))), 256, 0, ctx->stream>>>(ctx->sh, ctx->n, shard_pidx(ctx, ctx->P));
    ctx->launches++;
    shard_allreduce(ctx, SF_A, ctx->part, 1, 1, 1);          // (the value is irrelevant: its flag publishes the rows)
    CK(cudaGetLastError());
    return shard_check(ctx);
}

extern "C" int dsc_download(dsc_ctx* ctx, float* X1, float* X2, double* X1d, double* X2d,
                            double* scales, double* Tg7, double* update) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_problem) return fail(ctx, DSC_ERR_STATE, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    Globals g;
    CK(cudaMemcpyAsync(&g, ctx->Gcur, sizeof(Globals), cudaMemcpyDeviceToHost, ctx->stream));
    double upd = 0.0;
    { int src = shard_sync_state(ctx); if (src) return src; }
    if (n > 0) {
        int nb = grid_threads(ctx, n);
        // fp64 copies in the caller's order go through the CG scratch vector w ([n][6] doubles, dead between solves)
        double* dX1 = X1d ? ctx->vec[3] : nullptr;
        double* dX2 = X2d ? ctx->vec[3] + 3 * (size_t)n : nullptr;
        export_kernel<<<nb, kThreads, 0, ctx->stream>>>(n, ctx->P, ctx->P0, ctx->perm.empty() ? nullptr : ctx->d_perm,
                                                       ctx->X1f, ctx->X2f, dX1, dX2, ctx->part);
        ctx->launches++;
        CK(cudaGetLastError());
        if (X1) CK(cudaMemcpyAsync(X1, ctx->X1f, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (X2) CK(cudaMemcpyAsync(X2, ctx->X2f, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (X1d) CK(cudaMemcpyAsync(X1d, dX1, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (X2d) CK(cudaMemcpyAsync(X2d, dX2, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * nb, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        upd = host_sum(ctx->h_pinned, nb);
    } else CK(cudaStreamSynchronize(ctx->stream));
    if (scales) { scales[0] = g.s1; scales[1] = g.s2; }
    if (Tg7) for (int k = 0; k < 7; ++k) Tg7[k] = g.Tg[k];
    if (update) *update = upd;
    return DSC_OK;
}

extern "C" int dsc_pixel_sigma(dsc_ctx* ctx, double* sigma) {
    if (!ctx || !sigma) return DSC_ERR_INVALID_ARG;
    if (!ctx->have_problem) return fail(ctx, DSC_ERR_STATE, "no problem uploaded");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    if (n == 0) { sigma[0] = sigma[1] = std::numeric_limits<double>::quiet_NaN(); return DSC_OK; }
    { int src = shard_sync_state(ctx); if (src) return src; }
    int nb = grid_threads(ctx, n);
    pixel_sigma_kernel<<<nb, kThreads, 0, ctx->stream>>>(n, ctx->P, ctx->uv, ctx->pair, ctx->part);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_pinned, ctx->part, sizeof(double) * 4 * nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double s[4];
    for (int k = 0; k < 4; ++k) s[k] = host_sum(ctx->h_pinned, nb, 4, k);
    sigma[0] = 0.5 * (std::sqrt(s[0] / n) + std::sqrt(s[1] / n));
    sigma[1] = 0.5 * (std::sqrt(s[2] / n) + std::sqrt(s[3] / n));
    return DSC_OK;
}

// ------------------------------------------------------------------ test hooks
extern "C" int dsc_debug_linearize(dsc_ctx* ctx, const dsc_weights* w, double* b, double* hdiag, double* chi2) {
    int s = ready(ctx, w);
    if (s) return s;
    if (ctx->sharded) return fail(ctx, DSC_ERR_STATE, "test hook: not on a sharded context");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    WeightsDev W = make_weights(ctx, w);
    LinGlobal hl;
    s = run_linearize(ctx, W, &hl);
    if (s) return s;
    std::vector<double> hb(6 * (size_t)n), hD(21 * 32 * (((size_t)n + 31) / 32));
    if (n) {
        CK(cudaMemcpyAsync(hb.data(), ctx->b, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hD.data(), ctx->D, sizeof(double) * hD.size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (chi2) *chi2 = hl.chi2[0] + hl.chi2[1] + hl.chi2[2];
    for (int k = 0; k < 8; ++k) { if (b) b[k] = hl.bg[k]; if (hdiag) hdiag[k] = hl.C[k * 8 + k]; }
    for (int i = 0; i < n; ++i) {
        size_t d = ctx->perm.empty() ? (size_t)i : (size_t)ctx->perm[i];
        for (int k = 0; k < 6; ++k) {
            if (b) b[8 + 6 * d + k] = hb[6 * (size_t)i + k];
            if (hdiag) hdiag[8 + 6 * d + k] = hD[((size_t)(i >> 5) * 21 + pk<6>(k, k)) * 32 + (i & 31)];
        }
    }
    return DSC_OK;
}

extern "C" int dsc_debug_matvec(dsc_ctx* ctx, const dsc_weights* w, double lambda, const double* x, double* y) {
    int s = ready(ctx, w);
    if (s) return s;
    if (ctx->sharded) return fail(ctx, DSC_ERR_STATE, "test hook: not on a sharded context");
    if (!x || !y) return DSC_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    WeightsDev W = make_weights(ctx, w);
    LinGlobal hl;
    s = run_linearize(ctx, W, &hl);
    if (s) return s;
    CgVecs v = make_vecs(ctx);
    std::vector<double> hz(6 * (size_t)n);
    std::vector<float> hzf(f32(ctx) ? 6 * (size_t)n : 0);
    for (int i = 0; i < n; ++i) {
        size_t sidx = ctx->perm.empty() ? (size_t)i : (size_t)ctx->perm[i];
        for (int k = 0; k < 6; ++k) hz[6 * (size_t)i + k] = x[8 + 6 * sidx + k];
    }
    for (size_t k = 0; k < hzf.size(); ++k) hzf[k] = (float)hz[k];
    if (n && f32(ctx)) CK(cudaMemcpyAsync(ctx->vecF[2], hzf.data(), sizeof(float) * 6 * n, cudaMemcpyHostToDevice, ctx->stream));
    else if (n) CK(cudaMemcpyAsync(v.z, hz.data(), sizeof(double) * 6 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(v.zg, x, sizeof(double) * 8, cudaMemcpyHostToDevice, ctx->stream));
    int nbs = grid_spmv(ctx, n);
    if (f32(ctx)) launch_spmv<float>(ctx, W, lambda, nullptr); else launch_spmv<double>(ctx, W, lambda, nullptr);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<double> hw(6 * (size_t)n), hbp(8 * (size_t)nbs);
    if (n && f32(ctx)) CK(cudaMemcpyAsync(hzf.data(), ctx->vecF[3], sizeof(float) * 6 * n, cudaMemcpyDeviceToHost, ctx->stream));
    else if (n) CK(cudaMemcpyAsync(hw.data(), v.w, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hbp.data(), ctx->bpart, sizeof(double) * 8 * nbs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < hzf.size(); ++k) hw[k] = (double)hzf[k];
    for (int k = 0; k < 8; ++k) {
        double sum = 0.0;
        for (int bk = 0; bk < nbs; ++bk) sum += hbp[8 * (size_t)bk + k];
        y[k] = sum + ((k >= 6 ? hl.C[k * 8 + k] : 0.0) + lambda) * x[k];
    }
    for (int i = 0; i < n; ++i) {
        size_t d = ctx->perm.empty() ? (size_t)i : (size_t)ctx->perm[i];
        for (int k = 0; k < 6; ++k) y[8 + 6 * d + k] = hw[6 * (size_t)i + k];
    }
    return DSC_OK;
}

// ------------------------------------------------------------------ per-kernel timing (roofline)
extern "C" int dsc_problem_size(const dsc_ctx* ctx, long long* n, long long* n_edges) {
    if (!ctx) return DSC_ERR_INVALID_ARG;
    if (n) *n = ctx->n;
    if (n_edges) *n_edges = ctx->E;
    return DSC_OK;
}

extern "C" int dsc_profile_kernels(dsc_ctx* ctx, const dsc_weights* w, int warm, int reps, double* ms, double* bytes) {
    int s = ready(ctx, w);
    if (s) return s;
    if (!ms || reps < 1 || warm < 0) return DSC_ERR_INVALID_ARG;
    if (ctx->sharded) return fail(ctx, DSC_ERR_STATE, "per-kernel timing: not on a sharded context");
    CK(cudaSetDevice(ctx->device));
    int n = ctx->n;
    if (n == 0) return fail(ctx, DSC_ERR_STATE, "empty problem");
    WeightsDev W = make_weights(ctx, w);
    LinGlobal hl;
    s = run_linearize(ctx, W, &hl);
    if (s) return s;
    double lambda = 1e-5 * hl.maxdiag;
    CgVecs v = make_vecs(ctx);
    int nbv = grid_threads(ctx, n);
    const bool lo = f32(ctx);
    const double tb = lo ? 4.0 : 8.0;              // bytes per stored solver value
    CK(cudaMemsetAsync(ctx->errflag, 0, sizeof(int), ctx->stream));
    if (lo) launch_init<float>(ctx, lambda); else launch_init<double>(ctx, lambda);
    ctx->launches += 2;
    double N = (double)n, E = (double)ctx->E;
    double by[DSC_K_COUNT];
    double S = (double)ctx->nblk * 32.0;            // ELL slots (padding included)
    by[DSC_K_SPMV] = (32.0 + 26.0 * tb) * N + (4.0 + 9.0 * tb) * S;   // X1(32) | z(6) U(14) w(6) values | ELL blocks: col(4) + Je(9 values) per slot (padding included)
    const double mb = (lo || ctx->minv_float) ? 4.0 : 8.0;   // bytes per stored preconditioner value (float also in the fp64 mode)
    by[DSC_K_UPDATE] = (66.0 * tb + 21.0 * mb) * N;   // read z w p s x r (36 values) + Minv (21), write p s x r z (30)
    by[DSC_K_LINEARIZE] = (488.0 + (lo ? 56.0 : 0.0)) * N + 12.0 * S + (72.0 + (lo ? 36.0 : 0.0)) * E;   // P Q uv dm isg | ecol ewgt per slot | write b D U, Je per edge (+ the float copies in the fp32 mode)
    by[DSC_K_COST] = 136.0 * N + 12.0 * S;         // P Q uv dm isg | ecol ewgt per slot
    by[DSC_K_PRECOND] = (48.0 + 168.0 + 12.0 * tb + 21.0 * mb) * N;   // preconditioner + PCG start: b D -> Minv (21 values) r z (12)
    by[DSC_K_APPLY] = 224.0 * N;                  // P x b -> Ptrial
    by[DSC_K_ROTATIONS] = 96.0 * N + 12.0 * S;     // P | ecol ewgt per slot | write Q
    auto time_it = [&](int which, auto&& launch) -> int {
        for (int i = 0; i < warm; ++i) launch();
        CK(cudaEventRecord(ctx->evA, ctx->stream));
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(ctx->evB, ctx->stream));
        CK(cudaEventSynchronize(ctx->evB));
        CK(cudaGetLastError());
        ms[which] = ev_ms(ctx->evA, ctx->evB) / reps;
        ctx->launches += warm + reps;
        return DSC_OK;
    };
    s = time_it(DSC_K_SPMV, [&]() {
        if (lo) launch_spmv<float>(ctx, W, lambda, nullptr); else launch_spmv<double>(ctx, W, lambda, nullptr);
    });
    if (s) return s;
    // update: run real CG steps (first=1 keeps beta = 0, so the recurrences stay finite for any reps)
    s = time_it(DSC_K_UPDATE, [&]() {
        CgControl z{};                                 // a regular (not first) step: beta ~ 0, all seven vectors read
        z.lambda = lambda; z.gamma0 = 1e300; z.sc[1].gamma_prev = 1e300; z.sc[1].alpha_prev = 1.0;
        cudaMemcpyAsync(ctx->ctl, &z, sizeof(CgControl), cudaMemcpyHostToDevice, ctx->stream);
        if (lo) launch_update<float>(ctx, 0, 0, lambda, 0, 1, 0.0); else launch_update<double>(ctx, 0, 0, lambda, 0, 1, 0.0);
    });
    if (s) return s;
    s = time_it(DSC_K_LINEARIZE, [&]() {
        launch_linearize(ctx, W, grid_tiles(ctx, n, 1));
    });
    if (s) return s;
    s = time_it(DSC_K_COST, [&]() {
        cost_ell_kernel<<<grid_tiles(ctx, n, DSC_COST_BLOCKS), kEllThreads, kWinBytes, ctx->stream>>>(n, ctx->P, ctx->Q, ctx->uv, ctx->dm, ctx->isg, ctx->sliceptr, ctx->ecol,
                                                                                  ctx->ewgt, ctx->Gcur, ctx->pair, W, ctx->part, 0, -1);
    });
    if (s) return s;
    s = time_it(DSC_K_PRECOND, [&]() {
        if (lo) launch_init<float>(ctx, lambda); else launch_init<double>(ctx, lambda);
    });
    if (s) return s;
    s = time_it(DSC_K_APPLY, [&]() {
        apply_update_kernel<double><<<nbv, kThreads, 0, ctx->stream>>>(n, ctx->P, v.x, v.xg, ctx->b, ctx->lin, lambda, ctx->Gcur,
                                                                      ctx->Ptrial, ctx->Gtrial, ctx->gpart[0]);
    });
    if (s) return s;
    double* Qtmp = ctx->vec[3];                    // scratch: do not disturb the real rotations
    s = time_it(DSC_K_ROTATIONS, [&]() {
        rotations_ell_kernel<<<grid_tiles(ctx, n, DSC_ROT_BLOCKS), kEllThreads, kWinBytes, ctx->stream>>>(n, ctx->P, ctx->sliceptr, ctx->ecol, ctx->ewgt, Qtmp);
    });
    if (s) return s;
    if (bytes) for (int k = 0; k < DSC_K_COUNT; ++k) bytes[k] = by[k];
    return DSC_OK;
}

extern "C" int dsc_profile_triangulate(dsc_ctx* ctx, const dsc_tri_params* prm, int warm, int reps, double* ms, double* bytes) {
    if (!ctx || !prm || !ms || reps < 1 || warm < 0) return DSC_ERR_INVALID_ARG;
    if (ctx->tn == 0) return fail(ctx, DSC_ERR_STATE, "dsc_tri_upload first");
    CK(cudaSetDevice(ctx->device));
    TriParams tp{prm->method, prm->location, prm->gate, prm->min_cos, prm->depth_limit, prm->check_reproj};
    int nb = grid_threads(ctx, ctx->tn);
    auto launch = [&]() {
        triangulate_kernel<<<nb, kThreads, 0, ctx->stream>>>(ctx->tn, ctx->t_uv1, ctx->t_uv2, ctx->t_d1, ctx->t_d2, ctx->tpair, tp,
                                                            ctx->t_X1, ctx->t_X2, ctx->t_valid, ctx->t_cos);
    };
    for (int i = 0; i < warm; ++i) launch();
    CK(cudaEventRecord(ctx->evA, ctx->stream));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(ctx->evB, ctx->stream));
    CK(cudaEventSynchronize(ctx->evB));
    CK(cudaGetLastError());
    ctx->launches += warm + reps;
    ctx->t_done = true;
    *ms = ev_ms(ctx->evA, ctx->evB) / reps;
    if (bytes) *bytes = (double)ctx->tn * (16.0 + (prm->method == DSC_TRI_DEPTH ? 8.0 : 0.0) + 24.0 + 1.0 + 4.0);
    return DSC_OK;
}

// ------------------------------------------------------------------ kNN graph on the GPU
extern "C" int dsc_knn_build(dsc_ctx* ctx, int n, const float* X, int k, long long* n_edges) {
    if (!ctx || n < 0 || (n > 0 && !X) || k < 1 || k > kKnnMax) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_knn_build");
    CK(cudaSetDevice(ctx->device));
    ctx->knn_n = n; ctx->knn_E = 0;
    if (n_edges) *n_edges = 0;
    if (!ctx->knn_rowptr || (size_t)n > ctx->knn_cap_n) {
        const size_t cap = (size_t)n + (size_t)n / 8 + 16;
        ctx->knn_cap_n = 0;
        CK(dev_alloc(ctx->knn_rowptr, cap + 1));
        ctx->knn_cap_n = cap;
    }
    if (n == 0) { CK(cudaMemset(ctx->knn_rowptr, 0, sizeof(int))); return DSC_OK; }
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    int bad = 0;
#pragma omp parallel for reduction(min : x0, y0) reduction(max : x1, y1) reduction(| : bad) schedule(static)
    for (int i = 0; i < n; ++i) {
        double x = X[3 * (size_t)i], y = X[3 * (size_t)i + 1];
        if (!std::isfinite(x) || !std::isfinite(y)) { bad |= 1; continue; }
        x0 = std::min(x0, x); x1 = std::max(x1, x); y0 = std::min(y0, y); y1 = std::max(y1, y);
    }
    if (bad) return fail(ctx, DSC_ERR_INVALID_ARG, "dsc_knn_build: non-finite coordinate");
    double w = x1 - x0, h = y1 - y0;
    KnnGrid g{x0, y0, 1.0, 1, 1};
    if (w > 0 || h > 0) {
        double cell = std::sqrt(std::max(w, 1e-300) * std::max(h, 1e-300) / std::max(1.0, n / 2.0));
        cell = std::max(cell, std::max(w, h) / 4096.0);
        g.inv_cell = 1.0 / cell;
        g.nx = std::max(1, (int)std::ceil(w / cell) + 1);
        g.ny = std::max(1, (int)std::ceil(h / cell) + 1);
    }
    int ncells = g.nx * g.ny;
    auto release = [&]() {};
    int m = std::max(ncells, n) + 1;
    const size_t nsums = (size_t)(m / kScanBlock + 2);
    {
        size_t need = ws_round(sizeof(float) * 3 * (size_t)n);
        for (size_t c : {(size_t)n, (size_t)m, (size_t)m, (size_t)m, (size_t)n, (size_t)n * k, nsums, (size_t)n + 1}) need += ws_round(c * sizeof(int));
        cudaError_t e = ws_begin(ctx, need);
        if (e != cudaSuccess) return fail(ctx, DSC_ERR_ALLOC, cudaGetErrorString(e));
    }
    float* dX = ws_take<float>(ctx, 3 * (size_t)n);
    int *cell = ws_take<int>(ctx, n), *cnt = ws_take<int>(ctx, m), *start = ws_take<int>(ctx, m), *cursor = ws_take<int>(ctx, m), *order = ws_take<int>(ctx, n),
        *nbr = ws_take<int>(ctx, (size_t)n * k), *sums = ws_take<int>(ctx, nsums), *deg = ws_take<int>(ctx, (size_t)n + 1);
    auto scan = [&](int m, const int* in, int* out) -> int {      // exclusive scan of m ints
        int nb = (m + kScanBlock - 1) / kScanBlock;
        scan_block_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(m, in, out, sums);
        scan_sums_kernel<<<1, kScanBlock, 0, ctx->stream>>>(nb, sums);
        scan_add_kernel<<<nb, kScanBlock, 0, ctx->stream>>>(m, out, sums);
        ctx->launches += 3;
        return DSC_OK;
    };
    cudaError_t e = cudaSuccess;
    int nbt = grid_threads(ctx, n);
    { int rc = h2d(ctx, dX, X, sizeof(float) * 3 * (size_t)n); if (rc) return rc; }
    CK(cudaMemsetAsync(cnt, 0, sizeof(int) * m, ctx->stream));
    CK(cudaMemsetAsync(cursor, 0, sizeof(int) * m, ctx->stream));
    knn_cell_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, dX, g, cell, cnt);
    scan(ncells, cnt, start);
    knn_scatter_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, cell, start, cursor, order);
    knn_search_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, k, dX, g, start, order, ncells, nbr);
    knn_owncount_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, k, nbr, deg);
    knn_extra_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, k, nbr, deg);
    CK(cudaMemsetAsync(deg + n, 0, sizeof(int), ctx->stream));
    scan(n + 1, deg, ctx->knn_rowptr);
    ctx->launches += 5;
    int E = 0;
    CK(cudaMemcpyAsync(&E, ctx->knn_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->knn_col || (size_t)E > ctx->knn_cap_e) {
        const size_t cap = (size_t)E + (size_t)E / 8 + 16;
        ctx->knn_cap_e = 0;
        if ((e = dev_alloc(ctx->knn_col, cap))) { release(); return fail(ctx, DSC_ERR_ALLOC, cudaGetErrorString(e)); }
        ctx->knn_cap_e = cap;
    }
    CK(cudaMemsetAsync(cursor, 0, sizeof(int) * m, ctx->stream));
    knn_fill_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, k, nbr, ctx->knn_rowptr, cursor, ctx->knn_col);
    knn_sortrows_kernel<<<nbt, kThreads, 0, ctx->stream>>>(n, ctx->knn_rowptr, ctx->knn_col);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    release();
    ctx->knn_E = E;
    if (n_edges) *n_edges = E;
    return DSC_OK;
}

extern "C" int dsc_knn_download(dsc_ctx* ctx, int32_t* rowptr, int32_t* col) {
    if (!ctx || !rowptr) return DSC_ERR_INVALID_ARG;
    if (!ctx->knn_rowptr) return fail(ctx, DSC_ERR_STATE, "dsc_knn_download before dsc_knn_build");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(rowptr, ctx->knn_rowptr, sizeof(int) * ((size_t)ctx->knn_n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (col && ctx->knn_E > 0) CK(cudaMemcpyAsync(col, ctx->knn_col, sizeof(int) * (size_t)ctx->knn_E, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return DSC_OK;
}

// ------------------------------------------------------------------ batched frame pairs (include/dsc.h: dsc_batch_*)
// One sub-context per pair owns that pair's device buffers (set-up reuses dsc_problem_upload / dsc_set_graph /
// dsc_compute_rotations); the refinement of ALL pairs is one launch of lm_batch_kernel (dsc_batch.cuh).
struct dsc_batch {
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    std::vector<dsc_ctx*> ctxs;              // grown on demand, reused between uploads
    int n_problems = 0;
    int n_active = -1;                       // dsc_batch_set_active: the launch refines pairs [0, n_active) (-1: all)
    std::vector<long long> point_offset;
    BatchProblem* d_probs = nullptr; BatchIterRec* d_recs = nullptr; BatchResult* d_res = nullptr; int* d_queue = nullptr;
    double* d_part = nullptr;                // per problem: [kBatchPartStride] linearisation / trial partials
    size_t probs_cap = 0, recs_cap = 0, part_cap = 0;
    int cluster = kSmallCluster, n_clusters = 0;
    int last_cluster = 0;                    // CTAs per cluster of the last launch (few pairs get larger clusters)
    bool cluster_forced = false;             // DSC_BATCH_CLUSTER
    dsc_pcg_params pcg{1e-10, 4000, 32};
    int early_levels = 0;
    double early_rtol[4] = {0, 0, 0, 0}, early_margin[4] = {0, 0, 0, 0};
    long long launches = 0;
};

namespace {
constexpr int kBatchPartStride = kSmallCluster * (kLinPart + 4);
int bfail(dsc_batch* b, int code, const std::string& what) {
    if (b) b->err = std::string(status_str(code)) + ": " + what;
    return code;
}
int batch_active(const dsc_batch* b) { return b->n_active >= 0 ? std::min(b->n_active, b->n_problems) : b->n_problems; }
#define BCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return bfail(bt, DSC_ERR_CUDA, std::string(#call) + " -> " + cudaGetErrorString(e_)); \
    } while (0)
}  // namespace

extern "C" int dsc_batch_create(int device, dsc_batch** out) {
    if (!out) return DSC_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DSC_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return DSC_ERR_INVALID_ARG;
    dsc_batch* bt = new dsc_batch();
    bt->device = device;
    auto bail = [&](int code) { dsc_batch_destroy(bt); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(DSC_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(DSC_ERR_CUDA);
    bt->sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&bt->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(DSC_ERR_CUDA);
    if (cudaEventCreate(&bt->ev0) != cudaSuccess || cudaEventCreate(&bt->ev1) != cudaSuccess) return bail(DSC_ERR_CUDA);
    if (cudaFuncSetAttribute(lm_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinBytes) != cudaSuccess) return bail(DSC_ERR_CUDA);
    // CTAs per cluster.  Measured on B200 (64 pairs x 10k correspondences, profiles/r02_batch_cluster_sweep.txt): the
    // rate per SM is the same for 4, 8 and 16 CTAs (the phases are bound by load latency at 8 warps per SM, not by the
    // barriers), so the size that tiles the 148 SMs best wins: 4 (33 clusters = 132 SMs; 16 CTAs fit 7 times = 112 SMs).
    // DSC_BATCH_CLUSTER overrides (1, 2, 4, 8, 16; 16 is a non-portable size).
    cudaFuncSetAttribute(lm_batch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    bt->cluster = 4;
    if (const char* cs = std::getenv("DSC_BATCH_CLUSTER")) {
        int v = std::atoi(cs);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) { bt->cluster = v; bt->cluster_forced = true; }
    }
    *out = bt;
    return DSC_OK;
}

extern "C" void dsc_batch_destroy(dsc_batch* bt) {
    if (!bt) return;
    cudaSetDevice(bt->device);
    if (bt->stream) cudaStreamSynchronize(bt->stream);
    for (dsc_ctx* c : bt->ctxs) dsc_destroy(c);
    dev_free(bt->d_probs); dev_free(bt->d_recs); dev_free(bt->d_res); dev_free(bt->d_queue); dev_free(bt->d_part);
    if (bt->ev0) cudaEventDestroy(bt->ev0);
    if (bt->ev1) cudaEventDestroy(bt->ev1);
    if (bt->stream) cudaStreamDestroy(bt->stream);
    delete bt;
}

extern "C" const char* dsc_batch_last_error(const dsc_batch* bt) { return bt ? bt->err.c_str() : "null batch"; }

extern "C" int dsc_batch_size(const dsc_batch* bt, int* n_problems, long long* n_points, int* cluster_ctas, int* clusters) {
    if (!bt) return DSC_ERR_INVALID_ARG;
    if (n_problems) *n_problems = bt->n_problems;
    if (n_points) *n_points = bt->point_offset.empty() ? 0 : bt->point_offset.back();
    if (cluster_ctas) *cluster_ctas = bt->last_cluster ? bt->last_cluster : bt->cluster;
    if (clusters) *clusters = bt->n_clusters;
    return DSC_OK;
}

extern "C" int dsc_batch_set_pcg(dsc_batch* bt, const dsc_pcg_params* prm) {
    if (!bt || !prm || !(prm->rtol > 0.0) || prm->max_iters < 1) return bfail(bt, DSC_ERR_INVALID_ARG, "dsc_batch_set_pcg");
    bt->pcg = *prm;
    return DSC_OK;
}

extern "C" int dsc_batch_set_early_reject(dsc_batch* bt, int n_levels, const double* rtol_loose, const double* rho_margin) {
    if (!bt || n_levels < 0 || n_levels > 4 || (n_levels > 0 && (!rtol_loose || !rho_margin)))
        return bfail(bt, DSC_ERR_INVALID_ARG, "dsc_batch_set_early_reject");
    for (int l = 0; l < n_levels; ++l)
        if (!(rtol_loose[l] > 0.0) || !(rho_margin[l] >= 0.0) || (l > 0 && !(rtol_loose[l] < rtol_loose[l - 1])))
            return bfail(bt, DSC_ERR_INVALID_ARG, "tolerances must be positive and strictly decreasing, margins >= 0");
    bt->early_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) { bt->early_rtol[l] = rtol_loose[l]; bt->early_margin[l] = rho_margin[l]; }
    return DSC_OK;
}

extern "C" int dsc_batch_upload(dsc_batch* bt, int n_problems, const dsc_batch_pair* pairs, const long long* point_offset,
                                const float* X1, const float* X2, const float* uv1, const float* uv2,
                                const double* depth1, const double* depth2, const float* inv_sigma2_1, const float* inv_sigma2_2,
                                const long long* edge_offset, const int32_t* rowptr, const int32_t* col, const double* w, int reorder) {
    if (!bt || n_problems < 0 || (n_problems > 0 && (!pairs || !point_offset || !edge_offset || !rowptr)))
        return bfail(bt, DSC_ERR_INVALID_ARG, "dsc_batch_upload");
    BCK(cudaSetDevice(bt->device));
    for (int p = 0; p < n_problems; ++p)
        if (point_offset[p + 1] < point_offset[p] || edge_offset[p + 1] < edge_offset[p] || point_offset[p + 1] - point_offset[p] > 0x7fffffffLL)
            return bfail(bt, DSC_ERR_INVALID_ARG, "offsets must be non-decreasing prefix sums");
    while ((int)bt->ctxs.size() < n_problems) {
        dsc_ctx* c = nullptr;
        int st = dsc_create(bt->device, &c);
        if (st) return bfail(bt, st, "sub-context");
        bt->ctxs.push_back(c);
    }
    bt->n_problems = 0;
    if (n_problems == 0) bt->point_offset.assign(1, 0LL);
    else bt->point_offset.assign(point_offset, point_offset + n_problems + 1);
    for (int p = 0; p < n_problems; ++p) {
        dsc_ctx* c = bt->ctxs[p];
        const long long o = point_offset[p], e = edge_offset[p];
        const int n = (int)(point_offset[p + 1] - o);
        const dsc_batch_pair& bp = pairs[p];
        const bool ident = bp.Tg7[0] == 0 && bp.Tg7[1] == 0 && bp.Tg7[2] == 0 && bp.Tg7[3] == 0;
        int st = dsc_problem_upload(c, &bp.pair, n, X1 + 3 * o, X2 + 3 * o, uv1 + 2 * o, uv2 + 2 * o, depth1 + o, depth2 + o,
                                    inv_sigma2_1 ? inv_sigma2_1 + o : nullptr, inv_sigma2_2 ? inv_sigma2_2 + o : nullptr,
                                    bp.scale1, bp.scale2, ident ? nullptr : bp.Tg7);
        if (!st) st = dsc_set_graph(c, n, rowptr + o + p, col + e, w + e, bp.area, bp.n_triangles, reorder);
        if (!st) st = dsc_compute_rotations(c);
        if (st) return bfail(bt, st, "problem " + std::to_string(p) + ": " + c->err);
    }
    for (int p = 0; p < n_problems; ++p) BCK(cudaStreamSynchronize(bt->ctxs[p]->stream));
    bt->n_problems = n_problems;
    bt->n_active = -1;
    return DSC_OK;
}

extern "C" int dsc_batch_set_active(dsc_batch* bt, int n_active) {
    if (!bt) return DSC_ERR_INVALID_ARG;
    bt->n_active = n_active;
    return DSC_OK;
}

// calculatePixelsStandDev of every active pair (the weight search's objective): sigma[p][2]
extern "C" int dsc_batch_pixel_sigma(dsc_batch* bt, double* sigma) {
    if (!bt || !sigma) return DSC_ERR_INVALID_ARG;
    const int np = batch_active(bt);
    for (int p = 0; p < np; ++p) {
        int st = dsc_pixel_sigma(bt->ctxs[p], sigma + 2 * (size_t)p);
        if (st) return bfail(bt, st, bt->ctxs[p]->err);
    }
    return DSC_OK;
}

extern "C" int dsc_batch_reset_state(dsc_batch* bt) {
    if (!bt) return DSC_ERR_INVALID_ARG;
    for (int p = 0; p < batch_active(bt); ++p) {
        int st = dsc_reset_state(bt->ctxs[p]);
        if (st) return bfail(bt, st, bt->ctxs[p]->err);
    }
    return DSC_OK;
}

extern "C" int dsc_batch_optimize(dsc_batch* bt, const dsc_weights* weights, int n_weights, int n_iters, dsc_iter_record* records,
                                  dsc_opt_stats* stats, double* device_ms) {
    if (!bt) return DSC_ERR_INVALID_ARG;
    const int np = batch_active(bt);
    if (!weights || n_iters < 0 || (n_weights != 1 && n_weights != np && n_weights != bt->n_problems))
        return bfail(bt, DSC_ERR_INVALID_ARG, "dsc_batch_optimize: one dsc_weights for all pairs or one per (active) pair");
    BCK(cudaSetDevice(bt->device));
    if (device_ms) *device_ms = 0.0;
    if (np == 0) return DSC_OK;
    if ((size_t)np > bt->probs_cap) {
        BCK(dev_alloc(bt->d_probs, (size_t)np)); BCK(dev_alloc(bt->d_res, (size_t)np));
        BCK(dev_alloc(bt->d_queue, (size_t)2 + 4096));
        bt->probs_cap = (size_t)np;
    }
    if ((size_t)np * kBatchPartStride > bt->part_cap) { BCK(dev_alloc(bt->d_part, (size_t)np * kBatchPartStride)); bt->part_cap = (size_t)np * kBatchPartStride; }
    const size_t nrec = (size_t)np * (size_t)std::max(1, n_iters);
    if (nrec > bt->recs_cap) { BCK(dev_alloc(bt->d_recs, nrec)); bt->recs_cap = nrec; }
    std::vector<BatchProblem> hp((size_t)np);
    for (int p = 0; p < np; ++p) {
        dsc_ctx* c = bt->ctxs[p];
        const dsc_weights* w = weights + (n_weights == 1 ? 0 : p);
        int st = ready(c, w);
        if (st) return bfail(bt, st, "problem " + std::to_string(p) + ": " + c->err);
        BatchProblem& B = hp[p];
        B.n = c->n;
        B.P = c->P; B.Ptrial = c->Ptrial; B.Q = c->Q;
        B.uv = c->uv; B.dm = c->dm; B.isg = c->isg;
        B.sliceptr = c->sliceptr; B.ecol = c->ecol; B.ewgt = c->ewgt;
        B.Je = c->Je; B.U = c->U; B.D = c->D; B.Minv = c->Minv; B.b = c->b;
        B.v = make_vecs(c); B.Ginv = c->small + 48;
        B.G = c->Gcur; B.lin = c->lin; B.err = c->errflag;
        B.part = bt->d_part + (size_t)p * kBatchPartStride;
        B.gpart0 = c->gpart[0]; B.gpart1 = c->gpart[1]; B.dpart = c->dpart; B.bpart = c->bpart;
        B.pair = c->pair; B.W = make_weights(c, w);
    }
    BCK(cudaMemcpyAsync(bt->d_probs, hp.data(), sizeof(BatchProblem) * (size_t)np, cudaMemcpyHostToDevice, bt->stream));
    BCK(cudaMemsetAsync(bt->d_recs, 0, sizeof(BatchIterRec) * nrec, bt->stream));
    BCK(cudaMemsetAsync(bt->d_res, 0, sizeof(BatchResult) * (size_t)np, bt->stream));
    BCK(cudaMemsetAsync(bt->d_queue, 0, sizeof(int) * (2 + 4096), bt->stream));
    BatchParams prm{};
    prm.n_iters = n_iters; prm.max_pcg = bt->pcg.max_iters; prm.rtol = bt->pcg.rtol; prm.early_levels = bt->early_levels;
    for (int l = 0; l < 4; ++l) { prm.early_rtol[l] = bt->early_rtol[l]; prm.early_margin[l] = bt->early_margin[l]; }
    // persistent clusters: as many as the device keeps resident at once (one CTA per SM: 250 registers x 256 threads), at
    // most one per pair; the queue balances the load
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[1];
    // A few pairs (the weight search: up to four replicas) would leave most SMs idle with 4 CTAs each: larger clusters then
    int cl = bt->cluster;
    if (!bt->cluster_forced) while (cl < kSmallCluster && np * cl * 2 <= bt->sms) cl *= 2;
    for (;;) {
        cfg = cudaLaunchConfig_t{};
        cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kWinBytes; cfg.stream = bt->stream;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cfg.gridDim = dim3(cl);
        int maxc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, lm_batch_kernel, &cfg);
        if (e == cudaSuccess && maxc > 0) { bt->n_clusters = std::min(std::min(maxc, np), 4096); break; }
        cudaGetLastError();
        if (cl > 1) { cl /= 2; bt->cluster = std::min(bt->cluster, cl); continue; }
        return bfail(bt, DSC_ERR_CUDA, std::string("no cluster configuration of lm_batch_kernel is schedulable: ") + cudaGetErrorString(e));
    }
    bt->last_cluster = cl;
    cfg.gridDim = dim3(bt->n_clusters * cl);
    BCK(cudaEventRecord(bt->ev0, bt->stream));
    BCK(cudaLaunchKernelEx(&cfg, lm_batch_kernel, (const BatchProblem*)bt->d_probs, np, prm, bt->d_queue, bt->d_recs, bt->d_res));
    BCK(cudaEventRecord(bt->ev1, bt->stream));
    bt->launches++;
    std::vector<BatchResult> hr((size_t)np);
    BCK(cudaMemcpyAsync(hr.data(), bt->d_res, sizeof(BatchResult) * (size_t)np, cudaMemcpyDeviceToHost, bt->stream));
    if (records && n_iters > 0) {
        static_assert(sizeof(BatchIterRec) == sizeof(dsc_iter_record), "record layouts must agree");
        BCK(cudaMemcpyAsync(records, bt->d_recs, sizeof(BatchIterRec) * nrec, cudaMemcpyDeviceToHost, bt->stream));
    }
    BCK(cudaStreamSynchronize(bt->stream));
    float ms = 0.f;
    BCK(cudaEventElapsedTime(&ms, bt->ev0, bt->ev1));
    if (device_ms) *device_ms = (double)ms;
    int rc = DSC_OK;
    for (int p = 0; p < np; ++p) {
        const BatchResult& r = hr[p];
        if (stats) {
            dsc_opt_stats st{};
            st.iterations = r.iterations; st.total_trials = r.total_trials; st.total_pcg_iters = r.total_pcg_iters;
            st.terminated = r.terminated; st.final_chi2 = r.final_chi2; st.early_rejects = r.early_rejects;
            st.pcg_unconverged = r.pcg_unconverged; st.device_ms = (double)ms; st.kernel_launches = p == 0 ? 1 : 0;
            stats[p] = st;
        }
        if (r.status != 0 && rc == DSC_OK) rc = bfail(bt, r.status, "problem " + std::to_string(p) + ": cost is not finite");
    }
    return rc;
}

extern "C" int dsc_batch_download(dsc_batch* bt, float* X1, float* X2, double* scales, double* Tg7, double* update) {
    if (!bt) return DSC_ERR_INVALID_ARG;
    for (int p = 0; p < bt->n_problems; ++p) {
        const long long o = bt->point_offset[p];
        int st = dsc_download(bt->ctxs[p], X1 ? X1 + 3 * o : nullptr, X2 ? X2 + 3 * o : nullptr, nullptr, nullptr,
                              scales ? scales + 2 * (size_t)p : nullptr, Tg7 ? Tg7 + 7 * (size_t)p : nullptr, update ? update + p : nullptr);
        if (st) return bfail(bt, st, bt->ctxs[p]->err);
    }
    return DSC_OK;
}

// ------------------------------------------------------------------ classic bundle adjustment (include/dsc.h: dsc_ba_*)
#include "dsc_ba_api.cuh"
