/*
 * dsc.h -- C ABI of the B200-native deformable two-view hot path.
 *
 * One library (libdsc_b200.so, hand-written sm_100a CUDA) behind plain-C entry
 * points.  Every entry point names the reference code it replaces (paths are
 * relative to the reference repository, luicalrob/Triangulation-in-Deformable-Scenes):
 *
 *   dsc_triangulate*        Modules/Utils/Geometry.cc:216-230 useTriangulationMethod and the four
 *                           triangulators (:62-214), called per match from
 *                           Modules/Mapping/Mapping.cc:294-343 (+ gate :351-364) and
 *                           Modules/Mapping/MonocularMapInitializer.cc:303-368 (+ gates :315-360);
 *                           KannalaBrandt8::unproject/project Modules/Calibration/KannalaBrandt8.cc:32-83
 *   dsc_triangulate_rays    the same dispatcher with the reference's own argument list (rays, not pixels)
 *   dsc_depth_scale_init    Modules/Map/KeyFrame.cc:131-153 setInitialDepthScaleInSimulationImages
 *   dsc_problem_upload      the Map -> g2o graph gather of arapOptimization,
 *                           Modules/Optimization/g2oBundleAdjustment.cc:640-957
 *   dsc_set_graph           mesh adjacency + cot weights + area + triangle count,
 *                           g2oBundleAdjustment.cc:657-662,883-948 (Utils/Geometry.cc:272-368)
 *   dsc_compute_rotations   Modules/Utils/Geometry.cc:549-604 computeR (called :687-688)
 *   dsc_cost                g2o computeActiveErrors + activeRobustChi2 over the edges of
 *                           Modules/Optimization/g2oTypes.h:267-349,390-421
 *   dsc_optimize            optimizer.optimize(nOptIterations), g2oBundleAdjustment.cc:959-962
 *                           (g2o Levenberg-Marquardt + BlockSolverX + LinearSolverEigen, :619-628)
 *   dsc_download            the write-back :967-1007 (points cast to float, depth scales, T_global,
 *                           optimizationUpdate)
 *   dsc_pixel_sigma         Modules/Utils/Geometry.cc:370-498 calculatePixelsStandDev
 *
 * Conventions: every function returns an int status (DSC_OK == 0, negative = error, see
 * dsc_status); nothing throws across the boundary; all pointers are HOST pointers unless the
 * name says _dev; the caller keeps ownership of everything it passes in.  One context owns one
 * CUDA stream and must be driven by one host thread at a time; several contexts may run
 * concurrently.  There is NO CPU fallback: without a CUDA device dsc_create fails.
 */
#ifndef DSC_H_
#define DSC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsc_ctx dsc_ctx;

typedef enum {
    DSC_OK = 0,
    DSC_ERR_INVALID_ARG = -1,
    DSC_ERR_CUDA = -2,
    DSC_ERR_NO_DEVICE = -3,
    DSC_ERR_STATE = -4,          /* call order violated (e.g. optimize before set_graph)   */
    DSC_ERR_NONFINITE = -5,      /* cost became NaN/inf                                      */
    DSC_ERR_PCG_BREAKDOWN = -6,  /* p.Ap <= 0 or non-finite inside the linear solve          */
    DSC_ERR_GRAPH = -7,          /* neighbour graph is not symmetric / has bad indices       */
    DSC_ERR_ALLOC = -8,
    DSC_ERR_SHARD = -9           /* point-sharded pair: a peer rank did not answer in time   */
} dsc_status;

/* camera models: Modules/Calibration/{KannalaBrandt8,PinHole}.cc */
enum { DSC_CAM_KB8 = 0, DSC_CAM_PINHOLE = 1 };
/* Triangulation.method (Geometry.cc:220-228): unknown strings select NRSLAM */
enum { DSC_TRI_CLASSIC = 0, DSC_TRI_NRSLAM = 1, DSC_TRI_ORBSLAM = 2, DSC_TRI_DEPTH = 3 };
/* Triangulation.seed.location */
enum { DSC_LOC_INRAYS = 0, DSC_LOC_TWOPOINTS = 1, DSC_LOC_FARPOINTS = 2 };
/* validity gates */
enum { DSC_GATE_NONE = 0, DSC_GATE_SIM = 1 /* Mapping.cc:351-364 */, DSC_GATE_REAL = 2 /* MonocularMapInitializer.cc:315-360 */ };

typedef struct {
    int   model;        /* DSC_CAM_*                                  */
    float params[8];    /* fx fy cx cy k0 k1 k2 k3 (Settings.cc:38-50) */
} dsc_camera;

/* A key-frame pair.  Tcw are row-major 3x4 [R|t] float matrices (Sophus::SE3f KeyFrame::getPose()). */
typedef struct {
    dsc_camera cam1, cam2;
    float T1w[12];
    float T2w[12];
} dsc_pair;

typedef struct {
    int   method;          /* DSC_TRI_*                                               */
    int   location;        /* DSC_LOC_*                                               */
    int   gate;            /* DSC_GATE_*                                              */
    float min_cos;         /* Triangulation.minCos                                    */
    float depth_limit;     /* Triangulation.depthLimit (GATE_REAL)                    */
    int   check_reproj;    /* Triangulation.checks: reprojection^2 <= 5.991 (GATE_REAL) */
} dsc_tri_params;

/* the balance weights of arapOptimization(Map*, rep, global, arap, alpha, beta, DepthError, nIter, update) */
typedef struct {
    double rep;
    double global;         /* accepted, unused -- as in the reference body   */
    double arap;
    double alpha;          /* stored on the edge, unused (g2oTypes.h:346-347) */
    double beta;
    float  depth_sigma;    /* DepthError (metres); information = 1/sigma^2    */
} dsc_weights;

typedef struct {
    double rtol;           /* stop when sqrt(r.z / r0.z0) <= rtol  (default 1e-10)      */
    int    max_iters;      /* per linear solve                      (default 4000)       */
    int    check_every;    /* host polls convergence every this many iterations (32)    */
} dsc_pcg_params;

/* per LM iteration record (one per outer g2o iteration) */
typedef struct {
    double chi2_before;    /* activeRobustChi2 at the start of the iteration */
    double chi2_after;     /* after the accepted step (== before if none)    */
    double lambda;         /* lambda at the start of the iteration           */
    int    trials;         /* lambda trials                                   */
    int    accepted;
    int    pcg_iters;      /* summed over the trials                          */
} dsc_iter_record;

typedef struct {
    int    iterations;         /* LM iterations executed                       */
    int    total_trials;
    int    total_pcg_iters;
    int    terminated;         /* 1 if g2o's Terminate condition fired         */
    double final_chi2;
    double device_ms;          /* CUDA-event time of the whole call            */
    double linearize_ms, pcg_ms, trial_ms;   /* CUDA-event breakdown           */
    int    kernel_launches;
    int    early_rejects;      /* trials rejected at the loose tolerance (dsc_set_early_reject) */
    int    pcg_unconverged;    /* linear solves that hit max_iters; each is treated as a failed solve (trial rejected,
                                  lambda grows), as g2o treats a failed factorisation                                */
} dsc_opt_stats;

/* ---- context --------------------------------------------------------------------------- */
int  dsc_create(int device, dsc_ctx** out);
void dsc_destroy(dsc_ctx* ctx);
const char* dsc_last_error(const dsc_ctx* ctx);
const char* dsc_status_string(int status);
int  dsc_version(void);
int  dsc_synchronize(dsc_ctx* ctx);
/* CUDA-event stopwatch on the context's stream (bench.py times kernels with it) */
int  dsc_timer_start(dsc_ctx* ctx);
int  dsc_timer_stop(dsc_ctx* ctx, double* ms);
/* Page-lock caller memory (cudaHostRegister) / undo it.  The upload and download entry points move pinned or registered
 * caller memory by DMA straight from / to where it lies; pageable memory is staged through a pinned bounce buffer.  No
 * context needed; registering an already registered range is not an error. */
int  dsc_pin_host(const void* ptr, size_t bytes);
int  dsc_unpin_host(const void* ptr);
/* number of kernels this context has launched so far */
int  dsc_launch_count(const dsc_ctx* ctx, long long* count);

/* ---- K1: batched two-view triangulation -------------------------------------------------
 * Host arrays of n matches: uv1/uv2 [n][2] pixels, depth1/depth2 [n] (may be NULL unless
 * method == DSC_TRI_DEPTH).  Outputs X1/X2 [n][3] world points, valid[n], cos_parallax[n]
 * (any output may be NULL).  dsc_triangulate = upload + kernel + download (end-to-end);
 * the three-step form keeps data resident for repeated kernel timing. */
int dsc_triangulate(dsc_ctx* ctx, const dsc_pair* pair, const dsc_tri_params* prm, int n,
                    const float* uv1, const float* uv2, const float* depth1, const float* depth2,
                    float* X1, float* X2, uint8_t* valid, float* cos_parallax, int* n_valid);
int dsc_tri_upload(dsc_ctx* ctx, const dsc_pair* pair, int n, const float* uv1, const float* uv2,
                   const float* depth1, const float* depth2);
int dsc_tri_run(dsc_ctx* ctx, const dsc_tri_params* prm);
int dsc_tri_download(dsc_ctx* ctx, float* X1, float* X2, uint8_t* valid, float* cos_parallax, int* n_valid);
/* useTriangulationMethod as the reference declares it (Modules/Utils/Geometry.h:66-69): bearing rays in, no camera
 * model, no gates (Geometry.cc:216-230 always returns true).  T1w/T2w row-major 3x4 [R|t]; xn1/xn2 [n][3] rays (for
 * DSC_TRI_DEPTH: camera-frame points at the measured depth, Geometry.cc:189-214); X1/X2 [n][3] world points out. */
int dsc_triangulate_rays(dsc_ctx* ctx, const float* T1w, const float* T2w, int method, int location, int n,
                         const float* xn1, const float* xn2, float* X1, float* X2);
/* mean of depth/z_c over valid points with non-zero depth (KeyFrame.cc:131-153); which = 1|2 */
int dsc_depth_scale_init(dsc_ctx* ctx, int which, double* scale);

/* ---- the refinement problem ---------------------------------------------------------------
 * n compact correspondences (every one has both map points and both observations):
 * X1/X2 [n][3] float map-point positions, uv1/uv2 [n][2], depth1/depth2 [n] (metres, as returned
 * by KeyFrame::getDepthMeasure(..., false)), inv_sigma2_1/2 [n] (KeyFrame::getInvSigma2(octave),
 * NULL = 1), initial depth scales, T_global as (qx,qy,qz,qw,tx,ty,tz) (NULL = identity). */
int dsc_problem_upload(dsc_ctx* ctx, const dsc_pair* pair, int n,
                       const float* X1, const float* X2, const float* uv1, const float* uv2,
                       const double* depth1, const double* depth2,
                       const float* inv_sigma2_1, const float* inv_sigma2_2,
                       double scale1, double scale2, const double* Tg7);
/* symmetric CSR neighbour graph over the n correspondences: rowptr[n+1], col[E], w[E]
 * (w[i->j] == w[j->i]), mesh area, triangle count (ARAP information = arap * n_triangles^2).
 * reorder != 0 lets the library renumber the correspondences along a space-filling curve
 * internally; all results come back in the caller's numbering. */
int dsc_set_graph(dsc_ctx* ctx, int n, const int32_t* rowptr, const int32_t* col, const double* w,
                  double area, long long n_triangles, int reorder);
int dsc_compute_rotations(dsc_ctx* ctx);
/* rotations as unit quaternions (x,y,z,w) [n][4] */
int dsc_get_rotations(dsc_ctx* ctx, double* quat);
int dsc_set_rotations(dsc_ctx* ctx, const double* quat);
/* back to the uploaded points / scales / T_global (replaces Map::clone() for the weight search) */
int dsc_reset_state(dsc_ctx* ctx);
int dsc_set_pcg(dsc_ctx* ctx, const dsc_pcg_params* prm);
/* Linear solver of the LM step (the reference: g2o LinearSolverEigen, a sparse Cholesky, g2oBundleAdjustment.cc:619-628).
 * DSC_SOLVER_AUTO (default): dense Cholesky on the device up to DSC_DENSE_AUTO_MAX correspondences (the reference's own
 * problem sizes, where the PCG is bound by launch latency), PCG above; DSC_SOLVER_PCG / DSC_SOLVER_DENSE force one
 * (dense: n <= DSC_DENSE_MAX, else DSC_ERR_INVALID_ARG from dsc_optimize). */
#define DSC_SOLVER_AUTO 0
#define DSC_SOLVER_PCG 1
#define DSC_SOLVER_DENSE 2
#define DSC_DENSE_AUTO_MAX 600
#define DSC_DENSE_MAX 1000
int dsc_set_solver(dsc_ctx* ctx, int solver);
/* Storage precision of the linear solver (the north_star's second tolerance: "1e-3 px reprojection RMSE in fp32 mode").
 * DSC_PRECISION_F64 (default): everything double, the 1e-5 parity bar against the direct-solve oracle.
 * DSC_PRECISION_F32: the data the PCG streams -- the per-edge ARAP Jacobian records, the unary Hessian records, the
 * block-Jacobi preconditioner and the six PCG vectors -- are STORED as float (half the bytes per PCG iteration); every
 * sum, dot product, the 8x8 global block, the state (points, scales, T_global), the cost and the LM decisions stay
 * double, and the reprojection is float32 as in the reference (g2oTypes.h:284-289) in both modes.  An LM step is then an
 * inexact Newton step (relative error ~1e-6), which moves the per-iteration costs by ~1e-6 relative and the final
 * reprojection RMSE by far less than 1e-3 px (tests/test_gpu_parity.py).  A mode of the PCG path: the dense and the
 * one-launch cluster solvers of small problems are not used while it is set.  May be switched between dsc_optimize calls. */
#define DSC_PRECISION_F64 0
#define DSC_PRECISION_F32 1
int dsc_set_precision(dsc_ctx* ctx, int precision);
/* Optional (off by default): pause every linear solve at up to 4 loose tolerances rtol_loose[0] > rtol_loose[1] > ...,
 * evaluate the trial step there, and reject it at once when rho < -rho_margin[level]; otherwise resume the same CG.
 * The last pass always runs to the tight tolerance of dsc_set_pcg, so accepted steps are unchanged.  A rejected step
 * only uses the sign of rho (lambda *= ni either way), so the LM trace is unchanged unless rho changes sign between
 * a loose and the tight tolerance, which the margins guard against.  n_levels = 0 switches it off. */
int dsc_set_early_reject(dsc_ctx* ctx, int n_levels, const double* rtol_loose, const double* rho_margin);
/* total robust chi2 of the current state; parts[3] = reprojection, depth, ARAP (may be NULL) */
int dsc_cost(dsc_ctx* ctx, const dsc_weights* w, double* chi2, double* parts);
/* n_iters LM iterations; records[n_iters] and stats may be NULL */
int dsc_optimize(dsc_ctx* ctx, const dsc_weights* w, int n_iters, dsc_iter_record* records, dsc_opt_stats* stats);
/* X1/X2 [n][3] float (cast like the reference), fp64 copies X1d/X2d [n][3] (may be NULL),
 * scales[2], Tg7, update = sum |p_uploaded - p_now| over all 2n points (float norm, double sum) */
int dsc_download(dsc_ctx* ctx, float* X1, float* X2, double* X1d, double* X2d,
                 double* scales, double* Tg7, double* update);
/* calculatePixelsStandDev on the current state: sigma[2] = C1, C2 */
int dsc_pixel_sigma(dsc_ctx* ctx, double* sigma);

/* gradient b (8 + 6n, oracle layout [T(6) s1 s2 | X1_0 X2_0 ...]) and diagonal of H for the current
 * state -- test hook for the assembly kernels */
int dsc_debug_linearize(dsc_ctx* ctx, const dsc_weights* w, double* b, double* hdiag, double* chi2);
/* y = (H + lambda I) x with the matrix-free operator the PCG uses -- test hook */
int dsc_debug_matvec(dsc_ctx* ctx, const dsc_weights* w, double lambda, const double* x, double* y);

/* Per-kernel device times for the roofline: each hot kernel is launched `reps` times back to back on
 * the current state (after `warm` untimed launches) between two CUDA events on the context's stream.
 * ms[] receives the average launch duration in milliseconds, indexed by DSC_K_*; bytes[] the algorithmic
 * bytes one launch moves.  The solver state (points, vectors) is left unchanged
 * except for the CG scratch vectors. */
enum { DSC_K_SPMV = 0, DSC_K_UPDATE = 1, DSC_K_LINEARIZE = 2, DSC_K_COST = 3, DSC_K_PRECOND = 4,
       DSC_K_APPLY = 5, DSC_K_ROTATIONS = 6, DSC_K_COUNT = 7 };
int dsc_profile_kernels(dsc_ctx* ctx, const dsc_weights* w, int warm, int reps, double* ms, double* bytes);
/* same for the triangulation kernel (needs dsc_tri_upload first) */
int dsc_profile_triangulate(dsc_ctx* ctx, const dsc_tri_params* prm, int warm, int reps, double* ms, double* bytes);
/* ---- graph set-up on the GPU ----------------------------------------------------------------
 * Symmetrised k-nearest-neighbour graph in the plane (x, y) of n points X[n][3] (host, float), k <= 32: uniform grid +
 * ring search, rows ascending.  This is the adjacency the synthetic 100k / 1M configurations use in place of the
 * reference's Delaunay mesh (Modules/Utils/Geometry.cc:317-368).  dsc_knn_build leaves the CSR on the device and
 * reports the number of directed edges; dsc_knn_download copies rowptr[n+1] / col[E] out (feed them to dsc_set_graph). */
int dsc_knn_build(dsc_ctx* ctx, int n, const float* X, int k, long long* n_edges);
int dsc_knn_download(dsc_ctx* ctx, int32_t* rowptr, int32_t* col);
/* The reference's own neighbour graph on the GPU (Modules/Utils/Geometry.cc:272-368: ComputeDelaunayTriangulation3D -- the 2-D
 * Delaunay triangulation of the points' world (x, y), Qhull "d Qbb Qt" --, TriangleMesh adjacency, ComputeEdgeWeightsCot, mesh
 * area and triangle count, g2oBundleAdjustment.cc:657-662,942-946).  Every point clips its own Voronoi cell from the points
 * around it (uniform grid, ring by ring, certified by the security radius); the ~sqrt(n) cells that reach outside the point
 * cloud are finished in a second pass (n_second_pass reports how many); the thin triangles along the convex hull are kept, as
 * Qhull keeps them (the host triangulator host/Mesh.h loses the thinnest to its finite super triangle).  Rows ascending; w = mean over the adjacent
 * triangles of the cotangent of the opposite angle, clamped to >= min_weight; area in 3-D.  DSC_ERR_GRAPH on inputs this
 * construction does not handle (a cell with more than 32 vertices: e.g. many co-circular points).
 *   dsc_delaunay_build / dsc_delaunay_download   stand-alone, X[n][3] host floats -> CSR (feed it to dsc_set_graph)
 *   dsc_set_graph_delaunay                       the graph of the UPLOADED problem (its KF1 points), built and installed on
 *                                                the device: replaces host mesh + dsc_set_graph; reorder as for dsc_set_graph */
int dsc_delaunay_build(dsc_ctx* ctx, int n, const float* X, double min_weight, long long* n_edges, long long* n_triangles, double* area,
                       long long* n_second_pass);
int dsc_delaunay_download(dsc_ctx* ctx, int32_t* rowptr, int32_t* col, double* w);
int dsc_set_graph_delaunay(dsc_ctx* ctx, double min_weight, int reorder, double* area, long long* n_triangles, long long* n_edges);
/* problem size as seen by the library: n correspondences, E directed edges */
int dsc_problem_size(const dsc_ctx* ctx, long long* n, long long* n_edges);

/* ---- batches of independent frame pairs (BASELINE.json configs[4], and the <= 31 independent
 * refinements per outer iteration of the weight search, Modules/Optimization/nloptOptimization.cc:5-37) ---------------
 * Arrays of pair descriptors with prefix offsets: pair p owns correspondences point_offset[p] .. point_offset[p+1] of
 * the concatenated point arrays (same element layout as dsc_problem_upload) and the directed edges edge_offset[p] ..
 * edge_offset[p+1] of col / w; rowptr is the concatenation of the pairs' own CSR row pointers (n_p + 1 entries each,
 * every one starting at 0: pair p's begin at rowptr[point_offset[p] + p]).  dsc_batch_optimize refines ALL pairs in one
 * kernel launch: thread-block clusters take pairs from a device-side queue and run the whole Levenberg-Marquardt loop
 * of a pair (the same algorithm as dsc_optimize, PCG linear solves) without returning to the host.  n_weights = 1: one
 * dsc_weights for every pair; n_weights = n_problems: one each (the weight search).  records[n_problems][n_iters] and
 * stats[n_problems] may be NULL; device_ms = CUDA-event time of the launch. */
typedef struct dsc_batch dsc_batch;
typedef struct {
    dsc_pair  pair;
    double    scale1, scale2;      /* initial depth scales                                             */
    double    Tg7[7];              /* initial T_global (qx qy qz qw tx ty tz); all-zero quaternion = identity */
    double    area;                /* mesh area                                                        */
    long long n_triangles;         /* ARAP information = arap * n_triangles^2                          */
} dsc_batch_pair;
int  dsc_batch_create(int device, dsc_batch** out);
void dsc_batch_destroy(dsc_batch* batch);
const char* dsc_batch_last_error(const dsc_batch* batch);
int  dsc_batch_upload(dsc_batch* batch, int n_problems, const dsc_batch_pair* pairs, const long long* point_offset,
                      const float* X1, const float* X2, const float* uv1, const float* uv2,
                      const double* depth1, const double* depth2, const float* inv_sigma2_1, const float* inv_sigma2_2,
                      const long long* edge_offset, const int32_t* rowptr, const int32_t* col, const double* w, int reorder);
int  dsc_batch_set_pcg(dsc_batch* batch, const dsc_pcg_params* prm);
int  dsc_batch_set_early_reject(dsc_batch* batch, int n_levels, const double* rtol_loose, const double* rho_margin);
int  dsc_batch_reset_state(dsc_batch* batch);
/* the next launches (optimize / reset_state / pixel_sigma) work on pairs [0, n_active) only; -1 = all.  The weight search
 * keeps K replicas of one pair resident and refines as many as a Nelder-Mead step has candidate weights. */
int  dsc_batch_set_active(dsc_batch* batch, int n_active);
/* calculatePixelsStandDev (Modules/Utils/Geometry.cc:370-498) of every active pair's current state: sigma[n_active][2] */
int  dsc_batch_pixel_sigma(dsc_batch* batch, double* sigma);
int  dsc_batch_optimize(dsc_batch* batch, const dsc_weights* weights, int n_weights, int n_iters, dsc_iter_record* records,
                        dsc_opt_stats* stats, double* device_ms);
/* results in the callers' numbering, concatenated like the inputs: X1/X2 [sum n][3] float, scales [n_problems][2],
 * Tg7 [n_problems][7], update [n_problems] (any may be NULL) */
int  dsc_batch_download(dsc_batch* batch, float* X1, float* X2, double* scales, double* Tg7, double* update);
/* pairs uploaded, their correspondences, CTAs per cluster and clusters of the last launch */
int  dsc_batch_size(const dsc_batch* batch, int* n_problems, long long* n_points, int* cluster_ctas, int* clusters);

/* ---- classic bundle adjustment: bundleAdjustment / poseOnlyOptimization / localBundleAdjustment of
 * Modules/Optimization/g2oBundleAdjustment.cc:38-444 -- key-frame poses (g2o::VertexSE3Expmap, Tcw), map points
 * (VertexSBAPointXYZ, marginalised: BlockSolver_6_3) and reprojection edges (EdgeSE3ProjectXYZ / ...OnlyPose, g2oTypes.h:150-228)
 * with information invSigma2 * I and an optional Huber kernel.  Levenberg-Marquardt as g2o runs it; the points are
 * eliminated by a Schur complement summed on the device, the reduced camera system (6 x free poses) is factorised on the
 * host.  poses7: [n_poses][7] = unit quaternion (x y z w) + translation of Tcw; pose_fixed: setFixed(true) (key frame id 0,
 * the fixed key frames of the local map); points_fixed != 0: the points are constants of the edges (pose-only).
 * At most 512 free poses (the reduced system is dense on the host); a point may be observed once per pose.
 *   dsc_ba_set_levels   active[n_obs] != 0: edge at level 0 (optimised), 0: level 1 (left out); NULL = all at level 0
 *   dsc_ba_optimize     optimizer.optimize(n_iters); huber_delta <= 0: no robust kernel (setRobustKernel(0))
 *   dsc_ba_edge_chi2    e->chi2() and e->isDepthPositive() of every edge at the current estimate (any level)
 *   dsc_ba_set_poses    setEstimate of every pose (poseOnlyOptimization restarts each round from the frame's pose) */
typedef struct dsc_ba dsc_ba;
int  dsc_ba_create(int device, dsc_ba** out);
void dsc_ba_destroy(dsc_ba* ba);
const char* dsc_ba_last_error(const dsc_ba* ba);
int  dsc_ba_upload(dsc_ba* ba, int n_poses, const double* poses7, const uint8_t* pose_fixed, const dsc_camera* cams, int n_points,
                   const double* X, int points_fixed, long long n_obs, const int32_t* obs_pose, const int32_t* obs_point,
                   const float* obs_uv, const float* obs_inv_sigma2);
int  dsc_ba_set_poses(dsc_ba* ba, const double* poses7);
int  dsc_ba_set_levels(dsc_ba* ba, const uint8_t* active);
int  dsc_ba_optimize(dsc_ba* ba, int n_iters, double huber_delta, dsc_iter_record* records, dsc_opt_stats* stats);
int  dsc_ba_edge_chi2(dsc_ba* ba, double* chi2, uint8_t* depth_positive);
int  dsc_ba_download(dsc_ba* ba, double* poses7, double* X);
int  dsc_ba_launch_count(const dsc_ba* ba, long long* count);

/* ---- ONE frame pair over several GPUs (second row: "point-sharded edge evaluation with an allreduce of
 * the small reduced system over NVLink"; the graph it partitions is the one of g2oBundleAdjustment.cc:883-953) ------------
 * One process per GPU, one context per process.  Every rank uploads the SAME pair and graph and calls the SAME sequence
 * of entry points (dsc_problem_upload, dsc_set_graph, dsc_compute_rotations, dsc_optimize, dsc_download, ...); the
 * library splits the correspondences into contiguous ranges of tiles of its internal (space-filling-curve) order, each
 * rank linearises / solves / evaluates its own range, and everything that crosses NVLink is written by the kernels
 * themselves into peer-mapped memory: halo rows of the PCG vector z and of the trial state are pushed by their owners,
 * the PCG scalars and the 8 global rows are exchanged by the last block of the producing kernel (no collective library
 * call on the solve path; sums over the ranks are taken in rank order, so every rank and every run sees the same bits).
 *   dsc_shard_init     before dsc_problem_upload: allocates the rank's exchange arena for up to max_points
 *                      correspondences and returns its 64-byte inter-process handle (cudaIpcMemHandle_t);
 *   dsc_shard_attach   handles[world][64] of all ranks, in rank order (exchange them with any host-side all-gather);
 *   dsc_shard_partition  the row partition the library uses (pure host function): sliceptr[nslices + 1] of the sliced ELL
 *                      -> row_begin[world + 1], multiples of 512 rows, equal work per rank;
 *   dsc_shard_info     rank, world, own row range (internal numbering), own rows that are halo rows of a peer.
 * A rank that waits for a peer longer than a few seconds gives up: the entry point returns DSC_ERR_SHARD.
 * fp64 PCG path only (DSC_PRECISION_F32 and the dense / one-launch solvers are refused on a sharded context). */
#define DSC_SHARD_HANDLE_BYTES 64
#define DSC_SHARD_MAX_RANKS 8
int dsc_shard_init(dsc_ctx* ctx, int rank, int world, int max_points, void* handle_out);
int dsc_shard_attach(dsc_ctx* ctx, const void* handles);
int dsc_shard_partition(const int32_t* sliceptr, int nslices, int world, int32_t* row_begin);
int dsc_shard_info(const dsc_ctx* ctx, int* rank, int* world, int* row_begin, int* row_end, long long* halo_rows);

#ifdef __cplusplus
}
#endif
#endif /* DSC_H_ */
