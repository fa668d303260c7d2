// Optimization.h -- the reference's optimisation-call API, served by the CUDA library behind include/dsc.h.
//
// Same names, argument meaning and order as
//   Modules/Optimization/g2oBundleAdjustment.h:52-59   deformationOptimization, arapOptimization
//   Modules/Utils/Geometry.h:66-69                      useTriangulationMethod
// so Modules/System/SLAM.cc:127,145, Modules/Optimization/nloptOptimization.cc:20, Modules/Mapping/Mapping.cc:311
// and Modules/Mapping/MonocularMapInitializer.cc:313 compile against this header unchanged.
//
// Two builds of the same sources:
//   * stand-alone (this image: no Eigen / Sophus / OpenCV): Map.h / Settings.h / compat.h of this directory provide the
//     data model with the reference's names;
//   * -DDSC_IN_REFERENCE_TREE: the reference's own headers (Map/Map.h, Map/KeyFrame.h, Calibration/*.h, System/Settings.h,
//     Utils/CommonTypes.h) are used instead and nothing of that data model is redefined here.  tests/refapi/ holds a
//     header tree with exactly the reference's declarations of the members this shim calls; the conformance test
//     compiles Optimization.cc against it.
#pragma once
#include <memory>
#include <string>
#include <vector>

#ifdef DSC_IN_REFERENCE_TREE
#include "Calibration/CameraModel.h"
#include "Calibration/KannalaBrandt8.h"
#include "Calibration/PinHole.h"
#include "Map/KeyFrame.h"
#include "Map/Map.h"
#include "Map/MapPoint.h"
#include "Mapping/Frame.h"
#include "System/Settings.h"
#include "Utils/CommonTypes.h"
#include "Visualization/MapVisualizer.h"
#else
#include "Map.h"
#include "Settings.h"
#endif

/* The classic bundle-adjustment paths (g2oBundleAdjustment.h:36-47, bodies g2oBundleAdjustment.cc:38-444), on dsc_ba_*:
 * full BA of every key frame and map point (key frame 0 fixed, 20 iterations, Huber sqrt(5.99));  pose-only optimisation of a
 * frame (four rounds of ten iterations, inliers re-classified at chi2 5.991; returns the number of inliers, outliers lose their
 * map point);  BA of the local map of a key frame (5 robust + 10 plain iterations, outlier observations removed from the map). */
void bundleAdjustment(Map* pMap);
int poseOnlyOptimization(Frame& currFrame);
void localBundleAdjustment(Map* pMap, ID currKeyFrameId);

/* Performs an As-Rigid-As-Possible optimization using arapOptimization inside an external loop that optimizes the
 * balance weights (g2oBundleAdjustment.cc:446-606). */
void deformationOptimization(std::shared_ptr<Map> pMap, Settings& settings, std::shared_ptr<MapVisualizer>& mapVisualizer,
                             const std::vector<Eigen::Vector3f> originalPoints = {}, const std::vector<Eigen::Vector3f> movedPoints = {});

/* ARAP + reprojection + depth refinement of the map points of every key-frame pair (g2oBundleAdjustment.cc:608-1008). */
void arapOptimization(Map* pMap, double repBalanceWeight, double globalBalanceWeight, double arapBalanceWeight, double alphaWeight,
                      double betaWeight, float DepthError, int nOptIterations, double* optimizationUpdate = nullptr);

/* One correspondence (Geometry.cc:216-230), rays in, always true -- as in the reference.  Kept for API parity: it
 * launches a 1-element batch (dsc_triangulate_rays); callers with many matches should use triangulateMatches. */
bool useTriangulationMethod(const Eigen::Vector3f& xn1, const Eigen::Vector3f& xn2, const Sophus::SE3f& T1w, const Sophus::SE3f& T2w,
                            Eigen::Vector3f& x3D_1, Eigen::Vector3f& x3D_2, std::string TrianMethod, std::string TrianLocation);

/* Pixel-error standard deviation of the two cameras (Utils/Geometry.cc:370-498).  PixelsError is the reference's
 * struct (Utils/CommonTypes.h:23-30); it is only defined here when that header is not in the build. */
#ifndef DSC_IN_REFERENCE_TREE
struct PixelsError { double avgc1 = 0, avgc2 = 0, avg = 0, desvc1 = 0, desvc2 = 0, desv = 0; };
#endif
void calculatePixelsStandDev(std::shared_ptr<Map> Map, PixelsError& pixelsErrors);

namespace dsc_host {

struct TriangulationResult {
    std::vector<Eigen::Vector3f> x3D_1, x3D_2;
    std::vector<unsigned char> valid;
    std::vector<float> cosParallax;
    int nValid = 0;
};

/* Batched replacement of the per-match loops of Mapping::triangulateSimulatedMapPoints (Mapping.cc:294-343) and
 * MonocularMapInitializer::reconstructPoints (:303-368): key points in pixels, gates as DSC_GATE_*. */
TriangulationResult triangulateMatches(KeyFrame& refKF, KeyFrame& currKF, const std::vector<int>& matches /* curr index per ref key point, -1 = none */,
                                       const std::string& method, const std::string& location, int gate, float minCos,
                                       float depthLimit = 3.0e38f, bool checkReprojection = false);

/* Mapping::triangulateSimulatedMapPoints end to end: triangulate, create MapPoints/observations for the valid
 * matches, initial depth scales (KeyFrame.cc:131-153).  Returns the number of MapPoints created. */
int triangulateSimulatedMapPoints(Map& map, KeyFrame_ refKF, KeyFrame_ currKF, const std::string& method, const std::string& location, float minCos);

/* Real-image map initialisation, Mapping::monocularMapInitialization after the matching step (Mapping.cc:159-257):
 * MonocularMapInitializer::reconstructPoints (:303-368: triangulate every match with the given poses, gates
 * finite/non-zero, 0 <= z <= depthLimit in both cameras, optional reprojection^2 <= 5.991), then MapPoints and
 * observations for the matches whose depth measurements are positive and whose key points lie in (0.1, 1500)
 * (:183-209; the curr MapPoint is stored at the REF slot, its observation at the matched index), and the initial
 * depth scales = mean d_measured / z over the points whose parallax in degrees exceeds Triangulation.minCos
 * (:211-254).  Needs depth images on both key frames.  Returns the number of MapPoints created. */
int initializeMapFromMatches(Map& map, KeyFrame_ refKF, KeyFrame_ currKF, const std::vector<int>& matches, Settings& settings,
                             float* parallaxDegrees = nullptr);

/* Per-iteration record of the last arapOptimization call (the reference runs g2o with setVerbose(false); the
 * parity tests need the trace). */
struct LmRecord { double chi2_before, chi2_after, lambda; int trials, accepted, pcg_iters; };
const std::vector<LmRecord>& lastTrace();

/* Wall-clock phases of the last arapOptimization call (last key-frame pair): Map -> arrays, upload + mesh + graph set-up +
 * rotations (device), the LM iterations, download + Map write-back. */
struct PhaseTimes { double gather_ms, setup_ms, lm_ms, writeback_ms; long long correspondences; };
const PhaseTimes& lastPhaseTimes();

/* PCG controls of the linear solve (defaults rtol 1e-10, 6000 iterations). */
void setSolver(double rtol, int maxIters);

}  // namespace dsc_host
