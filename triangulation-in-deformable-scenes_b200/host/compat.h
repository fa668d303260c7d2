// compat.h -- the few Eigen / Sophus / OpenCV value types that appear in the reference's hot-path API.
//
// Inside the reference tree (Eigen, Sophus and OpenCV present) the real headers are used and this file
// only adds nothing.  In this image none of them exists, so minimal stand-ins with the SAME names and the
// members the hot-path API touches are provided; the call signatures of host/Optimization.h are then
// literally those of Modules/Optimization/g2oBundleAdjustment.h:52-59 and Modules/Utils/Geometry.h:66-69.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

#if __has_include(<Eigen/Core>) && __has_include(<sophus/se3.hpp>) && __has_include(<opencv2/core.hpp>)
#include <Eigen/Core>
#include <opencv2/core.hpp>
#include <sophus/se3.hpp>
#define DSC_HOST_REAL_MATH 1
#else
#define DSC_HOST_REAL_MATH 0

#include "compat_types.h"
#endif

namespace dsc_host {
// row-major 3x4 [R|t] of a pose, the layout the C ABI takes
inline void pose34(const Sophus::SE3f& T, float* out) {
    auto R = T.rotationMatrix();
    auto t = T.translation();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out[r * 4 + c] = R(r, c); out[r * 4 + 3] = t[r]; }
}
}  // namespace dsc_host
