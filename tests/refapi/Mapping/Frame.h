// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/Mapping/Frame.h:38-233, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
#pragma once
#include <memory>
#include <vector>
#include <opencv2/opencv.hpp>
#include <sophus/se3.hpp>
#include "Calibration/CameraModel.h"
#include "Map/MapPoint.h"

class Frame {
public:
    void setPose(Sophus::SE3f& Tcw);
    cv::KeyPoint getKeyPoint(const size_t idx);
    std::vector<std::shared_ptr<MapPoint>>& getMapPoints();
    void setMapPoint(size_t idx, std::shared_ptr<MapPoint> pMP);
    const Sophus::SE3f getPose() const;
    std::shared_ptr<CameraModel> getCalibration();
    float getInvSigma2(int octave);
};
