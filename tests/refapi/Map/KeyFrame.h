// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/Map/KeyFrame.h:29-233, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
// host/Optimization.cc is compiled against this tree with -DDSC_IN_REFERENCE_TREE: anything it calls that the reference
// does not declare fails that build.
#pragma once
#include <memory>
#include <vector>
#include <Eigen/Core>
#include <opencv2/opencv.hpp>
#include <sophus/se3.hpp>
#include "Calibration/CameraModel.h"
#include "Map/MapPoint.h"

class KeyFrame {
public:
    Sophus::SE3f getPose();
    void setPose(Sophus::SE3f& Tcw);
    cv::KeyPoint getKeyPoint(size_t idx);
    std::vector<cv::KeyPoint>& getKeyPoints();
    std::vector<float>& getDepthMeasurements();
    double getEstimatedDepthScale();
    void setEstimatedDepthScale(double scale);
    std::vector<std::shared_ptr<MapPoint>>& getMapPoints();
    void setMapPoint(size_t idx, std::shared_ptr<MapPoint> pMP);
    std::shared_ptr<CameraModel> getCalibration();
    long unsigned int getId();
    float getInvSigma2(int octave);
    float getDepthMeasure(size_t idx);
    double getDepthMeasure(float x, float y, bool scaled = true);
};
