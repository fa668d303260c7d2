// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/Calibration/CameraModel.h:37-147, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
// host/Optimization.cc is compiled against this tree with -DDSC_IN_REFERENCE_TREE: anything it calls that the reference
// does not declare fails that build.
#pragma once
#include <vector>
#include <Eigen/Core>

class CameraModel {
public:
    virtual void project(const Eigen::Vector3f& p3D, Eigen::Vector2f& p2D) = 0;
    float getParameter(const int i);
    int getNumberOfParameters();
};
