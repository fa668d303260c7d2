// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/System/Settings.h, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
// host/Optimization.cc is compiled against this tree with -DDSC_IN_REFERENCE_TREE: anything it calls that the reference
// does not declare fails that build.
#pragma once
#include <string>
#include "Calibration/CameraModel.h"

class Settings {
public:
    bool getCheckingSelection();
    float getDepthLimit();
    bool getDrawRaysSelection();
    float getMinCos();
    double getNloptAbsTolerance();
    double getNloptArapLowerBound();
    double getNloptArapUpperBound();
    double getNloptGlobalLowerBound();
    double getNloptGlobalUpperBound();
    double getNloptRelTolerance();
    double getNloptRepLowerBound();
    double getNloptRepUpperBound();
    int getNloptnOptimizations();
    double getOptAlphaWeight();
    double getOptArapWeight();
    double getOptBetaWeight();
    double getOptGlobalWeight();
    double getOptRepWeight();
    std::string getOptSelection();
    std::string getOptWeightsSelection();
    float getSimulatedDepthWeight();
    std::string getTrianLocation();
    std::string getTrianMethod();
    int getnOptIterations();
    int getnOptimizations();
};
