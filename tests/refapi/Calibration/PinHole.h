// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/Calibration/PinHole.h:34, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
// host/Optimization.cc is compiled against this tree with -DDSC_IN_REFERENCE_TREE: anything it calls that the reference
// does not declare fails that build.
#pragma once
#include "Calibration/CameraModel.h"

class PinHole : public CameraModel{
};
