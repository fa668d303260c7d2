// API-CONFORMANCE STUB (test infrastructure, tests/test_host_shim.py::test_shim_compiles_against_the_reference_api).
// Declares ONLY members that the reference declares in Modules/Map/Map.h:38-224, with the reference's own signatures
// (every declaration below is checked, line for line, against that header when /root/reference is present); no bodies.
// host/Optimization.cc is compiled against this tree with -DDSC_IN_REFERENCE_TREE: anything it calls that the reference
// does not declare fails that build.
#pragma once
#include <memory>
#include <set>
#include <unordered_map>
#include "Map/KeyFrame.h"
#include "Map/MapPoint.h"

typedef long unsigned int ID;

class Map {
public:
    void insertMapPoint(std::shared_ptr<MapPoint> pMP);
    std::shared_ptr<KeyFrame> getKeyFrame(ID id);
    void addObservation(ID kfId, ID mpId, size_t idx);
    void removeObservation(ID kfId, ID mpId);
    void getLocalMapOfKeyFrame(ID kfId, std::set<ID>& sLocalMapPointsIds, std::set<ID>& sLocalKeyFramesIds, std::set<ID>& sLocalFixedKeyFramesIds);
    void checkKeyFrame(ID kfId);
    std::unordered_map<ID,std::shared_ptr<MapPoint>>& getMapPoints();
    std::unordered_map<ID,std::shared_ptr<KeyFrame>>& getKeyFrames();
    int isMapPointInKeyFrame(ID mp, ID kf);
    void insertGlobalKeyFramesTransformation(ID kf1, ID kf2, const Sophus::SE3f& transformation);
    Sophus::SE3f getGlobalKeyFramesTransformation(ID kf1, ID kf2);
};
