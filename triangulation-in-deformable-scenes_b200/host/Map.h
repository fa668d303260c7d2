// Map.h -- host data model kept API-compatible with the reference for the members the hot path touches:
//   Modules/Calibration/CameraModel.h:37-147, KannalaBrandt8.h, PinHole.h
//   Modules/Map/MapPoint.h, Modules/Map/KeyFrame.h:29-233, Modules/Map/Map.h:38-224 (+ Map.cc:30-58,257-264,323-343)
// Image-side members (descriptors, grids, pyramids' images, covisibility) are out of scope.
#pragma once
#include <set>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "compat.h"

typedef long unsigned int ID;

// ------------------------------------------------------------------ calibration
class CameraModel {
public:
    CameraModel() {}
    explicit CameraModel(const std::vector<float>& p) : vParameters_(p) {}
    virtual ~CameraModel() {}
    virtual int modelId() const = 0;                       // DSC_CAM_KB8 / DSC_CAM_PINHOLE
    virtual void project(const Eigen::Vector3f& p3D, Eigen::Vector2f& p2D) = 0;
    virtual void unproject(const Eigen::Vector2f& p2D, Eigen::Vector3f& p3D) = 0;
    cv::Point2f project(Eigen::Vector3f& X) { Eigen::Vector2f uv; project(X, uv); return cv::Point2f(uv[0], uv[1]); }
    Eigen::Vector3f unproject(cv::Point2f puv) { Eigen::Vector3f r; unproject(Eigen::Vector2f(puv.x, puv.y), r); return r; }
    float getParameter(int i) const { return vParameters_[i]; }
    const std::vector<float>& getParameters() const { return vParameters_; }
    int getNumberOfParameters() const { return (int)vParameters_.size(); }
protected:
    std::vector<float> vParameters_;
};

// Modules/Calibration/KannalaBrandt8.cc:32-83.  Host copies exist for the callers' front end (key-point
// synthesis in SLAM::createKeyPoints); the hot path projects on the GPU.
class KannalaBrandt8 : public CameraModel {
public:
    explicit KannalaBrandt8(const std::vector<float>& p) : CameraModel(p) { vParameters_.resize(8, 0.f); }
    int modelId() const override { return 0; }
    using CameraModel::project;
    using CameraModel::unproject;
    void project(const Eigen::Vector3f& X, Eigen::Vector2f& uv) override {
        const std::vector<float>& P = vParameters_;
        const float x2y2 = X[0] * X[0] + X[1] * X[1];
        const float theta = atan2f(sqrtf(x2y2), X[2]);
        const float psi = atan2f(X[1], X[0]);
        const float t2 = theta * theta, t3 = theta * t2, t5 = t3 * t2, t7 = t5 * t2, t9 = t7 * t2;
        const float r = theta + P[4] * t3 + P[5] * t5 + P[6] * t7 + P[7] * t9;
        uv[0] = P[0] * r * cosf(psi) + P[2];
        uv[1] = P[1] * r * sinf(psi) + P[3];
    }
    void unproject(const Eigen::Vector2f& uv, Eigen::Vector3f& ray) override {
        const std::vector<float>& P = vParameters_;
        const float pwx = (uv[0] - P[2]) / P[0], pwy = (uv[1] - P[3]) / P[1];
        const float theta_d = sqrtf(pwx * pwx + pwy * pwy);
        if (!(theta_d > 1e-8)) { ray = Eigen::Vector3f(0.f, 0.f, 1.f); return; }
        float theta = theta_d;
        for (int j = 0; j < 10; ++j) {
            float t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
            float k0 = P[4] * t2, k1 = P[5] * t4, k2 = P[6] * t6, k3 = P[7] * t8;
            float fix = (theta * (1 + k0 + k1 + k2 + k3) - theta_d) / (1 + 3 * k0 + 5 * k1 + 7 * k2 + 9 * k3);
            theta -= fix;
            if (fabsf(fix) < 1e-6f) break;
        }
        ray = Eigen::Vector3f(sinf(theta) * pwx / theta_d, sinf(theta) * pwy / theta_d, cosf(theta));
    }
};

class PinHole : public CameraModel {
public:
    explicit PinHole(const std::vector<float>& p) : CameraModel(p) { vParameters_.resize(8, 0.f); }
    int modelId() const override { return 1; }
    using CameraModel::project;
    using CameraModel::unproject;
    void project(const Eigen::Vector3f& X, Eigen::Vector2f& uv) override {
        uv[0] = vParameters_[0] * X[0] / X[2] + vParameters_[2];
        uv[1] = vParameters_[1] * X[1] / X[2] + vParameters_[3];
    }
    void unproject(const Eigen::Vector2f& uv, Eigen::Vector3f& ray) override {
        ray = Eigen::Vector3f((uv[0] - vParameters_[2]) / vParameters_[0], (uv[1] - vParameters_[3]) / vParameters_[1], 1.f);
    }
};

// ------------------------------------------------------------------ MapPoint (MapPoint.h, MapPoint.cc:22-28)
class MapPoint {
public:
    explicit MapPoint(Eigen::Vector3f& p3d) : position3D_(p3d), id_(nNextId_++) {}
    MapPoint(const MapPoint& o) : position3D_(o.position3D_), id_(o.id_) {}
    MapPoint* clone() const { return new MapPoint(*this); }
    Eigen::Vector3f getWorldPosition() { return position3D_; }
    void setWorldPosition(Eigen::Vector3f& p3d) { position3D_ = p3d; }
    long unsigned int getId() { return id_; }
    static void restartIds() { nNextId_ = 0; }          // (tests: the ids of a fresh process)
private:
    Eigen::Vector3f position3D_;
    long unsigned int id_;
    inline static long unsigned int nNextId_ = 0;
};
typedef std::shared_ptr<MapPoint> MapPoint_;

// ------------------------------------------------------------------ KeyFrame (KeyFrame.h:29-233)
class KeyFrame {
public:
    // promoted from a frame: key points, pose, calibration, pyramid scale factor (Frame.cc:57-75)
    KeyFrame(const std::vector<cv::KeyPoint>& keys, const Sophus::SE3f& Tcw, std::shared_ptr<CameraModel> calib,
             int nScales = 8, float scaleFactor = 1.2f)
        : vKeys_(keys), Tcw_(Tcw), calibration_(calib), id_(nNextId_++) {
        vMapPoints_.assign(keys.size(), nullptr);
        vDepthMeasurements_.assign(keys.size(), 0.f);
        vInvSigma2_.resize(nScales);
        float s = 1.0f;
        for (int i = 0; i < nScales; ++i) { vInvSigma2_[i] = 1.0f / (s * s); s *= scaleFactor; }
    }
    KeyFrame* clone() const { return new KeyFrame(*this); }
    Sophus::SE3f getPose() { return Tcw_; }
    void setPose(Sophus::SE3f& Tcw) { Tcw_ = Tcw; }
    cv::KeyPoint getKeyPoint(size_t idx) { return vKeys_[idx]; }
    std::vector<cv::KeyPoint>& getKeyPoints() { return vKeys_; }
    // simulated images (KeyFrame.cc:123-129)
    float getDepthMeasure(size_t idx) { return vDepthMeasurements_[idx]; }
    std::vector<float>& getDepthMeasurements() { return vDepthMeasurements_; }
    void setDepthMeasure(float d, size_t idx) { vDepthMeasurements_[idx] = d; }
    // real images: bilinear sample of the depth image / 100, times imageDepthScale unless `scaled`
    // (KeyFrame.cc:181-202, Utils/Geometry.cc:607-619)
    void setDepthImage(const std::vector<float>& im, int cols, int rows, double imageDepthScale) {
        depthIm_ = im; depthCols_ = cols; depthRows_ = rows; imageDepthScale_ = imageDepthScale;
    }
    bool hasDepthImage() const { return !depthIm_.empty(); }
    double getDepthMeasure(float x, float y, bool scaled = true) {
        if (depthIm_.empty()) throw std::runtime_error("Depth image is not initialized.");
        if (x >= depthCols_ || y >= depthRows_) throw std::out_of_range("Pixel coordinates are out of range.");
        float xi, yi;
        float fx = modff(x, &xi), fy = modff(y, &yi);
        float w00 = (1.f - fx) * (1.f - fy), w01 = (1.f - fx) * fy, w10 = fx * (1.f - fy), w11 = 1.f - w00 - w01 - w10;
        const float* m = depthIm_.data();
        int c = depthCols_, X = (int)xi, Y = (int)yi;
        auto at = [&](int yy, int xx) { size_t k = (size_t)yy * c + xx; return k < depthIm_.size() ? m[k] : 0.f; };
        float g = at(Y, X) * w00 + at(Y, X + 1) * w10 + at(Y + 1, X) * w01 + at(Y + 1, X + 1) * w11;
        double depth = (double)g / 100;
        return scaled ? depth : depth * imageDepthScale_;
    }
    double getEstimatedDepthScale() { return estimatedDepthScale_; }
    void setEstimatedDepthScale(double s) { estimatedDepthScale_ = s; }
    std::vector<MapPoint_>& getMapPoints() { return vMapPoints_; }
    void setMapPoint(size_t idx, MapPoint_ pMP) { vMapPoints_[idx] = pMP; }
    MapPoint_ getMapPoint(size_t idx) { return vMapPoints_[idx]; }
    std::shared_ptr<CameraModel> getCalibration() { return calibration_; }
    long unsigned int getId() { return id_; }
    float getInvSigma2(int octave) { return vInvSigma2_[octave]; }
    int getNumberOfScales() { return (int)vInvSigma2_.size(); }
    void setInitialDepthScaleInSimulationImages();      // KeyFrame.cc:131-153 (runs on the GPU, see Optimization.cc)
    static void restartIds() { nNextId_ = 0; }          // (tests: the ids of a fresh process -- key frame 0 is the fixed one)
private:
    std::vector<cv::KeyPoint> vKeys_;
    std::vector<MapPoint_> vMapPoints_;
    std::vector<float> vDepthMeasurements_;
    std::vector<float> depthIm_;
    int depthCols_ = 0, depthRows_ = 0;
    double imageDepthScale_ = 1.0;
    double estimatedDepthScale_ = 0.0;
    Sophus::SE3f Tcw_;
    std::shared_ptr<CameraModel> calibration_;
    std::vector<float> vInvSigma2_;
    long unsigned int id_;
    inline static long unsigned int nNextId_ = 0;
};
typedef std::shared_ptr<KeyFrame> KeyFrame_;

// ------------------------------------------------------------------ Map (Map.h:38-224)
class Map {
public:
    Map() {}
    void insertMapPoint(MapPoint_ pMP) { mMapPoints_[pMP->getId()] = pMP; }
    void insertKeyFrame(KeyFrame_ pKF) { mKeyFrames_[pKF->getId()] = pKF; }
    KeyFrame_ getKeyFrame(ID id) { auto it = mKeyFrames_.find(id); return it == mKeyFrames_.end() ? nullptr : it->second; }
    MapPoint_ getMapPoint(ID id) { auto it = mMapPoints_.find(id); return it == mMapPoints_.end() ? nullptr : it->second; }
    // Map.cc:100-127: observation tables + covisibility counts
    void addObservation(ID kfId, ID mpId, size_t idx) {
        mKeyFrameObs_[kfId][mpId] = idx; mMapPointObs_[mpId][kfId] = idx;
        for (auto& kv : mMapPointObs_[mpId]) {
            if (kv.first == kfId) continue;
            mCovisibilityGraph_[kfId][kv.first]++; mCovisibilityGraph_[kv.first][kfId]++;
        }
    }
    // Map.cc:134-149
    void removeObservation(ID kfId, ID mpId) {
        mKeyFrameObs_[kfId].erase(mpId); mMapPointObs_[mpId].erase(kfId);
        for (auto& kv : mMapPointObs_[mpId]) { mCovisibilityGraph_[kfId][kv.first]--; mCovisibilityGraph_[kv.first][kfId]--; }
    }
    void checkKeyFrame(ID) {}                      // Map.h:142: debug only, its body is commented out upstream
    void setMinCommonObs(float v) { minCommonObs_ = v; }
    // Map.cc:178-209: the key frame, its covisible key frames (more than minCommonObs shared points), the points they
    // see, and -- fixed -- every other key frame that sees one of those points
    void getLocalMapOfKeyFrame(ID kfId, std::set<ID>& sLocalMapPointsIds, std::set<ID>& sLocalKeyFramesIds, std::set<ID>& sLocalFixedKeyFramesIds) {
        std::set<ID> sAllKFs;
        sLocalKeyFramesIds.insert(kfId);
        for (auto& kv : mKeyFrameObs_[kfId]) sLocalMapPointsIds.insert(kv.first);
        for (auto& kv : mCovisibilityGraph_[kfId])
            if (kv.second > minCommonObs_) {
                sLocalKeyFramesIds.insert(kv.first);
                for (auto& ob : mKeyFrameObs_[kv.first]) sLocalMapPointsIds.insert(ob.first);
            }
        for (ID mp : sLocalMapPointsIds) for (auto& kv : mMapPointObs_[mp]) sAllKFs.insert(kv.first);
        sLocalFixedKeyFramesIds.clear();
        for (ID k : sAllKFs) if (!sLocalKeyFramesIds.count(k)) sLocalFixedKeyFramesIds.insert(k);
    }
    std::unordered_map<ID, MapPoint_>& getMapPoints() { return mMapPoints_; }
    std::unordered_map<ID, KeyFrame_>& getKeyFrames() { return mKeyFrames_; }
    // Map.cc:257-264: index of the key point observing mp in kf, or -1
    int isMapPointInKeyFrame(ID mp, ID kf) {
        auto it = mKeyFrameObs_.find(kf);
        if (it == mKeyFrameObs_.end()) return -1;
        auto jt = it->second.find(mp);
        return jt == it->second.end() ? -1 : (int)jt->second;
    }
    // Map.cc:323-343
    void insertGlobalKeyFramesTransformation(ID kf1, ID kf2, const Sophus::SE3f& T) {
        mGTransformation_[kf1][kf2] = T;
        mGTransformation_[kf2][kf1] = T.inverse();
    }
    Sophus::SE3f getGlobalKeyFramesTransformation(ID kf1, ID kf2) {
        auto it = mGTransformation_.find(kf1);
        if (it != mGTransformation_.end()) { auto jt = it->second.find(kf2); if (jt != it->second.end()) return jt->second; }
        return Sophus::SE3f();          // identity when nothing has been stored yet
    }
    // Map.cc:30-58 deep copy.  The weight search of deformationOptimization does not need it here
    // (dsc_reset_state replaces it); kept for API parity.
    std::shared_ptr<Map> clone() const {
        auto m = std::make_shared<Map>();
        std::unordered_map<MapPoint*, MapPoint_> mp;
        for (auto& kv : mMapPoints_) { MapPoint_ c(kv.second->clone()); mp[kv.second.get()] = c; m->mMapPoints_[kv.first] = c; }
        for (auto& kv : mKeyFrames_) {
            KeyFrame_ c(kv.second->clone());
            auto& v = c->getMapPoints();
            for (auto& p : v) if (p) p = mp[p.get()];
            m->mKeyFrames_[kv.first] = c;
        }
        m->mKeyFrameObs_ = mKeyFrameObs_; m->mMapPointObs_ = mMapPointObs_; m->mGTransformation_ = mGTransformation_;
        m->mCovisibilityGraph_ = mCovisibilityGraph_; m->minCommonObs_ = minCommonObs_;
        return m;
    }
private:
    std::unordered_map<ID, MapPoint_> mMapPoints_;
    std::unordered_map<ID, KeyFrame_> mKeyFrames_;
    std::unordered_map<ID, std::unordered_map<ID, size_t>> mKeyFrameObs_, mMapPointObs_;
    std::unordered_map<ID, std::unordered_map<ID, Sophus::SE3f>> mGTransformation_;
    std::unordered_map<ID, std::unordered_map<ID, int>> mCovisibilityGraph_;
    float minCommonObs_ = 0.f;
};

// ------------------------------------------------------------------ Frame (Mapping/Frame.h:38-233): what poseOnlyOptimization touches
class Frame {
public:
    Frame(const std::vector<cv::KeyPoint>& keys, const Sophus::SE3f& Tcw, std::shared_ptr<CameraModel> calib, int nScales = 8, float scaleFactor = 1.2f)
        : vKeys_(keys), Tcw_(Tcw), calibration_(calib) {
        vMapPoints_.assign(keys.size(), nullptr);
        vInvSigma2_.resize(nScales);
        float s = 1.0f;
        for (int i = 0; i < nScales; ++i) { vInvSigma2_[i] = 1.0f / (s * s); s *= scaleFactor; }
    }
    void setPose(Sophus::SE3f& Tcw) { Tcw_ = Tcw; }
    const Sophus::SE3f getPose() const { return Tcw_; }
    cv::KeyPoint getKeyPoint(const size_t idx) { return vKeys_[idx]; }
    std::vector<std::shared_ptr<MapPoint>>& getMapPoints() { return vMapPoints_; }
    void setMapPoint(size_t idx, std::shared_ptr<MapPoint> pMP) { vMapPoints_[idx] = pMP; }
    std::shared_ptr<CameraModel> getCalibration() { return calibration_; }
    float getInvSigma2(int octave) { return vInvSigma2_[octave]; }
private:
    std::vector<cv::KeyPoint> vKeys_;
    std::vector<MapPoint_> vMapPoints_;
    Sophus::SE3f Tcw_;
    std::shared_ptr<CameraModel> calibration_;
    std::vector<float> vInvSigma2_;
};

// Visualisation is out of scope: the optimisation entry point keeps the parameter and calls update().
class MapVisualizer {
public:
    virtual ~MapVisualizer() {}
    virtual void update(bool /*drawRays*/) {}
};
