// Settings.h -- flat "%YAML:1.0" reader for the hot-path keys of Modules/System/Settings.cc:27-190.
// Like cv::FileStorage in the reference, a missing key silently becomes 0 / "".
#pragma once
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>

#include "Map.h"

class Settings {
public:
    Settings() {}
    explicit Settings(const std::string& configFile) { load(configFile); }
    bool load(const std::string& configFile) {
        std::ifstream f(configFile);
        if (!f.is_open()) return false;
        std::string line;
        while (std::getline(f, line)) {
            size_t h = line.find('#');
            if (h != std::string::npos) line = line.substr(0, h);
            size_t c = line.find(':');
            if (c == std::string::npos || line[0] == '%') continue;
            std::string k = trim(line.substr(0, c)), v = trim(line.substr(c + 1));
            if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
            if (!k.empty()) kv_[k] = v;
        }
        std::vector<float> cal = {num("Camera.fx"), num("Camera.fy"), num("Camera.cx"), num("Camera.cy"),
                                  num("Camera.d0"), num("Camera.d1"), num("Camera.d2"), num("Camera.d3")};
        calibration_ = std::make_shared<KannalaBrandt8>(cal);       // Settings.cc:43-51: always KB8 on the hot path
        pinHolecalibration_ = std::make_shared<PinHole>(std::vector<float>(cal.begin(), cal.begin() + 4));
        return true;
    }
    void set(const std::string& k, const std::string& v) { kv_[k] = v; }
    float num(const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? 0.f : (float)atof(it->second.c_str()); }
    double dnum(const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? 0.0 : atof(it->second.c_str()); }
    std::string str(const std::string& k) const { auto it = kv_.find(k); return it == kv_.end() ? std::string() : it->second; }

    std::shared_ptr<CameraModel> getCalibration() { return calibration_; }
    std::shared_ptr<CameraModel> getPHCalibration() { return pinHolecalibration_; }
    float getMinCos() { return num("Triangulation.minCos"); }
    float getDepthLimit() { return num("Triangulation.depthLimit"); }
    bool getCheckingSelection() { return str("Triangulation.checks") == "true"; }
    std::string getTrianMethod() { return str("Triangulation.method"); }
    std::string getTrianLocation() { return str("Triangulation.seed.location"); }
    float getSimulatedRepError() { return num("Keypoints.RepError"); }
    int getDecimalsRepError() { return (int)num("Keypoints.decimalsApproximation"); }
    float getSimulatedDepthError() { return num("Measurements.DepthError"); }
    float getSimulatedDepthWeight() { return num("Measurements.DepthWeight"); }
    float getSimulatedDepthScaleC1() { return num("Measurements.DepthScale.C1"); }
    float getSimulatedDepthScaleC2() { return num("Measurements.DepthScale.C2"); }
    double getOptRepWeight() { return dnum("Optimization.rep"); }
    double getOptArapWeight() { return dnum("Optimization.arap"); }
    double getOptGlobalWeight() { return dnum("Optimization.global"); }
    double getOptAlphaWeight() { return dnum("Optimization.alpha"); }
    double getOptBetaWeight() { return dnum("Optimization.beta"); }
    std::string getOptSelection() { return str("Optimization.selection"); }
    std::string getOptWeightsSelection() { return str("Optimization.weightsSelection"); }
    int getnOptimizations() { return (int)num("Optimization.numberOfOptimizations"); }
    int getnOptIterations() { return (int)num("Optimization.numberOfIterations"); }
    int getNloptnOptimizations() { return (int)num("Optimization.nlopt.numberOfIterations"); }
    double getNloptRelTolerance() { return dnum("Optimization.nlopt.relTolerance"); }
    double getNloptAbsTolerance() { return dnum("Optimization.nlopt.absTolerance"); }
    double getNloptRepLowerBound() { return dnum("Optimization.nlopt.rep.lowerBound"); }
    double getNloptRepUpperBound() { return dnum("Optimization.nlopt.rep.upperBound"); }
    double getNloptGlobalLowerBound() { return dnum("Optimization.nlopt.global.lowerBound"); }
    double getNloptGlobalUpperBound() { return dnum("Optimization.nlopt.global.upperBound"); }
    double getNloptArapLowerBound() { return dnum("Optimization.nlopt.arap.lowerBound"); }
    double getNloptArapUpperBound() { return dnum("Optimization.nlopt.arap.upperBound"); }
    bool getDrawRaysSelection() { return str("MapVisualizer.drawRays") == "true"; }
    std::string getExpFilePath() { return str("Experiment.Filepath"); }
    Eigen::Vector3f getFirstCameraPos() { return Eigen::Vector3f(num("Camera.FirstPose.x"), num("Camera.FirstPose.y"), num("Camera.FirstPose.z")); }
    Eigen::Vector3f getSecondCameraPos() { return Eigen::Vector3f(num("Camera.SecondPose.x"), num("Camera.SecondPose.y"), num("Camera.SecondPose.z")); }

private:
    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    std::map<std::string, std::string> kv_;
    std::shared_ptr<CameraModel> calibration_, pinHolecalibration_;
};
