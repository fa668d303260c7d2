// realpair_main.cc -- the real-image flow after matching (Mapping::monocularMapInitialization, Mapping.cc:159-257,
// then SLAM::processImage -> deformationOptimization / arapOptimization, SLAM.cc:108-131) on top of the host shim.
//   dsc_realpair <settings.yaml> <pair.txt> <depth1.f32> <depth2.f32>
// pair.txt: "n1 n2 W H imageDepthScale1 imageDepthScale2", 8 camera parameters, 12 + 12 pose entries (row-major
// 3x4 Tcw), n1 lines "x y octave", n2 lines "x y octave", n1 match indices (-1 = unmatched).
#include <cstdio>
#include <fstream>
#include <iostream>

#include "Optimization.h"

static std::vector<float> readF32(const std::string& path, size_t count) {
    std::vector<float> v(count);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(float)));
    if (!f) throw std::runtime_error("cannot read " + path);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s settings.yaml pair.txt depth1.f32 depth2.f32\n", argv[0]); return 2; }
    try {
        Settings settings(argv[1]);
        std::ifstream in(argv[2]);
        size_t n1, n2; int W, H; double ids1, ids2;
        in >> n1 >> n2 >> W >> H >> ids1 >> ids2;
        std::vector<float> cam(8);
        for (auto& c : cam) in >> c;
        auto readPose = [&]() {
            float m[12];
            for (float& v : m) in >> v;
            Eigen::Matrix3f R;
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = m[r * 4 + c];
            return Sophus::SE3f(R, Eigen::Vector3f(m[3], m[7], m[11]));
        };
        Sophus::SE3f T1w = readPose(), T2w = readPose();
        auto readKeys = [&](size_t n) {
            std::vector<cv::KeyPoint> k(n);
            for (auto& kp : k) { float x, y; int o; in >> x >> y >> o; kp = cv::KeyPoint(cv::Point2f(x, y), 1.f, o); }
            return k;
        };
        auto k1 = readKeys(n1), k2 = readKeys(n2);
        std::vector<int> matches(n1);
        for (auto& m : matches) in >> m;
        if (!in) throw std::runtime_error("pair file truncated");
        auto calib = std::make_shared<KannalaBrandt8>(cam);
        auto refKF = std::make_shared<KeyFrame>(k1, T1w, calib), currKF = std::make_shared<KeyFrame>(k2, T2w, calib);
        refKF->setDepthImage(readF32(argv[3], (size_t)W * H), W, H, ids1);
        currKF->setDepthImage(readF32(argv[4], (size_t)W * H), W, H, ids2);
        auto pMap = std::make_shared<Map>();
        pMap->insertKeyFrame(refKF); pMap->insertKeyFrame(currKF);
        float parallax = 0.f;
        int created = dsc_host::initializeMapFromMatches(*pMap, refKF, currKF, matches, settings, &parallax);
        double s1_0 = refKF->getEstimatedDepthScale(), s2_0 = currKF->getEstimatedDepthScale();
        std::vector<int> slots;
        std::vector<Eigen::Vector3f> t1, t2;
        for (size_t i = 0; i < n1; ++i) if (refKF->getMapPoint(i)) { slots.push_back((int)i); t1.push_back(refKF->getMapPoint(i)->getWorldPosition()); t2.push_back(currKF->getMapPoint(i)->getWorldPosition()); }
        dsc_host::setSolver(1e-12, 20000);
        double update = 0;
        arapOptimization(pMap.get(), settings.getOptRepWeight(), settings.getOptGlobalWeight(), settings.getOptArapWeight(), settings.getOptAlphaWeight(),
                         settings.getOptBetaWeight(), settings.getSimulatedDepthWeight() / 1000, settings.getnOptIterations(), &update);
        std::printf("{\"created\": %d, \"parallax\": %.9g, \"s1_init\": %.17g, \"s2_init\": %.17g, \"s1\": %.17g, \"s2\": %.17g, \"update\": %.17g,\n \"slots\": [",
                    created, parallax, s1_0, s2_0, refKF->getEstimatedDepthScale(), currKF->getEstimatedDepthScale(), update);
        for (size_t i = 0; i < slots.size(); ++i) std::printf("%s%d", i ? ", " : "", slots[i]);
        std::printf("],\n \"trace\": [");
        auto& tr = dsc_host::lastTrace();
        for (size_t i = 0; i < tr.size(); ++i) std::printf("%s[%.17g, %.17g, %d]", i ? ", " : "", tr[i].chi2_before, tr[i].lambda, tr[i].trials);
        auto dump = [&](const char* name, const std::vector<Eigen::Vector3f>& v) {
            std::printf("],\n \"%s\": [", name);
            for (size_t i = 0; i < v.size(); ++i) std::printf("%s%.9g, %.9g, %.9g", i ? ", " : "", v[i][0], v[i][1], v[i][2]);
        };
        dump("tri1", t1); dump("tri2", t2);
        std::vector<Eigen::Vector3f> f1, f2;
        for (int s : slots) { f1.push_back(refKF->getMapPoint((size_t)s)->getWorldPosition()); f2.push_back(currKF->getMapPoint((size_t)s)->getWorldPosition()); }
        dump("X1", f1); dump("X2", f2);
        std::printf("]}\n");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "dsc_realpair: %s\n", e.what());
        return 1;
    }
    return 0;
}
