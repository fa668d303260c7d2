// compat_types.h -- minimal stand-ins for the few Eigen / Sophus / OpenCV value types that appear in the reference's
// hot-path API (same names, only the members that API touches).  Used where the real libraries are absent: this image
// (compat.h) and the API-conformance build of tests/refapi (whose <Eigen/Core>, <sophus/se3.hpp>, <opencv2/*.hpp> forward here).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {
template <typename T, int N>
struct VecN {
    T v[N];
    VecN() { for (int i = 0; i < N; ++i) v[i] = T(0); }
    VecN(T a, T b) { static_assert(N == 2, "2"); v[0] = a; v[1] = b; }
    VecN(T a, T b, T c) { static_assert(N == 3, "3"); v[0] = a; v[1] = b; v[2] = c; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    T x() const { return v[0]; }
    T y() const { return v[1]; }
    T z() const { return v[2]; }
    VecN operator+(const VecN& o) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i]; return r; }
    VecN operator-(const VecN& o) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i]; return r; }
    VecN operator*(T s) const { VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] * s; return r; }
    T dot(const VecN& o) const { T s = 0; for (int i = 0; i < N; ++i) s += v[i] * o.v[i]; return s; }
    T squaredNorm() const { return dot(*this); }
    T norm() const { return std::sqrt(squaredNorm()); }
    VecN normalized() const { T n = norm(); VecN r; for (int i = 0; i < N; ++i) r.v[i] = v[i] / n; return r; }
    VecN cross(const VecN& o) const {
        static_assert(N == 3, "3");
        return VecN(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    template <typename U> VecN<U, N> cast() const { VecN<U, N> r; for (int i = 0; i < N; ++i) r.v[i] = U(v[i]); return r; }
    static VecN Zero() { return VecN(); }
};
typedef VecN<float, 2> Vector2f;
typedef VecN<float, 3> Vector3f;
typedef VecN<double, 2> Vector2d;
typedef VecN<double, 3> Vector3d;

struct Matrix3f {
    float m[9];   // row-major
    Matrix3f() { for (int i = 0; i < 9; ++i) m[i] = 0.f; }
    static Matrix3f Identity() { Matrix3f r; r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
    float& operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
    Vector3f operator*(const Vector3f& v) const {
        return Vector3f(m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[3] * v[0] + m[4] * v[1] + m[5] * v[2], m[6] * v[0] + m[7] * v[1] + m[8] * v[2]);
    }
    Matrix3f operator*(const Matrix3f& o) const {
        Matrix3f r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = m[i * 3] * o.m[j] + m[i * 3 + 1] * o.m[3 + j] + m[i * 3 + 2] * o.m[6 + j];
        return r;
    }
    Matrix3f transpose() const { Matrix3f r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = m[j * 3 + i]; return r; }
    void setCol(int c, const Vector3f& v) { m[c] = v[0]; m[3 + c] = v[1]; m[6 + c] = v[2]; }
};
}  // namespace Eigen

namespace Sophus {
// Rigid transform x' = R x + t (the reference stores camera poses as Tcw).
class SE3f {
public:
    SE3f() : R_(Eigen::Matrix3f::Identity()), t_() {}
    SE3f(const Eigen::Matrix3f& R, const Eigen::Vector3f& t) : R_(R), t_(t) {}
    const Eigen::Matrix3f& rotationMatrix() const { return R_; }
    const Eigen::Vector3f& translation() const { return t_; }
    SE3f inverse() const { Eigen::Matrix3f Rt = R_.transpose(); return SE3f(Rt, (Rt * t_) * -1.f); }
    SE3f operator*(const SE3f& o) const { return SE3f(R_ * o.R_, R_ * o.t_ + t_); }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return R_ * p + t_; }
private:
    Eigen::Matrix3f R_;
    Eigen::Vector3f t_;
};
}  // namespace Sophus

namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint {
    Point2f pt; float size; int octave;
    KeyPoint() : size(1.f), octave(0) {}
    KeyPoint(Point2f p, float s, int o = 0) : pt(p), size(s), octave(o) {}
};
}  // namespace cv
