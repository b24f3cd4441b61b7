// sophus/se3.hpp — STAND-IN (test infrastructure, not Sophus): rigid transform as rotation matrix + translation with
// the operations ORBmatcher.cc / the shim use.  See Eigen/Core in this directory.
#pragma once
#ifndef SOPHUS_SE3_HPP
#define SOPHUS_SE3_HPP
#include "Eigen/Core"

namespace Sophus {

template <typename T>
struct SO3 {
    static Eigen::Matrix<T, 3, 3> hat(const Eigen::Matrix<T, 3, 1> &v) {
        Eigen::Matrix<T, 3, 3> m;
        m << T(0), -v(2), v(1), v(2), T(0), -v(0), -v(1), v(0), T(0);
        return m;
    }
};
typedef SO3<float> SO3f;

template <typename T>
class SE3 {
public:
    typedef Eigen::Matrix<T, 3, 3> Mat3;
    typedef Eigen::Matrix<T, 3, 1> Vec3;
    SE3() : R_(Mat3::Identity()), t_() {}
    SE3(const Mat3 &R, const Vec3 &t) : R_(R), t_(t) {}
    Mat3 rotationMatrix() const { return R_; }
    const Vec3 &translation() const { return t_; }
    SE3 inverse() const { const Mat3 Rt = R_.transpose(); return SE3(Rt, -(Rt * t_)); }
    Vec3 operator*(const Vec3 &p) const { return R_ * p + t_; }
    SE3 operator*(const SE3 &o) const { return SE3(R_ * o.R_, R_ * o.t_ + t_); }
private:
    Mat3 R_;
    Vec3 t_;
};
typedef SE3<float> SE3f;

}  // namespace Sophus
#endif
