// sophus/sim3.hpp — STAND-IN (test infrastructure, not Sophus): similarity transform s * R * p + t.
#pragma once
#ifndef SOPHUS_SIM3_HPP
#define SOPHUS_SIM3_HPP
#include "sophus/se3.hpp"

namespace Sophus {

template <typename T>
class Sim3 {
public:
    typedef Eigen::Matrix<T, 3, 3> Mat3;
    typedef Eigen::Matrix<T, 3, 1> Vec3;
    Sim3() : s_(1), R_(Mat3::Identity()), t_() {}
    Sim3(T s, const Mat3 &R, const Vec3 &t) : s_(s), R_(R), t_(t) {}
    Mat3 rotationMatrix() const { return R_; }
    const Vec3 &translation() const { return t_; }
    T scale() const { return s_; }
    Sim3 inverse() const { const Mat3 Rt = R_.transpose(); const T is = T(1) / s_; return Sim3(is, Rt, -((Rt * t_) * is)); }
    Vec3 operator*(const Vec3 &p) const { return (R_ * p) * s_ + t_; }
private:
    T s_;
    Mat3 R_;
    Vec3 t_;
};
typedef Sim3<float> Sim3f;

}  // namespace Sophus
#endif
