// ORACLE SUPPORT (test infrastructure): compile-only stand-in for Sophus::Sim3 (Thirdparty/Sophus/sophus/sim3.hpp, rxso3.hpp).
// The loop-closing matchers of src/ORBmatcher.cc that take a Sim3 are compiled with the file but never executed by the tests.
#pragma once
#include "se3.hpp"
namespace Sophus {
template <class T, int Opt = 0> class Sim3 {
    SE3<T> se3_; T s_;
public:
    Sim3() : s_(T(1)) {}
    Sim3(const SE3<T>& se3, T s) : se3_(se3), s_(s) {}
    T scale() const { return s_; }
    Matrix3<T> rotationMatrix() const { return se3_.rotationMatrix(); }
    const Vector3<T>& translation() const { return se3_.translation(); }
    Sim3 inverse() const { const SO3<T> invR = se3_.so3().inverse(); return Sim3(SE3<T>(invR, invR * (se3_.translation() * (T(-1) / s_))), T(1) / s_); }
    template <class D> Vector3<T> operator*(const Eigen::MatrixBase<D>& p) const { return s_ * (se3_.so3() * p) + se3_.translation(); }
};
typedef Sim3<float> Sim3f;
typedef Sim3<double> Sim3d;
}  // namespace Sophus
