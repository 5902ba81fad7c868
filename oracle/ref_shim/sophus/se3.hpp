// ORACLE SUPPORT (test infrastructure, NOT product code): stand-in for the reference's vendored Sophus (Thirdparty/Sophus), which
// is written against Eigen internals (traits / Map specialisations) that the stand-in Eigen does not have.  Only the operations
// the hot path executes are restated, each after the vendored source (file:line under Thirdparty/Sophus/sophus/):
//   SO3(quaternion) normalises: coeffs /= norm            so3.hpp:481-487, 297-303
//   SO3(R) = Quaternion(R), no normalisation               so3.hpp:469-474
//   SO3::inverse() = SO3(conjugate) (normalises again)     so3.hpp:229-231
//   SO3 * SO3 explicit Hamilton product -> SO3(quaternion) so3.hpp:325-339 (the product is handed to the normalising constructor, :481-487)
//   SO3 * point = p + w*uv + vec x uv, uv = 2 (vec x p)    so3.hpp:358-367
//   SE3(q, t), SE3(R, t), SE3(so3, t)                      se3.hpp:466-490
//   SE3::inverse() = (invR, invR * (t * -1))               se3.hpp:208-211
//   SE3 * SE3 = (R1 R2, t1 + R1 t2)                        se3.hpp:304-308
//   SE3 * point = so3 * p + t                              se3.hpp:321-324
//   rotationMatrix() = quaternion.toRotationMatrix()       se3.hpp:363, so3.hpp:310-312
//   SO3::hat                                               so3.hpp:694-704
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>

namespace Sophus {
template <class T> using Vector3 = Eigen::Matrix<T, 3, 1>;
template <class T> using Matrix3 = Eigen::Matrix<T, 3, 3>;
template <class T> using Matrix4 = Eigen::Matrix<T, 4, 4>;

template <class T, int Opt = 0> class SO3 {
    Eigen::Quaternion<T> q_;
    struct Raw {};
    SO3(const Eigen::Quaternion<T>& q, Raw) : q_(q) {}
public:
    typedef T Scalar;
    SO3() : q_(T(1), T(0), T(0), T(0)) {}
    SO3(const Matrix3<T>& R) : q_(R) {}
    explicit SO3(const Eigen::Quaternion<T>& q) : q_(q) { normalize(); }
    void normalize() { const T length = q_.norm(); q_.coeffs() /= length; }
    const Eigen::Quaternion<T>& unit_quaternion() const { return q_; }
    void setQuaternion(const Eigen::Quaternion<T>& q) { q_ = q; normalize(); }
    Matrix3<T> matrix() const { return q_.toRotationMatrix(); }
    SO3 inverse() const { return SO3(q_.conjugate()); }
    SO3 operator*(const SO3& o) const {
        const Eigen::Quaternion<T>& a = q_; const Eigen::Quaternion<T>& b = o.q_;
        return SO3(Eigen::Quaternion<T>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                        a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                        a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                        a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));
    }
    template <class D> Vector3<T> operator*(const Eigen::MatrixBase<D>& p) const {
        const Vector3<T> qv = q_.vec();
        Vector3<T> uv = qv.cross(p);
        uv += uv;
        return Vector3<T>(p) + q_.w() * uv + qv.cross(uv);
    }
    static Matrix3<T> hat(const Vector3<T>& omega) {
        Matrix3<T> Omega;
        Omega << T(0), -omega(2), omega(1), omega(2), T(0), -omega(0), -omega(1), omega(0), T(0);
        return Omega;
    }
    template <class U> SO3<U> cast() const { return SO3<U>(q_.template cast<U>()); }
};
typedef SO3<float> SO3f;
typedef SO3<double> SO3d;

template <class T, int Opt = 0> class SE3 {
    SO3<T> so3_; Vector3<T> t_;
public:
    typedef T Scalar;
    SE3() { t_.setZero(); }
    template <class D> SE3(const SO3<T>& so3, const Eigen::MatrixBase<D>& t) : so3_(so3), t_(t) {}
    template <class D> SE3(const Matrix3<T>& R, const Eigen::MatrixBase<D>& t) : so3_(R), t_(t) {}
    template <class D> SE3(const Eigen::Quaternion<T>& q, const Eigen::MatrixBase<D>& t) : so3_(q), t_(t) {}
    explicit SE3(const Matrix4<T>& M) : so3_(Matrix3<T>(M.template topLeftCorner<3, 3>())), t_(M.template block<3, 1>(0, 3)) {}
    SO3<T>& so3() { return so3_; }
    const SO3<T>& so3() const { return so3_; }
    Vector3<T>& translation() { return t_; }
    const Vector3<T>& translation() const { return t_; }
    const Eigen::Quaternion<T>& unit_quaternion() const { return so3_.unit_quaternion(); }
    Matrix3<T> rotationMatrix() const { return so3_.matrix(); }
    void setQuaternion(const Eigen::Quaternion<T>& q) { so3_.setQuaternion(q); }
    SE3 inverse() const { const SO3<T> invR = so3_.inverse(); return SE3(invR, invR * (t_ * T(-1))); }
    SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, t_ + so3_ * o.t_); }
    SE3& operator*=(const SE3& o) { *this = *this * o; return *this; }
    template <class D> Vector3<T> operator*(const Eigen::MatrixBase<D>& p) const { return so3_ * p + t_; }
    Eigen::Matrix<T, 3, 4> matrix3x4() const { Eigen::Matrix<T, 3, 4> m; m.template topLeftCorner<3, 3>() = rotationMatrix(); m.col(3) = t_; return m; }
    Matrix4<T> matrix() const { Matrix4<T> m; m.setIdentity(); m.template topLeftCorner<3, 3>() = rotationMatrix(); m.template block<3, 1>(0, 3) = t_; return m; }
    template <class U> SE3<U> cast() const { return SE3<U>(so3_.template cast<U>(), t_.template cast<U>()); }
};
typedef SE3<float> SE3f;
typedef SE3<double> SE3d;
}  // namespace Sophus
