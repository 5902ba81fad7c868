// ORACLE SUPPORT (test infrastructure, NOT product code): a small stand-in for the parts of Eigen3 (un-vendored dependency of the
// reference, CMakeLists.txt:41 `find_package(Eigen3 3.1.0)`, absent from this image) that the reference's vendored g2o and the
// hot-path functions of src/ORBmatcher.cc, src/Optimizer.cc, src/Frame.cc, src/OptimizableTypes.cpp touch, so that those files
// compile UNMODIFIED from /root/reference (oracle/Makefile target `ref`).  Everything is evaluated eagerly (an expression returns
// a plain Matrix), which also makes every `noalias()` trivially true.
//
// Arithmetic conventions (stated because "what Eigen does" depends on its version and on vectorisation flags):
//  * fixed-size inner products / sums (a.dot(b), squaredNorm(), coefficient of a fixed-size matrix product) are added in the order
//    of Eigen's unrolled scalar reduction (redux_novec_unroller: split the range in halves, recursively): 3 terms -> t0 + (t1 + t2),
//    4 terms -> (t0 + t1) + (t2 + t3); dynamic sizes are summed left to right;
//  * cross(), Quaternion * Quaternion, Quaternion::_transformVector, toRotationMatrix(), Quaternion(Matrix3) and the cofactor
//    inverse of a 3x3 follow Eigen 3.3/3.4's scalar formulas (Geometry/OrthoMethods.h, Geometry/Quaternion.h, LU/InverseImpl.h);
//  * LDLT is Eigen's diagonally pivoted in-place factorisation (Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked) and its solve;
//  * no FMA contraction (the library is built with -ffp-contract=off like the rest of oracle/).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STRONG_INLINE inline
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 7
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_VERSION_AT_LEAST(x, y, z) 1

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Unaligned = 0, Aligned = 16, AlignedBit = 0x80 };
enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 3, Projective = 4 };
enum { EigenvaluesOnly = 0x40, ComputeEigenvectors = 0x80 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum { Lower = 1, Upper = 2 };
inline void initParallel() {}

template <class T> struct aligned_allocator : public std::allocator<T> {
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};

template <class S, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <class Xpr, int BR = Dynamic, int BC = Dynamic> class Block;
template <class M, int MapOpt = 0, class Stride = void> class Map;
template <class Xpr> class DiagonalView;
template <class S> class Quaternion;

namespace internal {
template <class T> struct traits;
template <class S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC>> { typedef S Scalar; enum { Rows = R, Cols = C }; };
template <class X, int BR, int BC> struct traits<Block<X, BR, BC>> { typedef typename traits<X>::Scalar Scalar; enum { Rows = BR, Cols = BC }; };
template <class M, int O, class St> struct traits<Map<M, O, St>> { typedef typename traits<typename std::remove_const<M>::type>::Scalar Scalar; enum { Rows = traits<typename std::remove_const<M>::type>::Rows, Cols = traits<typename std::remove_const<M>::type>::Cols }; };
template <class X> struct traits<DiagonalView<X>> { typedef typename traits<X>::Scalar Scalar; enum { Rows = Dynamic, Cols = 1 }; };
template <int A, int B> struct pick { enum { v = (A != Dynamic) ? A : B }; };
// Eigen's unrolled scalar reduction order for a compile-time length; left-to-right otherwise
template <class S, class F> inline S tree_sum(Index lo, Index n, const F& term) {
    if (n == 1) return term(lo);
    const Index h = n / 2;
    return tree_sum<S>(lo, h, term) + tree_sum<S>(lo + h, n - h, term);
}
template <class S, class F> inline S reduce(Index n, bool fixed, const F& term) {
    if (n == 0) return S(0);
    if (fixed) return tree_sum<S>(0, n, term);
    S s = term(0);
    for (Index i = 1; i < n; ++i) s = s + term(i);
    return s;
}
}  // namespace internal

template <class D> class LLT;
template <class D> class PartialPivLU;
template <class D, int UpLo = Lower> class LDLT;

template <class D> class MatrixBase {
public:
    typedef typename internal::traits<D>::Scalar Scalar;
    typedef Scalar RealScalar;
    enum { RowsAtCompileTime = internal::traits<D>::Rows, ColsAtCompileTime = internal::traits<D>::Cols,
           SizeAtCompileTime = (RowsAtCompileTime == Dynamic || ColsAtCompileTime == Dynamic) ? Dynamic : RowsAtCompileTime * ColsAtCompileTime,
           IsVectorAtCompileTime = (RowsAtCompileTime == 1 || ColsAtCompileTime == 1), Flags = 0 };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
    typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposeReturnType;

    D& derived() { return *static_cast<D*>(this); }
    const D& derived() const { return *static_cast<const D*>(this); }
    Index rows() const { return derived().rows(); }
    Index cols() const { return derived().cols(); }
    Index size() const { return rows() * cols(); }

    Scalar operator()(Index i, Index j) const { return derived().coeff(i, j); }
    Scalar& operator()(Index i, Index j) { return derived().coeffRef(i, j); }
    Scalar coeffLin(Index i) const { return cols() == 1 ? derived().coeff(i, 0) : rows() == 1 ? derived().coeff(0, i) : derived().coeff(i % rows(), i / rows()); }
    Scalar& coeffRefLin(Index i) { return cols() == 1 ? derived().coeffRef(i, 0) : rows() == 1 ? derived().coeffRef(0, i) : derived().coeffRef(i % rows(), i / rows()); }
    Scalar operator()(Index i) const { return coeffLin(i); }
    Scalar& operator()(Index i) { return coeffRefLin(i); }
    Scalar operator[](Index i) const { return coeffLin(i); }
    Scalar& operator[](Index i) { return coeffRefLin(i); }
    Scalar x() const { return coeffLin(0); }  Scalar& x() { return coeffRefLin(0); }
    Scalar y() const { return coeffLin(1); }  Scalar& y() { return coeffRefLin(1); }
    Scalar z() const { return coeffLin(2); }  Scalar& z() { return coeffRefLin(2); }
    Scalar w() const { return coeffLin(3); }  Scalar& w() { return coeffRefLin(3); }
    Scalar value() const { return derived().coeff(0, 0); }

    PlainObject eval() const { return PlainObject(derived()); }
    D& noalias() { return derived(); }
    D& array() { return derived(); }
    const D& array() const { return derived(); }
    D& matrix() { return derived(); }
    const D& matrix() const { return derived(); }
    const D& real() const { return derived(); }

    // ---- assignment-like
    template <class O> D& assign(const MatrixBase<O>& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = o.derived().coeff(i, j);
        return derived();
    }
    template <class O> D& operator+=(const MatrixBase<O>& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = derived().coeff(i, j) + o.derived().coeff(i, j);
        return derived();
    }
    template <class O> D& operator-=(const MatrixBase<O>& o) {
        assert(rows() == o.rows() && cols() == o.cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = derived().coeff(i, j) - o.derived().coeff(i, j);
        return derived();
    }
    D& operator+=(Scalar s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = derived().coeff(i, j) + s; return derived(); }   // array()
    D& operator*=(Scalar s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = derived().coeff(i, j) * s; return derived(); }
    D& operator/=(Scalar s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = derived().coeff(i, j) / s; return derived(); }
    D& setZero() { return setConstant(Scalar(0)); }
    D& setOnes() { return setConstant(Scalar(1)); }
    D& setConstant(Scalar v) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = v; return derived(); }
    void fill(Scalar v) { setConstant(v); }
    D& setIdentity() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0); return derived(); }
    template <class O> void swap(MatrixBase<O>& o) {
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) std::swap(derived().coeffRef(i, j), o.derived().coeffRef(i, j));
    }
    template <class O> void swap(MatrixBase<O>&& o) { swap(o); }

    // ---- comma initialiser
    struct CommaInit {
        D& m; Index k;
        CommaInit(D& m_, Scalar v) : m(m_), k(0) { put(v); }
        void put(Scalar v) { const Index c = m.cols(); m.coeffRef(k / c, k % c) = v; ++k; }      // row by row
        CommaInit& operator,(Scalar v) { put(v); return *this; }
        template <class O> CommaInit& operator,(const MatrixBase<O>& o) { for (Index i = 0; i < o.size(); ++i) put(o[i]); return *this; }
    };
    CommaInit operator<<(Scalar v) { return CommaInit(derived(), v); }
    template <class O> CommaInit operator<<(const MatrixBase<O>& o) { CommaInit c(derived(), o[0]); for (Index i = 1; i < o.size(); ++i) c.put(o[i]); return c; }

    // ---- views
    Block<D> block(Index i, Index j, Index r, Index c) { return Block<D>(derived(), i, j, r, c); }
    const Block<D> block(Index i, Index j, Index r, Index c) const { return Block<D>(const_cast<D&>(derived()), i, j, r, c); }
    template <int BR, int BC> Block<D, BR, BC> block(Index i, Index j) { return Block<D, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC> const Block<D, BR, BC> block(Index i, Index j) const { return Block<D, BR, BC>(const_cast<D&>(derived()), i, j, BR, BC); }
    Block<D, RowsAtCompileTime, 1> col(Index j) { return Block<D, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    const Block<D, RowsAtCompileTime, 1> col(Index j) const { return Block<D, RowsAtCompileTime, 1>(const_cast<D&>(derived()), 0, j, rows(), 1); }
    Block<D, 1, ColsAtCompileTime> row(Index i) { return Block<D, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    const Block<D, 1, ColsAtCompileTime> row(Index i) const { return Block<D, 1, ColsAtCompileTime>(const_cast<D&>(derived()), i, 0, 1, cols()); }
    // vector segments (column or row vector)
    enum { SegR = (ColsAtCompileTime == 1) ? 0 : 1 };
    template <int N> struct Seg { typedef Block<D, (ColsAtCompileTime == 1) ? N : 1, (ColsAtCompileTime == 1) ? 1 : N> type; };
    template <int N> typename Seg<N>::type segment(Index off) { return cols() == 1 ? typename Seg<N>::type(derived(), off, 0, N, 1) : typename Seg<N>::type(derived(), 0, off, 1, N); }
    template <int N> const typename Seg<N>::type segment(Index off) const { return const_cast<MatrixBase*>(this)->template segment<N>(off); }
    typename Seg<Dynamic>::type segment(Index off, Index n) { return cols() == 1 ? typename Seg<Dynamic>::type(derived(), off, 0, n, 1) : typename Seg<Dynamic>::type(derived(), 0, off, 1, n); }
    const typename Seg<Dynamic>::type segment(Index off, Index n) const { return const_cast<MatrixBase*>(this)->segment(off, n); }
    template <int N> typename Seg<N>::type head() { return segment<N>(0); }
    template <int N> const typename Seg<N>::type head() const { return segment<N>(0); }
    template <int N> typename Seg<N>::type tail() { return segment<N>(size() - N); }
    template <int N> const typename Seg<N>::type tail() const { return segment<N>(size() - N); }
    typename Seg<Dynamic>::type head(Index n) { return segment(0, n); }
    const typename Seg<Dynamic>::type head(Index n) const { return segment(0, n); }
    typename Seg<Dynamic>::type tail(Index n) { return segment(size() - n, n); }
    const typename Seg<Dynamic>::type tail(Index n) const { return segment(size() - n, n); }
    template <int BR, int BC> Block<D, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    template <int BR, int BC> const Block<D, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template <int BR, int BC> Block<D, BR, BC> topRightCorner() { return block<BR, BC>(0, cols() - BC); }
    template <int BR, int BC> const Block<D, BR, BC> topRightCorner() const { return block<BR, BC>(0, cols() - BC); }
    template <int BR, int BC> Block<D, BR, BC> bottomLeftCorner() { return block<BR, BC>(rows() - BR, 0); }
    template <int BR, int BC> Block<D, BR, BC> bottomRightCorner() { return block<BR, BC>(rows() - BR, cols() - BC); }
    Block<D> topLeftCorner(Index r, Index c) { return block(0, 0, r, c); }
    const Block<D> topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
    DiagonalView<D> diagonal() { return DiagonalView<D>(derived()); }
    const DiagonalView<D> diagonal() const { return DiagonalView<D>(const_cast<D&>(derived())); }

    // ---- eager expressions
    TransposeReturnType transpose() const {
        TransposeReturnType t; t.resize(cols(), rows());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(j, i) = derived().coeff(i, j);
        return t;
    }
    TransposeReturnType adjoint() const { return transpose(); }
    template <class T> Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
        Matrix<T, RowsAtCompileTime, ColsAtCompileTime> t; t.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(i, j) = static_cast<T>(derived().coeff(i, j));
        return t;
    }
    PlainObject operator-() const {
        PlainObject t; t.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(i, j) = -derived().coeff(i, j);
        return t;
    }
    PlainObject cwiseAbs() const {
        PlainObject t; t.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(i, j) = std::abs(derived().coeff(i, j));
        return t;
    }
    template <class O> PlainObject cwiseProduct(const MatrixBase<O>& o) const {
        PlainObject t; t.resize(rows(), cols());
        for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) t.coeffRef(i, j) = derived().coeff(i, j) * o.derived().coeff(i, j);
        return t;
    }
    // ---- reductions
    Scalar sum() const { return internal::reduce<Scalar>(size(), SizeAtCompileTime != Dynamic, [&](Index i) { return coeffLin(i); }); }
    Scalar trace() const { return internal::reduce<Scalar>(rows(), RowsAtCompileTime != Dynamic, [&](Index i) { return derived().coeff(i, i); }); }
    Scalar squaredNorm() const { return internal::reduce<Scalar>(size(), SizeAtCompileTime != Dynamic, [&](Index i) { const Scalar v = coeffLin(i); return v * v; }); }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    template <class O> Scalar dot(const MatrixBase<O>& o) const {
        assert(size() == o.size());
        return internal::reduce<Scalar>(size(), SizeAtCompileTime != Dynamic || O::SizeAtCompileTime != Dynamic, [&](Index i) { return coeffLin(i) * o.coeffLin(i); });
    }
    Scalar maxCoeff() const { Scalar m = coeffLin(0); for (Index i = 1; i < size(); ++i) if (coeffLin(i) > m) m = coeffLin(i); return m; }
    Scalar minCoeff() const { Scalar m = coeffLin(0); for (Index i = 1; i < size(); ++i) if (coeffLin(i) < m) m = coeffLin(i); return m; }
    template <class I> Scalar maxCoeff(I* idx) const { Scalar m = coeffLin(0); *idx = 0; for (Index i = 1; i < size(); ++i) if (coeffLin(i) > m) { m = coeffLin(i); *idx = (I)i; } return m; }
    bool allFinite() const { for (Index i = 0; i < size(); ++i) if (!std::isfinite(coeffLin(i))) return false; return true; }
    bool isZero(Scalar prec = Scalar(1e-12)) const { for (Index i = 0; i < size(); ++i) if (std::abs(coeffLin(i)) > prec) return false; return true; }
    void normalize() { const Scalar n = norm(); if (n > Scalar(0)) *this /= n; }
    PlainObject normalized() const { PlainObject t(derived()); t.normalize(); return t; }
    template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {        // Geometry/OrthoMethods.h
        Matrix<Scalar, 3, 1> r;
        r.coeffRef(0, 0) = coeffLin(1) * o.coeffLin(2) - coeffLin(2) * o.coeffLin(1);
        r.coeffRef(1, 0) = coeffLin(2) * o.coeffLin(0) - coeffLin(0) * o.coeffLin(2);
        r.coeffRef(2, 0) = coeffLin(0) * o.coeffLin(1) - coeffLin(1) * o.coeffLin(0);
        return r;
    }
    Scalar determinant() const;
    PlainObject inverse() const;
    LLT<PlainObject> llt() const;
    PartialPivLU<PlainObject> lu() const;
    PartialPivLU<PlainObject> partialPivLu() const;
    LDLT<PlainObject> ldlt() const;

    static PlainObject Zero() { PlainObject t; t.setZero(); return t; }
    static PlainObject Zero(Index r, Index c) { PlainObject t; t.resize(r, c); t.setZero(); return t; }
    static PlainObject Zero(Index n) { PlainObject t; t.resize(n); t.setZero(); return t; }
    static PlainObject Ones() { PlainObject t; t.setOnes(); return t; }
    static PlainObject Constant(Scalar v) { PlainObject t; t.setConstant(v); return t; }
    static PlainObject Identity() { PlainObject t; t.setIdentity(); return t; }
    static PlainObject Identity(Index r, Index c) { PlainObject t; t.resize(r, c); t.setIdentity(); return t; }
};

// ---- storage ----------------------------------------------------------------------------------------------------------------
namespace internal {
template <class S, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)> struct Storage;
template <class S, int R, int C> struct Storage<S, R, C, false> {
    S m[R * C > 0 ? R * C : 1];
    Storage() { for (int i = 0; i < R * C; ++i) m[i] = S(0); }          // deterministic (Eigen leaves fixed-size storage uninitialised)
    Index rows() const { return R; } Index cols() const { return C; }
    void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
    S* data() { return m; } const S* data() const { return m; }
};
template <class S, int R, int C> struct Storage<S, R, C, true> {
    std::vector<S> v; Index r_, c_;
    Storage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
    Index rows() const { return r_; } Index cols() const { return c_; }
    void resize(Index r, Index c) { if (r != r_ || c != c_) { r_ = r; c_ = c; v.assign((size_t)(r * c), S(0)); } }
    S* data() { return v.data(); } const S* data() const { return v.data(); }
};
}  // namespace internal

template <class S, int R, int C, int Opt, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt, MR, MC>> {
    internal::Storage<S, R, C> st_;
public:
    typedef MatrixBase<Matrix> Base;
    typedef S Scalar;
    typedef Map<Matrix, Unaligned> MapType;
    typedef Map<const Matrix, Unaligned> ConstMapType;
    typedef Map<Matrix, Aligned> AlignedMapType;
    typedef Map<const Matrix, Aligned> ConstAlignedMapType;
    using Base::operator+=; using Base::operator-=; using Base::operator*=; using Base::operator/=;

    Matrix() {}
    Matrix(const Matrix& o) : st_(o.st_) {}
    template <class O> Matrix(const MatrixBase<O>& o) { st_.resize(o.rows(), o.cols()); Base::assign(o); }
    // size / coefficient constructors (disambiguated like Eigen: fixed-size vectors take coefficients, dynamic take sizes)
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> explicit Matrix(T a) {
        if (R == Dynamic || C == Dynamic) { if (C == 1 || (R == Dynamic && C == Dynamic)) st_.resize((Index)a, C == Dynamic ? 1 : C); else st_.resize(R == Dynamic ? 1 : R, (Index)a); }
        else st_.data()[0] = (S)a;
    }
    template <class T0, class T1, class = typename std::enable_if<std::is_arithmetic<T0>::value && std::is_arithmetic<T1>::value>::type> Matrix(T0 a, T1 b) {
        if (R == Dynamic || C == Dynamic) st_.resize((Index)a, (Index)b);
        else { st_.data()[0] = (S)a; st_.data()[1] = (S)b; }
    }
    Matrix(S a, S b, S c) { S* d = st_.data(); d[0] = a; d[1] = b; d[2] = c; }
    Matrix(S a, S b, S c, S e) { S* d = st_.data(); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    explicit Matrix(const S* p) { std::memcpy(st_.data(), p, sizeof(S) * (size_t)(R * C)); }

    Matrix& operator=(const Matrix& o) { st_ = o.st_; return *this; }
    template <class O> Matrix& operator=(const MatrixBase<O>& o) {
        if ((const void*)&o == (const void*)this) return *this;
        st_.resize(o.rows(), o.cols()); Base::assign(o); return *this;
    }
    Index rows() const { return st_.rows(); }
    Index cols() const { return st_.cols(); }
    void resize(Index r, Index c) { st_.resize(r, c); }
    void resize(Index n) { if (C == 1) st_.resize(n, 1); else if (R == 1) st_.resize(1, n); else st_.resize(n, n); }
    void conservativeResize(Index r, Index c) {
        Matrix t; t.resize(r, c);
        for (Index j = 0; j < std::min(c, cols()); ++j) for (Index i = 0; i < std::min(r, rows()); ++i) t.coeffRef(i, j) = coeff(i, j);
        *this = t;
    }
    S coeff(Index i, Index j) const { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return st_.data()[i + j * st_.rows()]; }
    S& coeffRef(Index i, Index j) { assert(i >= 0 && i < rows() && j >= 0 && j < cols()); return st_.data()[i + j * st_.rows()]; }
    S* data() { return st_.data(); }
    const S* data() const { return st_.data(); }
    Index outerStride() const { return rows(); }
};

// ---- Block: an lvalue view into another expression ------------------------------------------------------------------------------
template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC>> {
    X* x_; Index i0_, j0_, r_, c_;
public:
    typedef MatrixBase<Block> Base;
    typedef typename internal::traits<X>::Scalar Scalar;
    using Base::operator+=; using Base::operator-=; using Base::operator*=; using Base::operator/=;
    Block(X& x, Index i0, Index j0, Index r, Index c) : x_(&x), i0_(i0), j0_(j0), r_(r), c_(c) { assert(i0 >= 0 && j0 >= 0 && i0 + r <= x.rows() && j0 + c <= x.cols()); }
    Block(const Block& o) : x_(o.x_), i0_(o.i0_), j0_(o.j0_), r_(o.r_), c_(o.c_) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Scalar coeff(Index i, Index j) const { return static_cast<const X*>(x_)->coeff(i0_ + i, j0_ + j); }
    Scalar& coeffRef(Index i, Index j) { return x_->coeffRef(i0_ + i, j0_ + j); }
    Scalar& coeffRef(Index i, Index j) const { return x_->coeffRef(i0_ + i, j0_ + j); }
    Block& operator=(const Block& o) { typename Base::PlainObject t(o); Base::assign(t); return *this; }
    template <class O> Block& operator=(const MatrixBase<O>& o) { typename MatrixBase<O>::PlainObject t(o.derived()); Base::assign(t); return *this; }
    void resize(Index r, Index c) { assert(r == r_ && c == c_); (void)r; (void)c; }
};

template <class X>
class DiagonalView : public MatrixBase<DiagonalView<X>> {
    X* x_;
public:
    typedef MatrixBase<DiagonalView> Base;
    typedef typename internal::traits<X>::Scalar Scalar;
    using Base::operator+=; using Base::operator-=; using Base::operator*=; using Base::operator/=;
    explicit DiagonalView(X& x) : x_(&x) {}
    Index rows() const { return std::min(x_->rows(), x_->cols()); }
    Index cols() const { return 1; }
    Scalar coeff(Index i, Index) const { return static_cast<const X*>(x_)->coeff(i, i); }
    Scalar& coeffRef(Index i, Index) { return x_->coeffRef(i, i); }
    Scalar& coeffRef(Index i, Index) const { return x_->coeffRef(i, i); }
    template <class O> DiagonalView& operator=(const MatrixBase<O>& o) { Base::assign(o); return *this; }
    void resize(Index, Index) {}
};

// ---- Map: a column-major view of raw memory -----------------------------------------------------------------------------------
template <class M, int MapOpt, class Stride>
class Map : public MatrixBase<Map<M, MapOpt, Stride>> {
    typedef typename std::remove_const<M>::type Plain;
public:
    typedef MatrixBase<Map> Base;
    typedef typename internal::traits<Plain>::Scalar Scalar;
    using Base::operator+=; using Base::operator-=; using Base::operator*=; using Base::operator/=;
private:
    Scalar* p_; Index r_, c_;
public:
    enum { R = internal::traits<Plain>::Rows, C = internal::traits<Plain>::Cols };
    Map(const Scalar* p) : p_(const_cast<Scalar*>(p)), r_(R), c_(C) {}
    Map(const Scalar* p, Index n) : p_(const_cast<Scalar*>(p)), r_(C == 1 ? n : (R == 1 ? 1 : n)), c_(C == 1 ? 1 : (R == 1 ? n : n)) {}
    Map(const Scalar* p, Index r, Index c) : p_(const_cast<Scalar*>(p)), r_(r), c_(c) {}
    Map(const Map& o) : p_(o.p_), r_(o.r_), c_(o.c_) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Scalar coeff(Index i, Index j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return p_[i + j * r_]; }
    Scalar& coeffRef(Index i, Index j) { return p_[i + j * r_]; }
    Scalar& coeffRef(Index i, Index j) const { return p_[i + j * r_]; }
    Scalar* data() { return p_; }
    const Scalar* data() const { return p_; }
    Map& operator=(const Map& o) { Plain t(o); Base::assign(t); return *this; }
    template <class O> Map& operator=(const MatrixBase<O>& o) { typename MatrixBase<O>::PlainObject t(o.derived()); Base::assign(t); return *this; }
    void resize(Index r, Index c) { assert(r == r_ && c == c_); (void)r; (void)c; }
};
}  // namespace Eigen

// placement new of a Map over an existing Map object (g2o: `new (&_hessian) HessianBlockType(d)`) works through the copy constructor.

namespace Eigen {
// ---- free operators ----------------------------------------------------------------------------------------------------------
#define MINI_EIGEN_RESULT(A, B) Matrix<typename A::Scalar, internal::pick<A::RowsAtCompileTime, B::RowsAtCompileTime>::v, internal::pick<A::ColsAtCompileTime, B::ColsAtCompileTime>::v>
template <class A, class B> MINI_EIGEN_RESULT(A, B) operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    MINI_EIGEN_RESULT(A, B) t; t.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) t.coeffRef(i, j) = a.derived().coeff(i, j) + b.derived().coeff(i, j);
    return t;
}
template <class A, class B> MINI_EIGEN_RESULT(A, B) operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    MINI_EIGEN_RESULT(A, B) t; t.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) t.coeffRef(i, j) = a.derived().coeff(i, j) - b.derived().coeff(i, j);
    return t;
}
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, T s) {
    typename MatrixBase<A>::PlainObject t; t.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) t.coeffRef(i, j) = a.derived().coeff(i, j) * (typename A::Scalar)s;
    return t;
}
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator*(T s, const MatrixBase<A>& a) {
    typename MatrixBase<A>::PlainObject t; t.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) t.coeffRef(i, j) = (typename A::Scalar)s * a.derived().coeff(i, j);
    return t;
}
template <class A, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, T s) {
    typename MatrixBase<A>::PlainObject t; t.resize(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) t.coeffRef(i, j) = a.derived().coeff(i, j) / (typename A::Scalar)s;
    return t;
}
// matrix product: coefficient (i, j) = reduction over k of a(i, k) * b(k, j)
template <class A, class B>
Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    assert(a.cols() == b.rows());
    typedef typename A::Scalar S;
    Matrix<S, A::RowsAtCompileTime, B::ColsAtCompileTime> t; t.resize(a.rows(), b.cols());
    const bool fixed = (A::ColsAtCompileTime != Dynamic) || (B::RowsAtCompileTime != Dynamic);
    const Index K = a.cols();
    for (Index j = 0; j < b.cols(); ++j)
        for (Index i = 0; i < a.rows(); ++i)
            t.coeffRef(i, j) = internal::reduce<S>(K, fixed, [&](Index k) { return a.derived().coeff(i, k) * b.derived().coeff(k, j); });
    return t;
}
template <class A, class B> bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b) {
    if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
    for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) if (a.derived().coeff(i, j) != b.derived().coeff(i, j)) return false;
    return true;
}
template <class A, class B> bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b) { return !(a == b); }
template <class A> std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
    for (Index i = 0; i < a.rows(); ++i) { for (Index j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.derived().coeff(i, j); if (i + 1 < a.rows()) os << "\n"; }
    return os;
}

// ---- determinant / inverse (LU/Determinant.h, LU/InverseImpl.h: closed forms up to 3x3; Gauss-Jordan with partial pivoting above)
template <class D> typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
    const Index n = rows(); assert(n == cols());
    const D& m = derived();
    if (n == 1) return m.coeff(0, 0);
    if (n == 2) return m.coeff(0, 0) * m.coeff(1, 1) - m.coeff(1, 0) * m.coeff(0, 1);
    if (n == 3) {
        auto det3 = [&](int a, int b, int c) { return m.coeff(0, a) * (m.coeff(1, b) * m.coeff(2, c) - m.coeff(1, c) * m.coeff(2, b)); };
        return det3(0, 1, 2) - det3(1, 0, 2) + det3(2, 0, 1);
    }
    PlainObject a(m); Scalar det = 1;                                    // partial-pivot LU
    for (Index k = 0; k < n; ++k) {
        Index p = k; for (Index i = k + 1; i < n; ++i) if (std::abs(a.coeff(i, k)) > std::abs(a.coeff(p, k))) p = i;
        if (a.coeff(p, k) == Scalar(0)) return Scalar(0);
        if (p != k) { for (Index j = 0; j < n; ++j) std::swap(a.coeffRef(k, j), a.coeffRef(p, j)); det = -det; }
        det *= a.coeff(k, k);
        for (Index i = k + 1; i < n; ++i) { const Scalar f = a.coeff(i, k) / a.coeff(k, k); for (Index j = k; j < n; ++j) a.coeffRef(i, j) -= f * a.coeff(k, j); }
    }
    return det;
}
template <class D> typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
    const Index n = rows(); assert(n == cols());
    const D& m = derived();
    PlainObject r; r.resize(n, n);
    if (n == 1) { r.coeffRef(0, 0) = Scalar(1) / m.coeff(0, 0); return r; }
    if (n == 2) {
        const Scalar invdet = Scalar(1) / determinant();
        r.coeffRef(0, 0) = m.coeff(1, 1) * invdet; r.coeffRef(1, 0) = -m.coeff(1, 0) * invdet;
        r.coeffRef(0, 1) = -m.coeff(0, 1) * invdet; r.coeffRef(1, 1) = m.coeff(0, 0) * invdet;
        return r;
    }
    if (n == 3) {                                                        // compute_inverse<Matrix3>: cofactors of column 0, det, then the rest
        auto cof = [&](int i, int j) { const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                                       return m.coeff(i1, j1) * m.coeff(i2, j2) - m.coeff(i1, j2) * m.coeff(i2, j1); };
        const Scalar c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const Scalar det = (c00 * m.coeff(0, 0) + c10 * m.coeff(1, 0)) + c20 * m.coeff(2, 0);
        const Scalar invdet = Scalar(1) / det;
        r.coeffRef(0, 0) = c00 * invdet; r.coeffRef(0, 1) = c10 * invdet; r.coeffRef(0, 2) = c20 * invdet;
        r.coeffRef(1, 0) = cof(0, 1) * invdet; r.coeffRef(1, 1) = cof(1, 1) * invdet; r.coeffRef(1, 2) = cof(2, 1) * invdet;
        r.coeffRef(2, 0) = cof(0, 2) * invdet; r.coeffRef(2, 1) = cof(1, 2) * invdet; r.coeffRef(2, 2) = cof(2, 2) * invdet;
        return r;
    }
    PlainObject a(m); r.setIdentity();
    for (Index k = 0; k < n; ++k) {
        Index p = k; for (Index i = k + 1; i < n; ++i) if (std::abs(a.coeff(i, k)) > std::abs(a.coeff(p, k))) p = i;
        if (p != k) for (Index j = 0; j < n; ++j) { std::swap(a.coeffRef(k, j), a.coeffRef(p, j)); std::swap(r.coeffRef(k, j), r.coeffRef(p, j)); }
        const Scalar d = a.coeff(k, k);
        for (Index j = 0; j < n; ++j) { a.coeffRef(k, j) /= d; r.coeffRef(k, j) /= d; }
        for (Index i = 0; i < n; ++i) if (i != k) { const Scalar f = a.coeff(i, k); if (f != Scalar(0)) for (Index j = 0; j < n; ++j) { a.coeffRef(i, j) -= f * a.coeff(k, j); r.coeffRef(i, j) -= f * r.coeff(k, j); } }
    }
    return r;
}

// ---- Cholesky -----------------------------------------------------------------------------------------------------------------
template <class M> class LLT {
    M L_; bool ok_;
public:
    typedef typename M::Scalar Scalar;
    LLT() : ok_(false) {}
    template <class O> explicit LLT(const MatrixBase<O>& a) { compute(a); }
    template <class O> LLT& compute(const MatrixBase<O>& a) {
        const Index n = a.rows(); L_ = a; ok_ = true;
        for (Index k = 0; k < n; ++k) {
            Scalar x = L_.coeff(k, k);
            for (Index p = 0; p < k; ++p) x -= L_.coeff(k, p) * L_.coeff(k, p);
            if (x <= Scalar(0)) { ok_ = false; return *this; }
            x = std::sqrt(x); L_.coeffRef(k, k) = x;
            for (Index i = k + 1; i < n; ++i) {
                Scalar s = L_.coeff(i, k);
                for (Index p = 0; p < k; ++p) s -= L_.coeff(i, p) * L_.coeff(k, p);
                L_.coeffRef(i, k) = s / x;
            }
        }
        return *this;
    }
    ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
        typename MatrixBase<B>::PlainObject x(b.derived()); const Index n = L_.rows();
        for (Index c = 0; c < x.cols(); ++c) {
            for (Index i = 0; i < n; ++i) { Scalar s = x.coeff(i, c); for (Index p = 0; p < i; ++p) s -= L_.coeff(i, p) * x.coeff(p, c); x.coeffRef(i, c) = s / L_.coeff(i, i); }
            for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, c); for (Index p = i + 1; p < n; ++p) s -= L_.coeff(p, i) * x.coeff(p, c); x.coeffRef(i, c) = s / L_.coeff(i, i); }
        }
        return x;
    }
    const M& matrixL() const { return L_; }
};

template <class M, int UpLo> class LDLT {       // Cholesky/LDLT.h: ldlt_inplace<Lower>::unblocked + solve (_solve_impl)
    M m_; std::vector<Index> tr_; int sign_; bool init_; ComputationInfo info_;
    enum { PositiveSemiDef, NegativeSemiDef, ZeroSign, Indefinite };
public:
    typedef typename M::Scalar Scalar;
    LDLT() : sign_(ZeroSign), init_(false), info_(Success) {}
    template <class O> explicit LDLT(const MatrixBase<O>& a) { compute(a); }
    template <class O> LDLT& compute(const MatrixBase<O>& a) {
        const Index size = a.rows();
        m_ = a; tr_.assign((size_t)size, 0); sign_ = ZeroSign; init_ = true; info_ = Success;
        M& mat = m_;
        if (size <= 1) {
            if (size == 1) { tr_[0] = 0; const Scalar d = mat.coeff(0, 0); sign_ = d > 0 ? PositiveSemiDef : d < 0 ? NegativeSemiDef : ZeroSign; }
            return *this;
        }
        bool found_zero_pivot = false, ret = true;
        std::vector<Scalar> temp((size_t)size);
        for (Index k = 0; k < size; ++k) {
            Index big = k; Scalar best = std::abs(mat.coeff(k, k));
            for (Index i = k + 1; i < size; ++i) if (std::abs(mat.coeff(i, i)) > best) { best = std::abs(mat.coeff(i, i)); big = i; }
            tr_[(size_t)k] = big;
            if (k != big) {                  // symmetric transposition touching the lower triangle only
                const Index s = size - big - 1;
                for (Index j = 0; j < k; ++j) std::swap(mat.coeffRef(k, j), mat.coeffRef(big, j));
                for (Index i = 0; i < s; ++i) std::swap(mat.coeffRef(size - s + i, k), mat.coeffRef(size - s + i, big));
                std::swap(mat.coeffRef(k, k), mat.coeffRef(big, big));
                for (Index i = k + 1; i < big; ++i) { const Scalar tmp = mat.coeff(i, k); mat.coeffRef(i, k) = mat.coeff(big, i); mat.coeffRef(big, i) = tmp; }
            }
            const Index rs = size - k - 1;
            if (k > 0) {
                for (Index j = 0; j < k; ++j) temp[(size_t)j] = mat.coeff(j, j) * mat.coeff(k, j);
                Scalar acc = Scalar(0);
                for (Index j = 0; j < k; ++j) acc = (j == 0) ? mat.coeff(k, 0) * temp[0] : acc + mat.coeff(k, j) * temp[(size_t)j];
                mat.coeffRef(k, k) -= acc;
                for (Index i = 0; i < rs; ++i) {
                    Scalar a2 = mat.coeff(k + 1 + i, 0) * temp[0];
                    for (Index j = 1; j < k; ++j) a2 = a2 + mat.coeff(k + 1 + i, j) * temp[(size_t)j];
                    mat.coeffRef(k + 1 + i, k) -= a2;
                }
            }
            const Scalar akk = mat.coeff(k, k);
            const bool pivot_is_valid = std::abs(akk) > Scalar(0);
            if (k == 0 && !pivot_is_valid) {         // the whole diagonal is zero
                sign_ = ZeroSign;
                for (Index j = 0; j < size; ++j) { tr_[(size_t)j] = j; for (Index i = j + 1; i < size; ++i) ret = ret && (mat.coeff(i, j) == Scalar(0)); }
                info_ = ret ? Success : NumericalIssue; return *this;
            }
            if (rs > 0 && pivot_is_valid) for (Index i = 0; i < rs; ++i) mat.coeffRef(k + 1 + i, k) /= akk;
            else if (rs > 0) for (Index i = 0; i < rs; ++i) ret = ret && (mat.coeff(k + 1 + i, k) == Scalar(0));
            if (found_zero_pivot && pivot_is_valid) ret = false;
            else if (!pivot_is_valid) found_zero_pivot = true;
            if (sign_ == PositiveSemiDef) { if (akk < Scalar(0)) sign_ = Indefinite; }
            else if (sign_ == NegativeSemiDef) { if (akk > Scalar(0)) sign_ = Indefinite; }
            else if (sign_ == ZeroSign) { if (akk > Scalar(0)) sign_ = PositiveSemiDef; else if (akk < Scalar(0)) sign_ = NegativeSemiDef; }
        }
        info_ = ret ? Success : NumericalIssue;
        return *this;
    }
    bool isPositive() const { return sign_ == PositiveSemiDef || sign_ == ZeroSign; }
    bool isNegative() const { return sign_ == NegativeSemiDef || sign_ == ZeroSign; }
    ComputationInfo info() const { return info_; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
        typename MatrixBase<B>::PlainObject x(b.derived()); const Index n = m_.rows();
        const Scalar tol = Scalar(1) / std::numeric_limits<Scalar>::max();
        for (Index c = 0; c < x.cols(); ++c) {
            for (Index k = 0; k < n; ++k) if (tr_[(size_t)k] != k) std::swap(x.coeffRef(k, c), x.coeffRef(tr_[(size_t)k], c));          // P b
            for (Index i = 0; i < n; ++i) { Scalar s = x.coeff(i, c); for (Index p = 0; p < i; ++p) s -= m_.coeff(i, p) * x.coeff(p, c); x.coeffRef(i, c) = s; }
            for (Index i = 0; i < n; ++i) { if (std::abs(m_.coeff(i, i)) > tol) x.coeffRef(i, c) /= m_.coeff(i, i); else x.coeffRef(i, c) = Scalar(0); }
            for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, c); for (Index p = i + 1; p < n; ++p) s -= m_.coeff(p, i) * x.coeff(p, c); x.coeffRef(i, c) = s; }
            for (Index k = n - 1; k >= 0; --k) if (tr_[(size_t)k] != k) std::swap(x.coeffRef(k, c), x.coeffRef(tr_[(size_t)k], c));      // P^T
        }
        return x;
    }
};
// LU with partial pivoting (used off the hot path only: g2o::Sim3::log)
template <class M> class PartialPivLU {
    M a_;
public:
    typedef typename M::Scalar Scalar;
    template <class O> explicit PartialPivLU(const MatrixBase<O>& a) : a_(a) {}
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
        M a(a_); typename MatrixBase<B>::PlainObject x(b.derived()); const Index n = a.rows();
        for (Index k = 0; k < n; ++k) {
            Index p = k; for (Index i = k + 1; i < n; ++i) if (std::abs(a.coeff(i, k)) > std::abs(a.coeff(p, k))) p = i;
            if (p != k) { for (Index j = 0; j < n; ++j) std::swap(a.coeffRef(k, j), a.coeffRef(p, j)); for (Index c = 0; c < x.cols(); ++c) std::swap(x.coeffRef(k, c), x.coeffRef(p, c)); }
            for (Index i = k + 1; i < n; ++i) {
                const Scalar f = a.coeff(i, k) / a.coeff(k, k);
                for (Index j = k; j < n; ++j) a.coeffRef(i, j) -= f * a.coeff(k, j);
                for (Index c = 0; c < x.cols(); ++c) x.coeffRef(i, c) -= f * x.coeff(k, c);
            }
        }
        for (Index c = 0; c < x.cols(); ++c)
            for (Index i = n - 1; i >= 0; --i) { Scalar s = x.coeff(i, c); for (Index j = i + 1; j < n; ++j) s -= a.coeff(i, j) * x.coeff(j, c); x.coeffRef(i, c) = s / a.coeff(i, i); }
        return x;
    }
};
template <class D> PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::lu() const { return PartialPivLU<PlainObject>(derived()); }
template <class D> PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::partialPivLu() const { return PartialPivLU<PlainObject>(derived()); }

// symmetric eigenvalues by cyclic Jacobi rotations (used off the hot path only: OptimizableGraph::verifyInformationMatrices)
template <class M> class SelfAdjointEigenSolver {
    Matrix<typename M::Scalar, Dynamic, 1> ev_;
public:
    typedef typename M::Scalar Scalar;
    SelfAdjointEigenSolver() {}
    template <class O> SelfAdjointEigenSolver& compute(const MatrixBase<O>& m, int = 0) {
        Matrix<Scalar, Dynamic, Dynamic> a(m.derived()); const Index n = a.rows();
        for (int sweep = 0; sweep < 64; ++sweep) {
            Scalar off = 0; for (Index i = 0; i < n; ++i) for (Index j = 0; j < i; ++j) off += a.coeff(i, j) * a.coeff(i, j);
            if (off < std::numeric_limits<Scalar>::min()) break;
            for (Index p = 0; p < n; ++p) for (Index q = p + 1; q < n; ++q) {
                if (a.coeff(p, q) == Scalar(0)) continue;
                const Scalar th = (a.coeff(q, q) - a.coeff(p, p)) / (Scalar(2) * a.coeff(p, q));
                const Scalar t = (th >= 0 ? Scalar(1) : Scalar(-1)) / (std::abs(th) + std::sqrt(th * th + Scalar(1)));
                const Scalar c = Scalar(1) / std::sqrt(t * t + Scalar(1)), s2 = t * c;
                for (Index k = 0; k < n; ++k) { const Scalar akp = a.coeff(k, p), akq = a.coeff(k, q); a.coeffRef(k, p) = c * akp - s2 * akq; a.coeffRef(k, q) = s2 * akp + c * akq; }
                for (Index k = 0; k < n; ++k) { const Scalar apk = a.coeff(p, k), aqk = a.coeff(q, k); a.coeffRef(p, k) = c * apk - s2 * aqk; a.coeffRef(q, k) = s2 * apk + c * aqk; }
            }
        }
        ev_.resize(n); for (Index i = 0; i < n; ++i) ev_[i] = a.coeff(i, i);
        std::sort(ev_.data(), ev_.data() + n);
        return *this;
    }
    const Matrix<Scalar, Dynamic, 1>& eigenvalues() const { return ev_; }
};

template <class D> LLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::llt() const { return LLT<PlainObject>(derived()); }
template <class D> LDLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::ldlt() const { return LDLT<PlainObject>(derived()); }

// ---- typedefs -----------------------------------------------------------------------------------------------------------------
#define MINI_EIGEN_TYPEDEFS(T, sfx) \
    typedef Matrix<T, 2, 1> Vector2##sfx; typedef Matrix<T, 3, 1> Vector3##sfx; typedef Matrix<T, 4, 1> Vector4##sfx; typedef Matrix<T, Dynamic, 1> VectorX##sfx; \
    typedef Matrix<T, 1, 2> RowVector2##sfx; typedef Matrix<T, 1, 3> RowVector3##sfx; typedef Matrix<T, 1, 4> RowVector4##sfx; typedef Matrix<T, 1, Dynamic> RowVectorX##sfx; \
    typedef Matrix<T, 2, 2> Matrix2##sfx; typedef Matrix<T, 3, 3> Matrix3##sfx; typedef Matrix<T, 4, 4> Matrix4##sfx; typedef Matrix<T, Dynamic, Dynamic> MatrixX##sfx;
MINI_EIGEN_TYPEDEFS(float, f) MINI_EIGEN_TYPEDEFS(double, d) MINI_EIGEN_TYPEDEFS(int, i)

// ---- Geometry -----------------------------------------------------------------------------------------------------------------
template <class S> class AngleAxis;
template <class S> class Quaternion {
    Matrix<S, 4, 1> c_;          // x, y, z, w
public:
    typedef S Scalar;
    typedef Matrix<S, 4, 1> Coefficients;
    typedef Matrix<S, 3, 1> Vector3;
    typedef Matrix<S, 3, 3> Matrix3;
    Quaternion() {}
    Quaternion(S w, S x, S y, S z) { c_[0] = x; c_[1] = y; c_[2] = z; c_[3] = w; }
    explicit Quaternion(const S* d) { for (int i = 0; i < 4; ++i) c_[i] = d[i]; }
    template <class O> explicit Quaternion(const MatrixBase<O>& m) { *this = m; }
    Quaternion(const Quaternion& o) : c_(o.c_) {}
    explicit Quaternion(const AngleAxis<S>& aa);
    Quaternion& operator=(const Quaternion& o) { c_ = o.c_; return *this; }
    template <class O> Quaternion& operator=(const MatrixBase<O>& m) {
        if (m.rows() == 4 && m.cols() == 1) { for (int i = 0; i < 4; ++i) c_[i] = m[i]; return *this; }
        assert(m.rows() == 3 && m.cols() == 3);                           // Geometry/Quaternion.h quaternionbase_assign_impl<Other,3,3>
        const O& mat = m.derived();
        S t = mat.trace();
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
            y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
            z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            Index i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const Index j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + S(1.0));
            c_[i] = S(0.5) * t;
            t = S(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            c_[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            c_[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
        return *this;
    }
    S x() const { return c_[0]; } S y() const { return c_[1]; } S z() const { return c_[2]; } S w() const { return c_[3]; }
    S& x() { return c_[0]; } S& y() { return c_[1]; } S& z() { return c_[2]; } S& w() { return c_[3]; }
    Coefficients& coeffs() { return c_; }
    const Coefficients& coeffs() const { return c_; }
    Vector3 vec() const { return Vector3(c_[0], c_[1], c_[2]); }
    Quaternion& setIdentity() { c_[0] = c_[1] = c_[2] = S(0); c_[3] = S(1); return *this; }
    static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
    S squaredNorm() const { return c_.squaredNorm(); }
    S norm() const { return c_.norm(); }
    void normalize() { c_ /= norm(); }                                    // MatrixBase::normalize (Eigen 3.3 guards z > 0)
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const { const S n2 = squaredNorm(); Quaternion q = conjugate(); q.c_ /= n2; return q; }
    Quaternion operator*(const Quaternion& b) const {                     // quat_product<Arch, Derived1, Derived2, Scalar>
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    template <class O> Vector3 _transformVector(const MatrixBase<O>& v) const {
        Vector3 uv = vec().cross(v);
        uv += uv;
        return Vector3(v) + w() * uv + vec().cross(uv);                   // ((v + w*uv) + vec x uv)
    }
    template <class O> Vector3 operator*(const MatrixBase<O>& v) const { return _transformVector(v); }
    Matrix3 toRotationMatrix() const {                                    // QuaternionBase::toRotationMatrix
        Matrix3 res;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w();
        const S txx = tx * x(), txy = ty * x(), txz = tz * x();
        const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.coeffRef(0, 0) = S(1) - (tyy + tzz); res.coeffRef(0, 1) = txy - twz; res.coeffRef(0, 2) = txz + twy;
        res.coeffRef(1, 0) = txy + twz; res.coeffRef(1, 1) = S(1) - (txx + tzz); res.coeffRef(1, 2) = tyz - twx;
        res.coeffRef(2, 0) = txz - twy; res.coeffRef(2, 1) = tyz + twx; res.coeffRef(2, 2) = S(1) - (txx + tyy);
        return res;
    }
    Matrix3 matrix() const { return toRotationMatrix(); }
    template <class T> Quaternion<T> cast() const { return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z()); }
    S dot(const Quaternion& o) const { return c_.dot(o.c_); }
    S angularDistance(const Quaternion& o) const { const Quaternion d = *this * o.conjugate(); return S(2) * std::atan2(d.vec().norm(), std::abs(d.w())); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S> class AngleAxis {
    Matrix<S, 3, 1> axis_; S angle_;
public:
    AngleAxis() : angle_(0) {}
    template <class O> AngleAxis(S angle, const MatrixBase<O>& axis) : axis_(axis), angle_(angle) {}
    S angle() const { return angle_; }
    const Matrix<S, 3, 1>& axis() const { return axis_; }
    Matrix<S, 3, 3> toRotationMatrix() const { return Quaternion<S>(*this).toRotationMatrix(); }
};
template <class S> Quaternion<S>::Quaternion(const AngleAxis<S>& aa) {
    const S ha = S(0.5) * aa.angle();
    w() = std::cos(ha); const S s = std::sin(ha);
    x() = s * aa.axis()[0]; y() = s * aa.axis()[1]; z() = s * aa.axis()[2];
}
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;

// Transform: only what g2o's type headers mention (conversion operators that the hot path never executes)
template <class S, int Dim, int Mode, int Opt = 0> class Transform {
    Matrix<S, Dim + 1, Dim + 1> m_;
public:
    Transform() { m_.setIdentity(); }
    Transform(const Quaternion<S>& q) { m_.setIdentity(); m_.template block<3, 3>(0, 0) = q.toRotationMatrix(); }
    template <class O> Transform(const MatrixBase<O>& m) { m_ = m; }
    Transform& operator=(const Quaternion<S>& q) { m_.setIdentity(); m_.template block<3, 3>(0, 0) = q.toRotationMatrix(); return *this; }
    static Transform Identity() { return Transform(); }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
    const Block<Matrix<S, Dim + 1, Dim + 1>, Dim, 1> translation() const { return m_.template block<Dim, 1>(0, Dim); }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
    const Block<Matrix<S, Dim + 1, Dim + 1>, Dim, Dim> linear() const { return m_.template block<Dim, Dim>(0, 0); }
    Matrix<S, Dim, Dim> rotation() const { return Matrix<S, Dim, Dim>(linear()); }
    Matrix<S, Dim + 1, Dim + 1>& matrix() { return m_; }
    const Matrix<S, Dim + 1, Dim + 1>& matrix() const { return m_; }
    Transform operator*(const Transform& o) const { Transform t; t.m_ = m_ * o.m_; return t; }
    Matrix<S, Dim, 1> operator*(const Matrix<S, Dim, 1>& p) const { return Matrix<S, Dim, 1>(linear() * p + translation()); }
    Transform inverse() const { Transform t; t.m_ = m_.inverse(); return t; }
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;
typedef Transform<float, 3, Isometry> Isometry3f;

template <class S> class Rotation2D {
    S a_;
public:
    Rotation2D(S a = 0) : a_(a) {}
    S angle() const { return a_; }
    Matrix<S, 2, 2> toRotationMatrix() const { Matrix<S, 2, 2> m; const S c = std::cos(a_), s = std::sin(a_); m << c, -s, s, c; return m; }
};
typedef Rotation2D<double> Rotation2Dd;

}  // namespace Eigen
