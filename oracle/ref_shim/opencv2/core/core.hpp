// TEST INFRASTRUCTURE.  Stand-in for the OpenCV headers, just large enough for two pieces of the REFERENCE to compile
// UNMODIFIED from /root/reference into oracle/_ref/ (oracle/Makefile, target `ref`):
//   * Thirdparty/DBoW2 (vocabulary tree, transform)            -> libref_dbow2.so
//   * src/ORBextractor.cc (ORBextractor, DistributeOctTree)    -> libref_orbextractor.so
// What is real and what is substituted: every line of control flow, table construction, cell geometry, quad-tree (std::list,
// std::sort), orientation and descriptor code that runs is the reference's.  The OpenCV *primitives* it calls are declared
// here and defined in oracle/ref_orbextractor_driver.cpp on top of the oracle's restatements (cv::resize INTER_LINEAR,
// cv::GaussianBlur 7x7, cv::FAST, cv::fastAtan2), each of which is pinned against the real OpenCV (python-cv2) by
// tests/test_oracle_golden.py.  cv::Mat here is a reference-counted byte/float matrix with views (rowRange / colRange /
// operator()(Rect) share storage, create() keeps a matching allocation), which is the part of its semantics the reference
// relies on (ComputePyramid resizes INTO a view of the bordered image).
#pragma once
// the real header pulls these in, and the reference relies on it
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_16U 2
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

namespace cv {

inline int cvRound(double v) { return (int)lrint(v); }          // round half to even, like OpenCV's SSE2 / lrint paths
inline int cvFloor(double v) { return (int)floor(v); }
inline int cvCeil(double v) { return (int)ceil(v); }

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }      // Point2f *= float: float arithmetic
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };

class KeyPoint {
public:
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

class Mat {
public:
    int rows, cols; size_t step; uchar* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr), type_(0) { create(sz.height, sz.width, type); }
    Mat(Size sz, int type, void* ext) : rows(sz.height), cols(sz.width), step(0), data((uchar*)ext), type_(type) { step = (size_t)cols * elemSize(); }   // user data, not owned
    struct Expr;                                                 // result of an arithmetic expression (stand-in for cv::MatExpr)
    Mat(const Expr& e);
    Mat& operator=(const Expr& e);                               // evaluates INTO a matching allocation (row views!), like MatExpr assignment
    Expr mul(const Mat& m) const;
    void convertTo(Mat& dst, int type) const;
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : (type_ == CV_16U ? 2 : 1); }
    size_t step1() const { return step / elemSize(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    void create(int r, int c, int type) {                       // keeps a matching allocation (OpenCV semantics)
        if (data && rows == r && cols == c && type_ == type) return;
        type_ = type; rows = r; cols = c; step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step + 64, (uchar)0);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }
    static Mat zeros(Size sz, int type) { Mat m(sz, type); return m; }
    static Mat ones(int r, int c, int type);
    Mat clone() const { Mat m; if (!empty()) { m.create(rows, cols, type_); for (int y = 0; y < rows; ++y) memcpy(m.data + y * m.step, data + y * step, cols * elemSize()); } return m; }
    void copyTo(Mat dst) const { dst.create(rows, cols, type_); for (int y = 0; y < rows; ++y) memmove(dst.data + y * dst.step, data + y * step, cols * elemSize()); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    template <class T> T& at(int y, int x) { return *reinterpret_cast<T*>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <class T> const T& at(int y, int x) const { return *reinterpret_cast<const T*>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

// cv::Mat_<T>(r, c) << a, b, ...  (row-major comma initialiser; Pinhole::toK)
template <class T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_8U), k_(0) {}
    Mat_& operator<<(T v) { return (*this, v); }
    Mat_& operator,(T v) { at<T>(k_ / cols, k_ % cols) = v; ++k_; return *this; }
private:
    int k_;
};
enum { NORM_L1 = 2, NORM_HAMMING = 6 };
inline double norm(const Mat& a, const Mat& b, int normType) {      // NORM_L1 on CV_8U windows (Frame::ComputeStereoMatches)
    assert(normType == NORM_L1 && a.type() == CV_8U && a.rows == b.rows && a.cols == b.cols); (void)normType;
    int s = 0;
    for (int y = 0; y < a.rows; ++y) { const uchar* pa = a.ptr(y); const uchar* pb = b.ptr(y); for (int x = 0; x < a.cols; ++x) s += abs((int)pa[x] - (int)pb[x]); }
    return (double)s;
}
struct Mat::Expr { Mat m; };                                     // always an evaluated CV_32F (or source-typed) temporary
inline Mat::Mat(const Expr& e) : rows(0), cols(0), step(0), data(nullptr), type_(0) { *this = e.m; }
inline Mat& Mat::operator=(const Expr& e) {
    if (data && rows == e.m.rows && cols == e.m.cols && type_ == e.m.type()) e.m.copyTo(*this);      // in place (e.g. M.row(0) = ...)
    else *this = e.m;
    return *this;
}
// arithmetic on CV_32F matrices; every operation is one float operation per element (what cv::arithm does for 32F), the matrix
// product accumulates each dot product in double and rounds once (cv::gemm for CV_32F; pinned against cv2.gemm in the oracle tests)
Mat::Expr operator*(const Mat& a, const Mat& b);
Mat::Expr operator-(double s, const Mat& m);
Mat::Expr operator/(double s, const Mat& m);
Mat::Expr operator/(const Mat& m, double s);
inline Mat::Expr operator/(double s, const Mat::Expr& e) { return s / e.m; }
std::ostream& operator<<(std::ostream& os, const Mat& m);

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
    _InputArray(const Mat::Expr& e) : m_(&e.m) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    Mat getMat() const { return *m_; }
    void release() const { m_->release(); }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TOZERO_INV = 4 };
enum { DIST_L2 = 2, DIST_MASK_5 = 5 };

// primitives: declared here, defined in oracle/ref_orbextractor_driver.cpp on the oracle's restatements
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, double value = 0);
double threshold(InputArray src, OutputArray dst, double thresh, double maxval, int type);
void dilate(InputArray src, OutputArray dst, InputArray kernel, Point anchor = Point(-1, -1), int iterations = 1);
Mat getStructuringElement(int shape, Size ksize);
void filter2D(InputArray src, OutputArray dst, int ddepth, InputArray kernel, Point anchor = Point(-1, -1), double delta = 0, int borderType = BORDER_DEFAULT);
void distanceTransform(InputArray src, OutputArray dst, OutputArray labels, int distanceType, int maskSize);
void minMaxLoc(InputArray src, double* minVal, double* maxVal);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };     // only in the unused ComputeKeyPointsOld

// cv::FileStorage stand-in: reads a plain text file of `key value` lines (what the tests write instead of a YAML settings file),
// so that the reference's own parameter parsing runs (DepthModule::ParseRGBLParameters / ParseUpsamplingParameters).
class FileNode {
public:
    enum { NONE = 0, SEQ = 4, MAP = 5 };
    FileNode() : has_(false) {}
    explicit FileNode(const std::string& v) : has_(true), v_(v) {}
    bool empty() const { return !has_; }
    bool isReal() const { if (!has_) return false; char* e = nullptr; strtod(v_.c_str(), &e); return e && *e == 0 && !v_.empty(); }
    double real() const { return has_ ? strtod(v_.c_str(), nullptr) : 0.0; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    int type() const { return NONE; }
    size_t size() const { return 0; }
    operator int() const { return (int)real(); }
    operator float() const { return (float)real(); }
    operator double() const { return real(); }
    operator std::string() const { return v_; }
private:
    bool has_; std::string v_;
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string& path, int mode) {
        if (mode != READ) return;
        std::ifstream f(path.c_str());
        std::string k, v;
        while (f >> k >> v) { keys_.push_back(k); vals_.push_back(v); }
    }
    bool isOpened() const { return !keys_.empty(); }
    void release() {}
    FileNode operator[](const std::string& k) const { for (size_t i = 0; i < keys_.size(); ++i) if (keys_[i] == k) return FileNode(vals_[i]); return FileNode(); }
    FileNode operator[](const char* k) const { return (*this)[std::string(k)]; }
private:
    std::vector<std::string> keys_, vals_;
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv

using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;
