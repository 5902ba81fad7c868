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
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    size_t step1() const { return step / elemSize(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf_.reset(); }
    void create(int r, int c, int type) {                       // keeps a matching allocation (OpenCV semantics)
        if (data && rows == r && cols == c && type_ == type) return;
        type_ = type; rows = r; cols = c; step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar> >((size_t)r * step + 64, (uchar)0);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); return m; }
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

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(&m) {}
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

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

// primitives: declared here, defined in oracle/ref_orbextractor_driver.cpp on the oracle's restatements
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };     // only in the unused ComputeKeyPointsOld

class FileNode {
public:
    enum { NONE = 0, SEQ = 4, MAP = 5 };
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    int type() const { return NONE; }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator std::string() const { return std::string(); }
};

class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv

using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;
