// TEST INFRASTRUCTURE.  Stand-in for <opencv2/core/core.hpp>, just large enough for the reference's vendored DBoW2
// (Thirdparty/DBoW2) to compile UNMODIFIED from /root/reference into oracle/_ref/libref_dbow2.so (oracle/Makefile, target
// `ref`): a row-major byte matrix with value semantics and inert cv::FileStorage / cv::FileNode types (the YAML save/load
// members of TemplatedVocabulary are virtual, so they must compile, but nothing calls them: ORB-SLAM3 loads ORBvoc.txt with
// loadFromTextFile).  DBoW2 only reads/writes descriptors through ptr<T>(), so deep copies are equivalent to cv::Mat's
// reference counting here.
#pragma once
// the real header pulls these in, and DBoW2 relies on it
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <cstddef>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) { rows = r; cols = c; elem_ = (type == CV_32F) ? 4 : 1; data_.assign((size_t)r * c * elem_, 0); }
    void release() { rows = cols = 0; data_.clear(); }
    bool empty() const { return data_.empty(); }
    Mat clone() const { return *this; }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    template <class T> T* ptr(int row = 0) { return reinterpret_cast<T*>(data_.data() + (size_t)row * cols * elem_); }
    template <class T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(data_.data() + (size_t)row * cols * elem_); }
private:
    int elem_ = 1;
    std::vector<uint8_t> data_;
};

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
