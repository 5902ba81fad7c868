// TEST INFRASTRUCTURE.  Definitions of the stand-in OpenCV operations src/DepthModule.cc calls (declared in core/core.hpp next to this
// file), each written to the arithmetic the oracle's restatements were pinned to against python-cv2: float matrix product with double
// accumulation, one float operation per element for s - M and the row scaling, THRESH_TOZERO_INV, cv::dilate ignoring out-of-image
// taps, getStructuringElement.  Included by oracle/ref_depthmodule_driver.cpp (the reference's DepthModule as a whole) and by the
// binding test (tests/shim/shim_driver.cpp: the reference's DepthModule constructor + parsers around the replaced hot function).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

namespace cv {

Mat Mat::ones(int r, int c, int type) { Mat m(r, c, type); if (type == CV_32F) for (int y = 0; y < r; ++y) for (int x = 0; x < c; ++x) m.at<float>(y, x) = 1.f; else abort(); return m; }
Mat::Expr Mat::mul(const Mat& b) const {
    Mat o(rows, cols, CV_32F);
    for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) o.at<float>(y, x) = at<float>(y, x) * b.at<float>(y, x);
    return Expr{o};
}
void Mat::convertTo(Mat&, int) const { abort(); }
Mat::Expr operator*(const Mat& a, const Mat& b) {               // gemm, CV_32F: double accumulation, one rounding
    Mat o(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            double acc = 0;
            for (int k = 0; k < a.cols; ++k) acc += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
            o.at<float>(i, j) = (float)acc;
        }
    return Mat::Expr{o};
}
Mat::Expr operator-(double s, const Mat& m) {
    Mat o(m.rows, m.cols, CV_32F);
    const float fs = (float)s;
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) o.at<float>(y, x) = fs - m.at<float>(y, x);
    return Mat::Expr{o};
}
Mat::Expr operator/(double s, const Mat& m) {
    Mat o(m.rows, m.cols, CV_32F);
    const float fs = (float)s;
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) o.at<float>(y, x) = fs / m.at<float>(y, x);
    return Mat::Expr{o};
}
Mat::Expr operator/(const Mat&, double) { abort(); }
std::ostream& operator<<(std::ostream& os, const Mat&) { return os; }

double threshold(InputArray src_, OutputArray dst_, double thresh, double, int type) {
    if (type != THRESH_TOZERO_INV) abort();
    Mat src = src_.getMat();
    dst_.create(src.rows, src.cols, CV_32F);
    Mat dst = dst_.getMat();
    const float t = (float)thresh;
    for (int y = 0; y < src.rows; ++y) for (int x = 0; x < src.cols; ++x) { const float v = src.at<float>(y, x); dst.at<float>(y, x) = (v > t) ? 0.f : v; }
    return thresh;
}

void dilate(InputArray src_, OutputArray dst_, InputArray kernel_, Point anchor, int iterations) {
    if (iterations != 1 || anchor.x != -1 || anchor.y != -1) abort();
    Mat src = src_.getMat().clone(), k = kernel_.getMat();       // clone: the reference dilates in place
    dst_.create(src.rows, src.cols, CV_32F);
    Mat dst = dst_.getMat();
    const int ax = k.cols / 2, ay = k.rows / 2;
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            float best = -FLT_MAX;                                // BORDER_CONSTANT with morphologyDefaultBorderValue(): outside never wins
            for (int j = 0; j < k.rows; ++j)
                for (int i = 0; i < k.cols; ++i) {
                    if (!k.at<uchar>(j, i)) continue;
                    const int yy = y + j - ay, xx = x + i - ax;
                    if (yy < 0 || yy >= src.rows || xx < 0 || xx >= src.cols) continue;
                    const float v = src.at<float>(yy, xx);
                    if (v > best) best = v;
                }
            dst.at<float>(y, x) = best;
        }
}

Mat getStructuringElement(int shape, Size ksize) {             // OpenCV imgproc/src/morph.dispatch.cpp, anchor at the centre
    Mat e(ksize.height, ksize.width, CV_8U);
    const int r = ksize.height / 2, c = ksize.width / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < ksize.height; ++i) {
        int j1 = 0, j2 = 0;
        if (shape == MORPH_RECT || (shape == MORPH_CROSS && i == r)) j2 = ksize.width;
        else if (shape == MORPH_CROSS) { j1 = c; j2 = c + 1; }
        else {
            const int dy = i - r;
            if (std::abs(dy) <= r) { const int dx = cvRound(c * std::sqrt((r * r - dy * dy) * inv_r2)); j1 = std::max(c - dx, 0); j2 = std::min(c + dx + 1, ksize.width); }
        }
        for (int j = j1; j < j2; ++j) e.at<uchar>(i, j) = 1;
    }
    return e;
}

void copyMakeBorder(InputArray, OutputArray, int, int, int, int, int, double) { abort(); }       // NearestNeighborPixel only
void filter2D(InputArray, OutputArray, int, InputArray, Point, double, int) { abort(); }         // AverageFiltering only
void distanceTransform(InputArray, OutputArray, OutputArray, int, int) { abort(); }
void minMaxLoc(InputArray, double*, double*) { abort(); }

}  // namespace cv
