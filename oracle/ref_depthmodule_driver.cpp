// ORACLE SUPPORT (test infrastructure, NOT product code): the REFERENCE's own DepthModule, compiled unmodified from
// /root/reference/src/DepthModule.cc (oracle/Makefile target `ref` -> oracle/_ref/libref_depthmodule.so), behind a C entry.
// What runs is the reference's code: the parameter parsing (projection matrix = K [R|t] from the settings, distance limits,
// method and structuring element selection), ProjectPointcloudToImage (matrix product, normalisation, the sequential
// last-writer-wins scatter), Upsample_InverseDilation (invert, threshold, dilate, invert, threshold) and
// GetFeatureDepthFromDepthMap.  The OpenCV operations under it are the stand-ins defined below, each written to the arithmetic
// that the oracle's restatements were pinned to against python-cv2 (float matrix product with double accumulation, one float
// operation per element for s - M and the row scaling, THRESH_TOZERO_INV, cv::dilate ignoring out-of-image taps,
// getStructuringElement).  Only the InverseDilation method (the default of Examples/RGB-L/*.yaml) is supported here; the
// primitives of the two other methods abort.  Used to pin oracle.depth_from_pcd (rows a10-a12, a14 of DESIGN.md).
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

#include "DepthModule.cc"           // found through -I/root/reference/src: compiled from where it lies

extern "C" {

// settings: path of a `key value` text file with the keys the reference parses (Camera.fx/fy/cx/cy/bf, LiDAR.Tr11..Tr34,
// LiDAR.min_dist/max_dist, LiDAR.Method, LiDAR.MethodInverseDilation.*).  pts: 4 x n planar float rows (x | y | z | 1) like
// the cv::Mat the RGB-L example builds; kps / kps_un: n_kp x 2 (x, y).  Outputs: the reference's LidarProjectionMatrix (first 12
// of its 16 floats = the 3 x 4 part), RawDepthMap, ProcessedDepthMap (W x H floats), mvDepth, mvuRight.
int ref_depth_from_pcd(const char* settings, const float* pts, int n, int W, int H, const float* kps, const float* kps_un, int n_kp,
                       float* P_out, float* raw_out, float* processed_out, float* depth_out, float* uright_out) {
    ORB_SLAM3::DepthModule dm(settings, 0);
    cv::Mat cloud(4, n, CV_32F);
    for (int r = 0; r < 4; ++r) memcpy(cloud.ptr<float>(r), pts + (size_t)r * n, (size_t)n * sizeof(float));
    std::vector<cv::KeyPoint> k(n_kp), ku(n_kp);
    for (int i = 0; i < n_kp; ++i) { k[i].pt = cv::Point2f(kps[2 * i], kps[2 * i + 1]); ku[i].pt = cv::Point2f(kps_un[2 * i], kps_un[2 * i + 1]); }
    dm.CalculateDepthFromPcd(k, ku, cloud, W, H);
    if (dm.RawDepthMap.empty() || dm.ProcessedDepthMap.empty() || (int)dm.mvDepth.size() != n_kp) return -1;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P_out[4 * r + c] = dm.LidarProjectionMatrix.at<float>(r, c);
    for (int y = 0; y < H; ++y) {
        memcpy(raw_out + (size_t)y * W, dm.RawDepthMap.ptr<float>(y), (size_t)W * sizeof(float));
        memcpy(processed_out + (size_t)y * W, dm.ProcessedDepthMap.ptr<float>(y), (size_t)W * sizeof(float));
    }
    for (int i = 0; i < n_kp; ++i) { depth_out[i] = dm.mvDepth[i]; uright_out[i] = dm.mvuRight[i]; }
    return 0;
}

}  // extern "C"
