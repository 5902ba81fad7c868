// ORACLE SUPPORT (test infrastructure, NOT product code): the REFERENCE's own ORBextractor, compiled unmodified from
// /root/reference/src/ORBextractor.cc (oracle/Makefile target `ref` -> oracle/_ref/libref_orbextractor.so), behind a C entry.
// Everything that runs inside ORBextractor::operator() is the reference's code: the scale / feature-quota / umax tables, the
// bordered pyramid and its views, the 35 px cell grid with the iniTh -> minTh fallback, DistributeOctTree with its std::list
// and std::sort, IC_Angle, the steered BRIEF sampling, the mono/stereo output order.  The OpenCV primitives it calls are the
// stand-ins below, built on the oracle's restatements (orb_extractor_oracle.cpp, linked into this library), which are pinned
// against the real OpenCV separately (tests/test_oracle_golden.py).  Used to pin oracle.Extractor (rows a1-a9 of DESIGN.md).
#include <cstdint>
#include <cstring>
#include <vector>

#include <opencv2/core/core.hpp>

extern "C" {   // oracle primitives (orb_extractor_oracle.cpp)
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int orc_fast_window(const uint8_t* win, int w, int h, int stride, int th, int* out, int cap);
float orc_fast_atan2(float y, float x);
}

namespace cv {

void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int) {
    Mat src = src_.getMat();
    dst_.create(dsize.height, dsize.width, src.type());          // a view of the right size is written in place
    Mat dst = dst_.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static int reflect101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * (n - 1) - p; return p; }

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int, double) {
    Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = dst_.getMat();
    // interior first (src may already BE the interior of dst: ComputePyramid passes a view of `temp` as the source)
    for (int y = 0; y < src.rows; ++y) memmove(dst.data + (size_t)(y + top) * dst.step + left, src.data + (size_t)y * src.step, src.cols);
    for (int y = 0; y < dst.rows; ++y) {
        const int sy = reflect101(y - top, src.rows);
        uchar* d = dst.data + (size_t)y * dst.step;
        const uchar* s = dst.data + (size_t)(sy + top) * dst.step + left;      // interior row of dst
        for (int x = 0; x < dst.cols; ++x) {
            if (y >= top && y < top + src.rows && x >= left && x < left + src.cols) continue;
            d[x] = s[reflect101(x - left, src.cols)];
        }
    }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int borderType) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat src = src_.getMat();
    Mat tmp(src.rows, src.cols, src.type());
    orc_gaussian_blur7_u8(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
    dst_.create(src.rows, src.cols, src.type());
    tmp.copyTo(dst_.getMat());
}

void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    assert(nonmaxSuppression);
    Mat img = image_.getMat();
    std::vector<int> out((size_t)img.rows * img.cols * 3 + 3);
    const int n = orc_fast_window(img.data, img.cols, img.rows, (int)img.step, threshold, out.data(), img.rows * img.cols);
    keypoints.clear();
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint(Point2f((float)out[3 * i], (float)out[3 * i + 1]), 7.f, -1.f, (float)out[3 * i + 2]));
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }
void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) { abort(); }

}  // namespace cv

#include "ORBextractor.cc"          // found through -I/root/reference/src: compiled from where it lies

struct RefKp { float x, y, size, angle, response; int32_t octave, class_id; };

extern "C" {

// -> monoIndex (the return value of ORBextractor::operator()) or -2 when cap is too small; n_out = number of keypoints
int ref_orb_extract(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, const uint8_t* img, int w, int h, int stride,
                    int lap0, int lap1, RefKp* kps_out, uint8_t* desc_out, int cap, int* n_out) {
    ORB_SLAM3::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
    cv::Mat image(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) memcpy(image.data + (size_t)y * image.step, img + (size_t)y * stride, w);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc, mask;
    std::vector<int> lapping = {lap0, lap1};
    const int mono = ex(image, mask, kps, desc, lapping);
    *n_out = (int)kps.size();
    if ((int)kps.size() > cap) return -2;
    for (size_t i = 0; i < kps.size(); ++i) {
        const cv::KeyPoint& k = kps[i];
        kps_out[i] = RefKp{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
        memcpy(desc_out + 32 * i, desc.ptr((int)i), 32);
    }
    return mono;
}

// ORBextractor's tables as the reference computes them (include/ORBextractor.h:61-81 getters; umax / quotas are protected)
int ref_orb_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
    ORB_SLAM3::ORBextractor ex(nfeatures, scale_factor, nlevels, 20, 7);
    std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(), d = ex.GetInverseScaleSigmaSquares();
    for (int i = 0; i < nlevels; ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
    return ex.GetLevels();
}

}  // extern "C"
