// Reference-side binding of librgbl_b200.so: drop-in replacements for the four hot-path classes of
// ORB_SLAM3 (TUMFTM/ORB_SLAM3_RGBL) with the reference's exact method signatures.  This header needs the
// reference's own dependencies (OpenCV, Eigen, Sophus, the ORB_SLAM3 headers) and is therefore NOT compiled
// in this repository (none of them exist in the build image); see INTEGRATION.md.  It contains glue only:
// marshalling of cv::Mat / std::vector<cv::KeyPoint> / MapPoint* into the flat buffers of include/rgbl_b200.h.
#pragma once
#ifdef RGBL_B200_WITH_ORBSLAM3

#include <opencv2/core.hpp>
#include <stdexcept>
#include <vector>

#include "rgbl_b200.h"

namespace ORB_SLAM3 {

static_assert(sizeof(cv::KeyPoint) == sizeof(rgbl_keypoint), "cv::KeyPoint layout");

// include/ORBextractor.h:45-108 -----------------------------------------------------------------------
class ORBextractor {
public:
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : prm_{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST} {
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        std::vector<int32_t> q(nlevels); int32_t um[16];
        rgbl_orb_tables(&prm_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), q.data(), um);
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() { if (ctx_) rgbl_destroy(ctx_); }

    int operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
        if (_image.empty()) return -1;                                       // src/ORBextractor.cc:1090-1091
        cv::Mat image = _image.getMat();
        CV_Assert(image.type() == CV_8UC1);
        ensure(image.cols, image.rows);
        const int cap = rgbl_keypoint_capacity(ctx_);
        _keypoints.resize(cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0, mono = 0;
        check(rgbl_orb_extract(ctx_, image.data, image.cols, image.rows, (int)image.step, vLappingArea[0], vLappingArea[1],
                               reinterpret_cast<rgbl_keypoint*>(_keypoints.data()), desc.data, cap, &n, &mono));
        _keypoints.resize(n);
        if (n == 0) _descriptors.release(); else desc.rowRange(0, n).copyTo(_descriptors);
        pyramid_valid_ = false;                                             // mvImagePyramid is fetched lazily
        return mono;
    }
    // Frame::ComputeStereoMatches reads mvImagePyramid (src/Frame.cc:908,998-1013): fetch on first use
    const std::vector<cv::Mat>& ImagePyramid() {
        if (!pyramid_valid_) {
            for (int l = 0; l < prm_.nlevels; ++l) {
                cv::Mat padded(h_ + 38, w_ + 38, CV_8U);
                int lw = 0, lh = 0;
                check(rgbl_orb_get_pyramid(ctx_, 0, l, padded.data, (int)padded.step, &lw, &lh));
                mvImagePyramid[l] = padded(cv::Rect(19, 19, lw, lh));      // ROI view like src/ORBextractor.cc:1178
            }
            pyramid_valid_ = true;
        }
        return mvImagePyramid;
    }
    int inline GetLevels() { return prm_.nlevels; }
    float inline GetScaleFactor() { return prm_.scale_factor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    std::vector<cv::Mat> mvImagePyramid;
    rgbl_ctx* context() { return ctx_; }

private:
    void ensure(int w, int h) {
        if (ctx_ && w == w_ && h == h_) return;
        if (ctx_) rgbl_destroy(ctx_);
        rgbl_config cfg{}; cfg.device = 0; cfg.width = w; cfg.height = h; cfg.max_batch = 1; cfg.max_points = 300000; cfg.orb = prm_;
        check(rgbl_create(&cfg, &ctx_)); w_ = w; h_ = h;
    }
    void check(int rc) { if (rc != RGBL_OK) throw std::runtime_error(rgbl_last_error(ctx_)); }
    rgbl_orb_params prm_; rgbl_ctx* ctx_ = nullptr; int w_ = 0, h_ = 0; bool pyramid_valid_ = false;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

// include/DepthModule.h:30-164 (hot path only; the YAML parsing of src/DepthModule.cc:281-601 stays as is) ----------
struct DepthModuleB200 {
    rgbl_ctx* ctx; rgbl_depth_params prm; cv::Mat LidarProjectionMatrix, RawDepthMap, ProcessedDepthMap;
    std::vector<float> mvuRight, mvDepth;
    void CalculateDepthFromPcd(const std::vector<cv::KeyPoint>& mvKeys, const std::vector<cv::KeyPoint>& mvKeysUn,
                               const cv::Mat& PointCloud /*4xN CV_32F*/, int imwidth, int imheight) {
        const int N = (int)mvKeys.size();
        mvuRight.assign(N, -1.f); mvDepth.assign(N, -1.f);
        RawDepthMap.create(imheight, imwidth, CV_32F); ProcessedDepthMap.create(imheight, imwidth, CV_32F);
        cv::Mat pts = PointCloud.isContinuous() ? PointCloud : PointCloud.clone();
        int rc = rgbl_depth_from_pcd(ctx, pts.ptr<float>(), pts.cols, LidarProjectionMatrix.ptr<float>(), imwidth, imheight, &prm,
                                     reinterpret_cast<const rgbl_keypoint*>(mvKeys.data()), reinterpret_cast<const rgbl_keypoint*>(mvKeysUn.data()), N,
                                     mvDepth.data(), mvuRight.data(), RawDepthMap.ptr<float>(), ProcessedDepthMap.ptr<float>());
        if (rc != RGBL_OK) throw std::runtime_error(rgbl_last_error(ctx));
    }
};

}  // namespace ORB_SLAM3
#endif  // RGBL_B200_WITH_ORBSLAM3
