// Drop-in replacement of include/ORBextractor.h (+ src/ORBextractor.cc): same class name, constructor, operator(), getters and the
// public mvImagePyramid (include/ORBextractor.h:45-108), forwarding to the C ABI.  Put this directory before the reference's include/
// on the include path and drop src/ORBextractor.cc from the library sources.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv2/core/core.hpp>

#include "rgbl_shim_common.h"

namespace ORB_SLAM3 {

static_assert(sizeof(cv::KeyPoint) == sizeof(rgbl_keypoint), "cv::KeyPoint layout (pt.x, pt.y, size, angle, response, octave, class_id)");

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
        : prm_{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST} {
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        std::vector<int32_t> quota(nlevels); int32_t umax[16];
        if (rgbl_orb_tables(&prm_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), quota.data(), umax) != RGBL_OK)
            throw std::runtime_error("librgbl_b200: invalid ORB parameters");
        mvImagePyramid.resize(nlevels);
    }
    ~ORBextractor() {
        if (ctx_) { if (rgbl_shim::shared_context() == ctx_) rgbl_shim::shared_context() = nullptr; rgbl_destroy(ctx_); }
    }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (src/ORBextractor.cc:1086-1168).  Mask is ignored, like in the reference.
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
                   std::vector<int>& vLappingArea) {
        if (_image.empty()) return -1;                                       // :1090-1091
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);                                     // :1094
        ensure(image.cols, image.rows);
        const int cap = rgbl_keypoint_capacity(ctx_);
        _keypoints.resize(cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0, mono = 0;
        rgbl_shim::check(ctx_, rgbl_orb_extract(ctx_, image.data, image.cols, image.rows, (int)image.step, vLappingArea[0], vLappingArea[1],
                                                reinterpret_cast<rgbl_keypoint*>(_keypoints.data()), desc.data, cap, &n, &mono));
        _keypoints.resize(n);
        if (n == 0) _descriptors.release();                                  // :1108-1109
        else { _descriptors.create(n, 32, CV_8U); desc.rowRange(0, n).copyTo(_descriptors.getMat()); }
        if (mbExportPyramid) fetch_pyramid();
        rgbl_shim::shared_context() = ctx_;
        return mono;
    }

    int inline GetLevels() { return prm_.nlevels; }
    float inline GetScaleFactor() { return prm_.scale_factor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Frame::ComputeStereoMatches reads these directly (src/Frame.cc:908,998-1013): views into 19-px bordered planes like
    // src/ORBextractor.cc:1178.  Filled after every operator() while mbExportPyramid is set (a 1.7 MB device-to-host copy per frame);
    // RGB-L / RGB-D / monocular tracking never reads them and may switch the export off.
    std::vector<cv::Mat> mvImagePyramid;
    bool mbExportPyramid = true;

    rgbl_ctx* context() { return ctx_; }

protected:
    void ensure(int w, int h) {
        if (ctx_ && w == w_ && h == h_) return;
        if (ctx_) { if (rgbl_shim::shared_context() == ctx_) rgbl_shim::shared_context() = nullptr; rgbl_destroy(ctx_); ctx_ = nullptr; }
        rgbl_config cfg{};
        cfg.device = 0; cfg.width = w; cfg.height = h; cfg.max_batch = 1; cfg.max_points = 300000; cfg.orb = prm_;
        rgbl_ctx* c = nullptr;
        const int rc = rgbl_create(&cfg, &c);
        if (rc != RGBL_OK) throw std::runtime_error(std::string("librgbl_b200: ") + rgbl_last_error(nullptr));
        ctx_ = c; w_ = w; h_ = h;
        padded_.assign(prm_.nlevels, cv::Mat());
    }
    void fetch_pyramid() {
        for (int l = 0; l < prm_.nlevels; ++l) {
            int lw = 0, lh = 0;
            if (padded_[l].empty()) padded_[l].create(h_ + 38, w_ + 38, CV_8U);          // large enough for every level
            rgbl_shim::check(ctx_, rgbl_orb_get_pyramid(ctx_, 0, l, padded_[l].data, (int)padded_[l].step, &lw, &lh));
            mvImagePyramid[l] = padded_[l](cv::Rect(19, 19, lw, lh));
        }
    }
    rgbl_orb_params prm_;
    rgbl_ctx* ctx_ = nullptr;
    int w_ = 0, h_ = 0;
    std::vector<cv::Mat> padded_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3
#endif
