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

#include <opencv2/depth_primitives_impl.hpp>

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
