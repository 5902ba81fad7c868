// Replacement of ONE function of the reference's src/DepthModule.cc: DepthModule::CalculateDepthFromPcd (:50-79), same signature, over
// the reference's own include/DepthModule.h.  The constructor and the two settings parsers of src/DepthModule.cc (:30-46, :281-601)
// stay the reference's code (delete lines 50-79 from src/DepthModule.cc and add this file to the library sources; the private helpers
// ProjectPointcloudToImage / Upsample_* / GetFeatureDepthFromDepthMap become dead code).  Everything the function leaves behind in
// the object - mvDepth, mvuRight, RawDepthMap, ProcessedDepthMap - is filled as before (src/Frame.cc:331-333, src/Tracking.cc:1584
// read them).
#include "DepthModule.h"

#include "rgbl_shim_common.h"

namespace ORB_SLAM3 {

void DepthModule::CalculateDepthFromPcd(std::vector<cv::KeyPoint> mvKeys, std::vector<cv::KeyPoint> mvKeysUn, const cv::Mat& PointCloud,
                                        const int imwidth, const int imheight) {
    // Check if all required Parameters are available (:52-55: message + early return leaving stale outputs)
    if (!b_parse_LiDARUpsampling || !b_parse_LiDAR) {
        std::cout << "*Cannot perform LiDAR Upsampling since parameters were missing in the config file.*" << std::endl;
        return;
    }
    rgbl_depth_params prm{};
    prm.min_dist = opt_min_dist; prm.max_dist = opt_max_dist; prm.bf = mbf;
    prm.inv_dilation_scale = ParamUpsampling_InverseDilation_ScaleFactor;
    switch (SelectedUpsamlingMethod) {
        case DepthModule::None: prm.method = RGBL_DEPTH_NONE; break;
        case DepthModule::NearestNeighborPixel:
            prm.method = RGBL_DEPTH_NEAREST_NEIGHBOR_PIXEL; prm.nn_search_radius = ParamUpsampling_NearestNeighborPixel_SearchRadius; break;
        case DepthModule::AverageFiltering:
            prm.method = RGBL_DEPTH_AVERAGE_FILTERING; prm.avg_kernel = ParamUpsampling_AverageFilter_KernelSize; break;
        case DepthModule::InverseDilation: {
            prm.method = RGBL_DEPTH_INVERSE_DILATION;
            prm.ku = ParamUpsampling_InverseDilation_KernelSize_u; prm.kv = ParamUpsampling_InverseDilation_KernelSize_v;
            if (rgbl_depth_structuring_element(ParamUpsampling_InverseDilation_KernelType.c_str(), prm.ku, prm.kv, prm.mask) != RGBL_OK) {
                std::cout << "*Invalid Kernel Type.*" << std::endl;            // src/DepthModule.cc:254-258
                return;
            }
            break;
        }
        default:
            // the reference projects the cloud and then prints this (:73-76); the maps stay as the projection left them
            std::cout << "*Desired Upsampling Method was not yet implemented.*";
            prm.method = RGBL_DEPTH_NONE;
            break;
    }
    const int N = (int)mvKeys.size();
    if (SelectedUpsamlingMethod != DepthModule::None) { mvuRight.assign(N, -1.f); mvDepth.assign(N, -1.f); }     // :84-85, :147-148
    RawDepthMap.create(imheight, imwidth, CV_32F);
    ProcessedDepthMap.create(imheight, imwidth, CV_32F);
    // PointCloud: 4 x N CV_32F, rows x, y, z, 1 (Examples/RGB-L/rgbl_kitti.cc:168-177); LidarProjectionMatrix: the reference's own
    // K [R|t] product (:434), its first three rows go across the ABI unchanged
    cv::Mat pts = (PointCloud.step == (size_t)PointCloud.cols * sizeof(float)) ? PointCloud : PointCloud.clone();
    float depth_dummy = 0, ur_dummy = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_depth_from_pcd(ctx, pts.ptr<float>(), pts.cols, LidarProjectionMatrix.ptr<float>(), imwidth, imheight, &prm,
                                              reinterpret_cast<const rgbl_keypoint*>(mvKeys.data()), reinterpret_cast<const rgbl_keypoint*>(mvKeysUn.data()), N,
                                              N ? mvDepth.data() : &depth_dummy, N ? mvuRight.data() : &ur_dummy, RawDepthMap.ptr<float>(), ProcessedDepthMap.ptr<float>()));
}

}  // namespace ORB_SLAM3
