// Replacement of ONE function of the reference's src/Optimizer.cc: Optimizer::PoseOptimization(Frame*) (:814-1114), same signature.
// The g2o problem of the reference (one SE3 vertex, one unary edge per map point, 4 x 10 Levenberg-Marquardt iterations, Huber kernel,
// outlier classification) runs in librgbl_b200's FP64 kernel; this function only gathers the edges and writes the results back.
// Frames with a second camera (fisheye rigs, pFrame->mpCamera2) are outside the accelerated path and keep the reference's body.
#include "Optimizer.h"

#include <mutex>

#include "rgbl_shim_common.h"

namespace ORB_SLAM3 {

int Optimizer::PoseOptimization(Frame* pFrame) {
    if (pFrame->mpCamera2) throw std::runtime_error("librgbl_b200: PoseOptimization of two-camera rigs is not accelerated; keep the reference's function for them");
    const int N = pFrame->N;
    std::vector<float> xw, obs, inv_sigma2;
    std::vector<uint8_t> stereo;
    std::vector<int> idx;
    xw.reserve(3 * N); obs.reserve(3 * N); inv_sigma2.reserve(N); stereo.reserve(N); idx.reserve(N);
    {
        std::unique_lock<std::mutex> lock(MapPoint::mGlobalMutex);           // :856
        for (int i = 0; i < N; i++) {
            MapPoint* pMP = pFrame->mvpMapPoints[i];
            if (!pMP) continue;
            pFrame->mvbOutlier[i] = false;                                   // :866, :897
            const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
            const Eigen::Vector3f X = pMP->GetWorldPos();
            xw.push_back(X(0)); xw.push_back(X(1)); xw.push_back(X(2));
            const float ur = pFrame->mvuRight[i];
            obs.push_back(kpUn.pt.x); obs.push_back(kpUn.pt.y); obs.push_back(ur);
            inv_sigma2.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
            stereo.push_back(ur < 0 ? 0 : 1);                               // monocular observation iff mvuRight[i] < 0 (:863)
            idx.push_back(i);
        }
    }
    const int nInitialCorrespondences = (int)idx.size();
    if (nInitialCorrespondences < 3) return 0;                               // :995-996
    float Tin[7], Tout[7];
    rgbl_shim::to_pose7(pFrame->GetPose(), Tin);
    std::vector<uint8_t> outlier(nInitialCorrespondences);
    int n_inliers = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_pose_optimize(ctx, Tin, nInitialCorrespondences, xw.data(), obs.data(), inv_sigma2.data(), stereo.data(),
                                             Frame::fx, Frame::fy, Frame::cx, Frame::cy, pFrame->mbf, Tout, outlier.data(), &n_inliers));
    for (int k = 0; k < nInitialCorrespondences; ++k) pFrame->mvbOutlier[idx[k]] = outlier[k] != 0;
    // the kernel returns the float32 pose already normalised like Sophus::SE3f(quaternion, t) does (:1108-1110)
    Sophus::SE3<float> pose(Eigen::Quaternionf(Tout[3], Tout[0], Tout[1], Tout[2]), Eigen::Vector3f(Tout[4], Tout[5], Tout[6]));
    pFrame->SetPose(pose);
    return n_inliers;
}

}  // namespace ORB_SLAM3
