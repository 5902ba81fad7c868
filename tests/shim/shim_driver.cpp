// TEST INFRASTRUCTURE: the reference-side binding (shim/) compiled and driven the way the reference drives the four classes.
// What is compiled here:
//   * shim/ORBextractor.h, shim/DepthModule_b200.cc, shim/ORBmatcher_b200.cc, shim/Optimizer_b200.cc - the binding (product glue);
//   * the reference's own code AROUND the replaced functions, from /root/reference: src/DepthModule.cc and src/ORBmatcher.cc with
//     exactly the replaced definitions cut out (oracle/extract_ref.py --except), include/DepthModule.h, include/ORBmatcher.h,
//     CameraModels/Pinhole.*, Frame::{SetPose, UpdatePoseMatrices, AssignFeaturesToGrid, isInFrustum, UnprojectStereo ...} and
//     MapPoint::PredictScale verbatim - so the binding is checked against the declarations it has to match;
//   * stand-ins for OpenCV / Eigen / Sophus and the class shells (oracle/ref_shim).
// main() is caller code shaped like the reference's: the RGB-L Frame constructor (src/Frame.cc:289-377), the map points of
// Tracking::StereoInitialization / UpdateLastFrame (src/Tracking.cc:2384-2431, 2826-2886), TrackWithMotionModel (:2888-2981) and
// TrackLocalMap / SearchLocalPoints (:2983-3050, 3377-3460).  It links against librgbl_b200.so; without a CUDA device rgbl_create
// fails and the program reports that and exits with status 3 (the binding has no CPU fallback).
//   shim_driver SETTINGS INPUT OUTPUT
#define RGBL_B200_SHIM
#include "orbslam_standins.hpp"
#include <opencv2/depth_primitives_impl.hpp>

#include "ORBextractor.h"              // shim/ORBextractor.h (the include path lists shim/ first)

#include "gen/DepthModule_rest.cc"     // the reference's DepthModule constructor + settings parsers (+ now-dead helpers)
#include "gen/ORBmatcher_rest.cc"      // the reference's ORBmatcher.cc without the five tracking-thread matchers
#include "CameraModels/Pinhole.cpp"

namespace ORB_SLAM3 {
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
#define mbIsSet mbHasPose
#include "gen/frame_functions.inc"
#undef mbIsSet
#include "gen/keyframe_functions.inc"
#include "gen/mappoint_functions.inc"
}  // namespace ORB_SLAM3

#include "DepthModule_b200.cc"
#include "ORBmatcher_b200.cc"
#include "Optimizer_b200.cc"

#include <cstdio>

using namespace ORB_SLAM3;

namespace {
template <class T> void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short input\n"); exit(2); } }
template <class T> void wr(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

// the part of Frame::Frame (RGB-L, src/Frame.cc:289-377) that runs around the replaced classes
void construct_frame(Frame& F, const cv::Mat& imGray, const cv::Mat& PointCloud, ORBextractor* extractor, DepthModule* pDepthHandler,
                     GeometricCamera* pCamera, float bf, const float K[4]) {
    F.mpORBextractorLeft = extractor; F.mpCamera = pCamera; F.mpCamera2 = nullptr; F.mbf = bf;
    F.mnScaleLevels = extractor->GetLevels();                              // :303-309
    F.mfScaleFactor = extractor->GetScaleFactor();
    F.mfLogScaleFactor = log(F.mfScaleFactor);
    F.mvScaleFactors = extractor->GetScaleFactors();
    F.mvInvScaleFactors = extractor->GetInverseScaleFactors();
    F.mvLevelSigma2 = extractor->GetScaleSigmaSquares();
    F.mvInvLevelSigma2 = extractor->GetInverseScaleSigmaSquares();
    std::vector<int> vLapping = {0, 0};                                    // Frame::ExtractORB(0, im, 0, 0), :508-515
    (*extractor)(imGray, cv::Mat(), F.mvKeys, F.mDescriptors, vLapping);
    F.N = (int)F.mvKeys.size();
    if (F.mvKeys.empty()) return;
    F.mvKeysUn = F.mvKeys;                                                 // UndistortKeyPoints with k1 == 0 (:837-843)
    pDepthHandler->CalculateDepthFromPcd(F.mvKeys, F.mvKeysUn, PointCloud, imGray.cols, imGray.rows);       // :331-333
    F.mvDepth = pDepthHandler->mvDepth;
    F.mvuRight = pDepthHandler->mvuRight;
    F.mvpMapPoints = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    F.mvbOutlier = std::vector<bool>(F.N, false);
    Frame::mnMinX = 0.0f; Frame::mnMaxX = imGray.cols; Frame::mnMinY = 0.0f; Frame::mnMaxY = imGray.rows;   // ComputeImageBounds, no distortion (:893-897)
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3]; Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
    F.mb = F.mbf / Frame::fx;
    F.Nleft = -1; F.Nright = -1;
    F.AssignFeaturesToGrid();
}
}  // namespace

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: shim_driver SETTINGS INPUT OUTPUT\n"); return 2; }
    FILE* fi = fopen(argv[2], "rb");
    if (!fi) { fprintf(stderr, "cannot open input\n"); return 2; }
    int hdr[4]; float cam[5], pose0[7];
    rd(fi, hdr, 4); rd(fi, cam, 5); rd(fi, pose0, 7);
    const int W = hdr[0], H = hdr[1], n0 = hdr[2], n1 = hdr[3];
    cv::Mat im0(H, W, CV_8U), im1(H, W, CV_8U), pc0(4, n0, CV_32F), pc1(4, n1, CV_32F);
    rd(fi, im0.data, (size_t)W * H); rd(fi, im1.data, (size_t)W * H); rd(fi, pc0.ptr<float>(), (size_t)4 * n0); rd(fi, pc1.ptr<float>(), (size_t)4 * n1);
    fclose(fi);
    try {
        ORBextractor extractor(2000, 1.2f, 8, 12, 7);                      // src/Tracking.cc:595-601 with Examples/RGB-L/KITTI00-02.yaml
        DepthModule depth(argv[1], 0);                                    // src/System.cc:220-221
        Pinhole camera(std::vector<float>{cam[0], cam[1], cam[2], cam[3]});
        Frame last, cur;
        construct_frame(last, im0, pc0, &extractor, &depth, &camera, cam[4], cam);
        // the image pyramid stays readable for Frame::ComputeStereoMatches-style callers (public mvImagePyramid)
        const int pyr0_w = extractor.mvImagePyramid[0].cols, pyr7_h = extractor.mvImagePyramid[7].rows;
        const unsigned char pyr_probe = extractor.mvImagePyramid[1].at<unsigned char>(10, 10);
        last.SetPose(Sophus::SE3f(Eigen::Quaternionf(pose0[3], pose0[0], pose0[1], pose0[2]), Eigen::Vector3f(pose0[4], pose0[5], pose0[6])));
        // map points from the LiDAR depths of the first frame (Tracking::StereoInitialization, src/Tracking.cc:2398-2417)
        std::vector<MapPoint*> all;
        for (int i = 0; i < last.N; i++) {
            Eigen::Vector3f x3D;
            if (!last.UnprojectStereo(i, x3D)) continue;
            MapPoint* p = new MapPoint();
            p->mWorldPos = x3D; p->mDescriptor = last.mDescriptors.row(i).clone(); p->nObs = 1;
            const Eigen::Vector3f PC = x3D - last.mOw;                     // MapPoint::UpdateNormalAndDepth, one observation (src/MapPoint.cc:437-490)
            const float dist = PC.norm();
            p->mNormalVector = PC / dist;
            p->mfMaxDistance = dist * last.mvScaleFactors[last.mvKeysUn[i].octave];
            p->mfMinDistance = p->mfMaxDistance / last.mvScaleFactors[last.mnScaleLevels - 1];
            last.mvpMapPoints[i] = p; all.push_back(p);
        }
        construct_frame(cur, im1, pc1, &extractor, &depth, &camera, cam[4], cam);
        // ---- Tracking::TrackWithMotionModel (src/Tracking.cc:2888-2981), zero velocity ----
        cur.SetPose(last.GetPose());
        ORBmatcher matcher(0.9, true);
        const int nmatches = matcher.SearchByProjection(cur, last, 15, false);
        std::vector<int> match1(cur.N, -1);
        for (int i = 0; i < cur.N; i++) if (cur.mvpMapPoints[i]) for (int j = 0; j < last.N; j++) if (last.mvpMapPoints[j] == cur.mvpMapPoints[i]) { match1[i] = j; break; }
        const int inliers1 = Optimizer::PoseOptimization(&cur);
        std::vector<unsigned char> outlier1(cur.N, 0);
        float pose1[7]; rgbl_shim::to_pose7(cur.GetPose(), pose1);
        std::set<MapPoint*> matched;
        for (int i = 0; i < cur.N; i++) {                                  // discard outliers (:2944-2966)
            if (!cur.mvpMapPoints[i]) continue;
            if (cur.mvbOutlier[i]) { outlier1[i] = 1; cur.mvpMapPoints[i] = static_cast<MapPoint*>(NULL); cur.mvbOutlier[i] = false; }
            else matched.insert(cur.mvpMapPoints[i]);
        }
        // ---- Tracking::TrackLocalMap / SearchLocalPoints (:2983-3050, 3377-3460): local map = the first frame's points in index order ----
        std::vector<MapPoint*> vpLocal;
        std::vector<int> local_src;
        for (int j = 0; j < last.N; j++) {
            MapPoint* p = last.mvpMapPoints[j];
            if (!p) continue;
            vpLocal.push_back(p); local_src.push_back(j);
            if (matched.count(p)) { p->mbTrackInView = false; continue; }  // "do not search map points already matched" (:3380-3398)
            cur.isInFrustum(p, 0.5);                                       // the reference's own function (stays on the host in this binding)
        }
        ORBmatcher matcher2(0.8);
        const int nlocal = matcher2.SearchByProjection(cur, vpLocal, 3, false, 50.0f);
        std::vector<int> match2(cur.N, -1);
        for (int i = 0; i < cur.N; i++) if (cur.mvpMapPoints[i]) for (int j = 0; j < last.N; j++) if (last.mvpMapPoints[j] == cur.mvpMapPoints[i]) { match2[i] = j; break; }
        const int inliers2 = Optimizer::PoseOptimization(&cur);
        float pose2[7]; rgbl_shim::to_pose7(cur.GetPose(), pose2);
        unsigned char d_ab[32], d_ba[32];
        memcpy(d_ab, last.mDescriptors.ptr<unsigned char>(0), 32); memcpy(d_ba, cur.mDescriptors.ptr<unsigned char>(0), 32);
        const int hd = ORBmatcher::DescriptorDistance(last.mDescriptors.row(0), cur.mDescriptors.row(0));

        FILE* fo = fopen(argv[3], "wb");
        const int out_hdr[10] = {last.N, cur.N, nmatches, inliers1, nlocal, inliers2, hd, pyr0_w, pyr7_h, (int)pyr_probe};
        wr(fo, out_hdr, 10);
        for (Frame* F : {&last, &cur}) {
            wr(fo, reinterpret_cast<const rgbl_keypoint*>(F->mvKeys.data()), (size_t)F->N);
            wr(fo, F->mDescriptors.data, (size_t)F->N * 32);
            wr(fo, F->mvDepth.data(), (size_t)F->N); wr(fo, F->mvuRight.data(), (size_t)F->N);
        }
        wr(fo, match1.data(), (size_t)cur.N); wr(fo, outlier1.data(), (size_t)cur.N); wr(fo, pose1, 7);
        wr(fo, match2.data(), (size_t)cur.N); wr(fo, pose2, 7);
        std::vector<unsigned char> ol2(cur.N); for (int i = 0; i < cur.N; i++) ol2[i] = cur.mvbOutlier[i] ? 1 : 0;
        wr(fo, ol2.data(), (size_t)cur.N);
        wr(fo, depth.ProcessedDepthMap.ptr<float>(), (size_t)W * H);       // src/Tracking.cc:1584 reads it
        fclose(fo);
        for (MapPoint* p : all) delete p;
    } catch (const std::exception& e) {
        fprintf(stderr, "shim_driver: %s\n", e.what());
        return 3;
    }
    return 0;
}
