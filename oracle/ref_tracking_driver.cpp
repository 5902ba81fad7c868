// ORACLE SUPPORT (test infrastructure, NOT product code): the REFERENCE's own tracking-thread code behind C entry points with the
// same argument lists as the oracle restatements (matcher_oracle.cpp, pose_oracle.cpp, stereo_oracle.cpp, mapping_oracle.cpp), so
// that tests/test_oracle_tracking_ref.py can feed both the same arrays and compare (oracle/Makefile target `ref` ->
// oracle/_ref/libref_tracking.so; needs /root/reference, built in this container, travels to the GPU box as a binary).
//
// What is the reference's code, compiled unmodified from where it lies:
//   * src/ORBmatcher.cc — the whole file (#include below): SearchByProjection (last frame, local map, relocalisation),
//     SearchByBoW, SearchForTriangulation, Fuse, ComputeThreeMaxima, DescriptorDistance;
//   * src/OptimizableTypes.cpp — the whole file (EdgeSE3ProjectXYZOnlyPose::linearizeOplus ...), include/OptimizableTypes.h;
//   * src/CameraModels/Pinhole.cpp — the whole file, over the real GeometricCamera.h / Pinhole.h;
//   * Thirdparty/g2o — every source of its library (linked from _ref/obj_g2o/*.o): SparseOptimizer, OptimizableGraph, BlockSolver,
//     LinearSolverDense, OptimizationAlgorithmLevenberg, RobustKernelHuber, SE3Quat, VertexSE3Expmap,
//     EdgeStereoSE3ProjectXYZOnlyPose ...;
//   * Optimizer::PoseOptimization (src/Optimizer.cc:814-1114), Frame::{SetPose, UpdatePoseMatrices, AssignFeaturesToGrid,
//     isInFrustum, GetFeaturesInArea, PosInGrid, ComputeStereoMatches, ComputeStereoFromRGBD, UnprojectStereo} (src/Frame.cc),
//     KeyFrame::GetFeaturesInArea (src/KeyFrame.cc), MapPoint::{PredictScale, ComputeDistinctiveDescriptors} (src/MapPoint.cc):
//     cut verbatim at build time by oracle/extract_ref.py into _ref/gen/*.inc and compiled against the stand-in class declarations of
//     ref_shim/orbslam_standins.hpp.
// What is substituted: Eigen (ref_shim/Eigen/mini_eigen.hpp), Sophus (ref_shim/sophus/se3.hpp), OpenCV's cv::Mat / KeyPoint
// (ref_shim/opencv2), and the class shells around the functions.  Their arithmetic conventions are stated in those headers.
#include "orbslam_standins.hpp"

#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_dense.h"
#include "Thirdparty/g2o/g2o/types/types_sba.h"
// g2o::LinearSolverEigen (solvers/linear_solver_eigen.h) wraps Eigen's sparse SimplicialLDLT with an AMD ordering, which the stand-in
// Eigen does not have: Optimizer::LocalBundleAdjustment gets the same linear system solved by the dense LDLT instead (same solution up
// to rounding; the BA comparison is tolerance-based).
#define G2O_LINEAR_SOLVER_EIGEN_H
namespace g2o { template <typename MatrixType> class LinearSolverEigen : public LinearSolverDense<MatrixType> {}; }
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"

#include "ORBmatcher.cc"                    // found through -I/root/reference/src
#include "OptimizableTypes.cpp"
#include "CameraModels/Pinhole.cpp"

namespace ORB_SLAM3 {
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
// SetPose touches one member the stand-in does not model
#define mbIsSet mbHasPose
#include "gen/frame_functions.inc"
#undef mbIsSet
#include "gen/keyframe_functions.inc"
#include "gen/mappoint_functions.inc"
#include "gen/optimizer_functions.inc"
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {

struct RefKp { float x, y, size, angle, response; int32_t octave, class_id; };

extern "C" struct ref_frame_view {           // = orc_frame_view (matcher_oracle.cpp)
    int n;
    const void* keys_un; const float* uright; const uint8_t* desc;
    float min_x, max_x, min_y, max_y;
    int n_levels; const float* scale_factors;
    float fx, fy, cx, cy, bf;
    float log_scale_factor;
};

cv::Mat desc_mat(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8U);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

Sophus::SE3f to_se3(const float p[7]) {      // (qx, qy, qz, qw, tx, ty, tz); SE3f(quaternion, t) normalises like Frame::SetPose's argument was built
    return Sophus::SE3f(Eigen::Quaternionf(p[3], p[0], p[1], p[2]), Eigen::Vector3f(p[4], p[5], p[6]));
}

// A Frame as the RGB-L / stereo constructor leaves it (src/Frame.cc:289-377): Nleft == -1, pinhole camera, grid assigned.
struct FrameHolder {
    Frame F;
    Pinhole cam;
    std::vector<MapPoint*> owned;
    explicit FrameHolder(const ref_frame_view* v) : cam(std::vector<float>{v->fx, v->fy, v->cx, v->cy}) {
        const RefKp* k = (const RefKp*)v->keys_un;
        F.N = v->n;
        F.mvKeysUn.resize(v->n);
        for (int i = 0; i < v->n; ++i) F.mvKeysUn[i] = cv::KeyPoint(cv::Point2f(k[i].x, k[i].y), k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id);
        F.mvKeys = F.mvKeysUn;
        F.mvuRight.assign(v->uright, v->uright + v->n);
        F.mvDepth.assign(v->n, -1.f);
        F.mDescriptors = desc_mat(v->desc, v->n);
        F.mvpMapPoints.assign(v->n, nullptr);
        F.mvbOutlier.assign(v->n, false);
        F.mpCamera = &cam;
        Frame::fx = v->fx; Frame::fy = v->fy; Frame::cx = v->cx; Frame::cy = v->cy;
        Frame::invfx = 1.0f / v->fx; Frame::invfy = 1.0f / v->fy;                                         // src/Frame.cc:357-358
        Frame::mnMinX = v->min_x; Frame::mnMaxX = v->max_x; Frame::mnMinY = v->min_y; Frame::mnMaxY = v->max_y;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);    // :351-352
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        F.mbf = v->bf; F.mb = v->bf / v->fx;                                                             // :360
        F.mnScaleLevels = v->n_levels;
        F.mfLogScaleFactor = v->log_scale_factor;
        F.mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->n_levels);
        F.mvInvScaleFactors.resize(v->n_levels); F.mvLevelSigma2.resize(v->n_levels); F.mvInvLevelSigma2.resize(v->n_levels);
        for (int l = 0; l < v->n_levels; ++l) {                                                          // src/ORBextractor.cc:416-429
            F.mvLevelSigma2[l] = F.mvScaleFactors[l] * F.mvScaleFactors[l];
            F.mvInvScaleFactors[l] = 1.0f / F.mvScaleFactors[l];
            F.mvInvLevelSigma2[l] = 1.0f / F.mvLevelSigma2[l];
        }
        F.AssignFeaturesToGrid();
    }
    MapPoint* add_point() { owned.push_back(new MapPoint()); return owned.back(); }
    ~FrameHolder() { for (MapPoint* p : owned) delete p; }
};

}  // namespace

extern "C" {

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(desc_mat(a, 1), desc_mat(b, 1));
}

int ref_features_in_area(const ref_frame_view* fv, float x, float y, float r, int min_level, int max_level, int* out, int cap) {
    FrameHolder h(fv);
    const std::vector<size_t> v = h.F.GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono): arguments as orc_search_by_projection_last.
// match[i2]: index i of the last-frame point now held by slot i2, else -1 (slot holds what it held on entry), -2 (slot was occupied
// on entry and is NULL now = cleared by the rotation check).
int ref_search_by_projection_last(const ref_frame_view* cur, const float cur_pose[7], const float last_pose[7],
                                  int n_last, const uint8_t* valid, const float* xw, const uint8_t* mp_desc,
                                  const int* last_octave, const float* last_angle, const uint8_t* obs_pos,
                                  float th, int mono, int check_orientation, const uint8_t* cur_state, int* match) {
    FrameHolder C(cur);
    C.F.SetPose(to_se3(cur_pose));
    for (int i = 0; i < cur->n; ++i)
        if (cur_state[i]) { MapPoint* d = C.add_point(); d->nObs = (cur_state[i] == 1) ? 1 : 0; C.F.mvpMapPoints[i] = d; }
    // the last frame: only the members the function reads
    Frame L;
    L.N = n_last; L.Nleft = -1;
    L.mvKeysUn.resize(n_last); L.mvpMapPoints.assign(n_last, nullptr); L.mvbOutlier.assign(n_last, false);
    std::map<MapPoint*, int> index_of;
    for (int i = 0; i < n_last; ++i) {
        L.mvKeysUn[i].octave = last_octave[i]; L.mvKeysUn[i].angle = last_angle[i];
        if (!valid[i]) continue;
        MapPoint* p = C.add_point();
        p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p->mDescriptor = desc_mat(mp_desc + 32 * i, 1);
        p->nObs = obs_pos[i] ? 1 : 0;
        L.mvpMapPoints[i] = p; index_of[p] = i;
    }
    L.mvKeys = L.mvKeysUn;
    L.SetPose(to_se3(last_pose));
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int n = matcher.SearchByProjection(C.F, L, th, mono != 0);
    for (int i = 0; i < cur->n; ++i) {
        MapPoint* p = C.F.mvpMapPoints[i];
        if (p && index_of.count(p)) match[i] = index_of[p];
        else if (!p && cur_state[i]) match[i] = -2;
        else match[i] = -1;
    }
    return n;
}

// Frame::isInFrustum + MapPoint::PredictScale, arguments as orc_is_in_frustum (the frame pose enters as mRcw, mtcw, mOw).
void ref_is_in_frustum(const ref_frame_view* fv, const float* Rcw, const float* tcw, const float* Ow, int n,
                       const float* xw, const float* normal, const float* mf_min_dist, const float* mf_max_dist,
                       float cos_limit, uint8_t* in_view, float* px, float* py, float* pxr, float* depth,
                       int* level, float* view_cos) {
    FrameHolder h(fv);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) h.F.mRcw(r, c) = Rcw[3 * r + c]; h.F.mtcw(r) = tcw[r]; h.F.mOw(r) = Ow[r]; }
    for (int i = 0; i < n; ++i) {
        MapPoint p;
        p.mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p.mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p.mfMinDistance = mf_min_dist[i]; p.mfMaxDistance = mf_max_dist[i];
        p.mTrackProjXR = 0; p.mTrackDepth = 0; p.mnTrackScaleLevel = 0; p.mTrackViewCos = 0;
        const bool ok = h.F.isInFrustum(&p, cos_limit);
        in_view[i] = ok ? 1 : 0; px[i] = p.mTrackProjX; py[i] = p.mTrackProjY; pxr[i] = p.mTrackProjXR; depth[i] = p.mTrackDepth;
        level[i] = p.mnTrackScaleLevel; view_cos[i] = p.mTrackViewCos;
    }
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), arguments as orc_search_by_projection_local.
int ref_search_by_projection_local(const ref_frame_view* fv, int n, const uint8_t* in_view, const float* px,
                                   const float* py, const float* pxr, const float* track_depth, const int* level,
                                   const float* view_cos, const uint8_t* mp_desc, const uint8_t* obs_pos, float th,
                                   float nn_ratio, int far_points, float th_far, const uint8_t* cur_state, int* match) {
    FrameHolder h(fv);
    for (int i = 0; i < fv->n; ++i)
        if (cur_state[i]) { MapPoint* d = h.add_point(); d->nObs = (cur_state[i] == 1) ? 1 : 0; h.F.mvpMapPoints[i] = d; }
    std::vector<MapPoint*> pts(n);
    std::map<MapPoint*, int> index_of;
    for (int i = 0; i < n; ++i) {
        MapPoint* p = h.add_point();
        p->mbTrackInView = in_view[i] != 0; p->mbTrackInViewR = false;
        p->mTrackProjX = px[i]; p->mTrackProjY = py[i]; p->mTrackProjXR = pxr[i]; p->mTrackDepth = track_depth[i];
        p->mnTrackScaleLevel = level[i]; p->mTrackViewCos = view_cos[i];
        p->mDescriptor = desc_mat(mp_desc + 32 * i, 1);
        p->nObs = obs_pos[i] ? 1 : 0;
        pts[i] = p; index_of[p] = i;
    }
    ORBmatcher matcher(nn_ratio, true);
    const int nm = matcher.SearchByProjection(h.F, pts, th, far_points != 0, th_far);
    for (int i = 0; i < fv->n; ++i) {
        MapPoint* p = h.F.mvpMapPoints[i];
        match[i] = (p && index_of.count(p)) ? index_of[p] : -1;
    }
    return nm;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), arguments as orc_search_by_bow.
int ref_search_by_bow(int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                      int n_nodes_kf, const uint32_t* kf_node_ids, const int* kf_node_start, const int* kf_node_feat,
                      int n_f, const uint8_t* f_desc, const float* f_angle,
                      int n_nodes_f, const uint32_t* f_node_ids, const int* f_node_start, const int* f_node_feat,
                      float nn_ratio, int check_orientation, int* match) {
    KeyFrame kf;
    kf.N = n_kf; kf.mvKeysUn.resize(n_kf); kf.mvKeys.resize(n_kf); kf.mDescriptors = desc_mat(kf_desc, n_kf); kf.mvpMapPoints.assign(n_kf, nullptr);
    std::vector<MapPoint> store(n_kf);
    std::map<MapPoint*, int> index_of;
    for (int i = 0; i < n_kf; ++i) {
        kf.mvKeysUn[i].angle = kf_angle[i]; kf.mvKeys[i].angle = kf_angle[i];
        if (kf_valid[i]) { kf.mvpMapPoints[i] = &store[i]; index_of[&store[i]] = i; }
    }
    for (int a = 0; a < n_nodes_kf; ++a)
        for (int j = kf_node_start[a]; j < kf_node_start[a + 1]; ++j) kf.mFeatVec.addFeature(kf_node_ids[a], (unsigned)kf_node_feat[j]);
    Frame F;
    F.N = n_f; F.Nleft = -1; F.mvKeysUn.resize(n_f); F.mvKeys.resize(n_f); F.mDescriptors = desc_mat(f_desc, n_f);
    for (int i = 0; i < n_f; ++i) { F.mvKeysUn[i].angle = f_angle[i]; F.mvKeys[i].angle = f_angle[i]; }
    for (int a = 0; a < n_nodes_f; ++a)
        for (int j = f_node_start[a]; j < f_node_start[a + 1]; ++j) F.mFeatVec.addFeature(f_node_ids[a], (unsigned)f_node_feat[j]);
    std::vector<MapPoint*> out;
    ORBmatcher matcher(nn_ratio, check_orientation != 0);
    const int nm = matcher.SearchByBoW(&kf, F, out);
    for (int i = 0; i < n_f; ++i) match[i] = (out[i] && index_of.count(out[i])) ? index_of[out[i]] : -1;
    return nm;
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist), arguments as orc_search_by_projection_reloc.
int ref_search_by_projection_reloc(const ref_frame_view* cur, const float cur_pose[7], int n, const uint8_t* valid,
                                   const float* xw, const uint8_t* mp_desc, const float* kf_angle, const float* mf_min_dist,
                                   const float* mf_max_dist, float th, int orb_dist, int check_orientation,
                                   const uint8_t* cur_occupied, int* match) {
    FrameHolder C(cur);
    C.F.SetPose(to_se3(cur_pose));
    for (int i = 0; i < cur->n; ++i) if (cur_occupied[i]) C.F.mvpMapPoints[i] = C.add_point();
    KeyFrame kf;
    kf.N = n; kf.mvKeysUn.resize(n); kf.mvKeys.resize(n); kf.mvpMapPoints.assign(n, nullptr);
    std::map<MapPoint*, int> index_of;
    std::set<MapPoint*> found;
    for (int i = 0; i < n; ++i) {
        kf.mvKeysUn[i].angle = kf_angle[i]; kf.mvKeys[i].angle = kf_angle[i];
        if (!valid[i]) continue;
        MapPoint* p = C.add_point();
        p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p->mDescriptor = desc_mat(mp_desc + 32 * i, 1);
        p->mfMinDistance = mf_min_dist[i]; p->mfMaxDistance = mf_max_dist[i];
        kf.mvpMapPoints[i] = p; index_of[p] = i;
    }
    ORBmatcher matcher(0.9f, check_orientation != 0);
    const int nm = matcher.SearchByProjection(C.F, &kf, found, th, orb_dist);
    for (int i = 0; i < cur->n; ++i) {
        MapPoint* p = C.F.mvpMapPoints[i];
        if (p && index_of.count(p)) match[i] = index_of[p];
        else if (!p && cur_occupied[i]) match[i] = -2;
        else match[i] = -1;
    }
    return nm;
}

// Optimizer::PoseOptimization(Frame*), arguments as orc_pose_optimize: one keypoint per edge, keypoint i on its own pyramid level i
// so that mvInvLevelSigma2[kpUn.octave] is the caller's inv_sigma2[i].
int ref_pose_optimize(const float pose_in[7], int n, const float* xw, const float* obs, const float* inv_sigma2,
                      const uint8_t* stereo, float fx, float fy, float cx, float cy, float bf,
                      float pose_out[7], uint8_t* outlier) {
    Frame F;
    Pinhole cam(std::vector<float>{fx, fy, cx, cy});
    F.N = n; F.Nleft = -1; F.mpCamera = &cam; F.mpCamera2 = nullptr;
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy; F.mbf = bf;
    F.mvKeysUn.resize(n); F.mvuRight.resize(n); F.mvInvLevelSigma2.resize(n); F.mvbOutlier.assign(n, false);
    std::vector<MapPoint> store(n);
    F.mvpMapPoints.resize(n);
    for (int i = 0; i < n; ++i) {
        F.mvKeysUn[i].pt = cv::Point2f(obs[3 * i], obs[3 * i + 1]); F.mvKeysUn[i].octave = i;
        F.mvuRight[i] = stereo[i] ? obs[3 * i + 2] : -1.0f;
        F.mvInvLevelSigma2[i] = inv_sigma2[i];
        store[i].mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        F.mvpMapPoints[i] = &store[i];
    }
    F.mvKeys = F.mvKeysUn;
    F.SetPose(to_se3(pose_in));
    const int r = Optimizer::PoseOptimization(&F);
    const Sophus::SE3f T = F.GetPose();
    pose_out[0] = T.unit_quaternion().x(); pose_out[1] = T.unit_quaternion().y(); pose_out[2] = T.unit_quaternion().z(); pose_out[3] = T.unit_quaternion().w();
    pose_out[4] = T.translation()(0); pose_out[5] = T.translation()(1); pose_out[6] = T.translation()(2);
    for (int i = 0; i < n; ++i) outlier[i] = F.mvbOutlier[i] ? 1 : 0;
    return r;
}

// Frame::ComputeStereoMatches, arguments as orc_stereo_matches (un-padded level images, row stride = width).
void ref_stereo_matches(int n_l, const void* kps_l_, const uint8_t* desc_l, int n_r, const void* kps_r_, const uint8_t* desc_r,
                        int n_levels, const float* scale, const float* inv_scale, const uint8_t* const* levels_l,
                        const uint8_t* const* levels_r, const int* lw, const int* lh, float mb, float mbf, float* depth, float* uright) {
    const RefKp* kl = (const RefKp*)kps_l_; const RefKp* kr = (const RefKp*)kps_r_;
    Frame F;
    ORBextractor exl, exr;
    for (int l = 0; l < n_levels; ++l) {
        exl.mvImagePyramid.push_back(cv::Mat(cv::Size(lw[l], lh[l]), CV_8U, (void*)levels_l[l]));
        exr.mvImagePyramid.push_back(cv::Mat(cv::Size(lw[l], lh[l]), CV_8U, (void*)levels_r[l]));
    }
    F.mpORBextractorLeft = &exl; F.mpORBextractorRight = &exr;
    F.N = n_l; F.mvKeys.resize(n_l); F.mvKeysRight.resize(n_r);
    for (int i = 0; i < n_l; ++i) F.mvKeys[i] = cv::KeyPoint(cv::Point2f(kl[i].x, kl[i].y), kl[i].size, kl[i].angle, kl[i].response, kl[i].octave, kl[i].class_id);
    for (int i = 0; i < n_r; ++i) F.mvKeysRight[i] = cv::KeyPoint(cv::Point2f(kr[i].x, kr[i].y), kr[i].size, kr[i].angle, kr[i].response, kr[i].octave, kr[i].class_id);
    F.mDescriptors = desc_mat(desc_l, n_l); F.mDescriptorsRight = desc_mat(desc_r, n_r);
    F.mvScaleFactors.assign(scale, scale + n_levels); F.mvInvScaleFactors.assign(inv_scale, inv_scale + n_levels);
    F.mb = mb; F.mbf = mbf;
    F.ComputeStereoMatches();
    for (int i = 0; i < n_l; ++i) { depth[i] = F.mvDepth[i]; uright[i] = F.mvuRight[i]; }
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:1074-1095): depth image H x W float32, distorted + undistorted keypoint positions.
void ref_stereo_from_rgbd(int n, const float* kp_xy, const float* kp_un_xy, const float* depth_map, int w, int h, float mbf, float* depth, float* uright) {
    Frame F;
    F.N = n; F.mvKeys.resize(n); F.mvKeysUn.resize(n); F.mbf = mbf;
    for (int i = 0; i < n; ++i) { F.mvKeys[i].pt = cv::Point2f(kp_xy[2 * i], kp_xy[2 * i + 1]); F.mvKeysUn[i].pt = cv::Point2f(kp_un_xy[2 * i], kp_un_xy[2 * i + 1]); }
    cv::Mat im(cv::Size(w, h), CV_32F, (void*)depth_map);
    F.ComputeStereoFromRGBD(im);
    for (int i = 0; i < n; ++i) { depth[i] = F.mvDepth[i]; uright[i] = F.mvuRight[i]; }
}

// Frame::UnprojectStereo for every keypoint (the map points the next frame's SearchByProjection projects): pose (qx..tz), out 3 floats
// per keypoint, ok[i] = return value.
void ref_unproject_stereo(const float pose[7], int n, const float* kp_un_xy, const float* depth, float fx, float fy, float cx, float cy, float* x3d, uint8_t* ok) {
    Frame F;
    F.N = n; F.mvKeysUn.resize(n); F.mvDepth.assign(depth, depth + n);
    for (int i = 0; i < n; ++i) F.mvKeysUn[i].pt = cv::Point2f(kp_un_xy[2 * i], kp_un_xy[2 * i + 1]);
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy; Frame::invfx = 1.0f / fx; Frame::invfy = 1.0f / fy;
    F.SetPose(to_se3(pose));
    for (int i = 0; i < n; ++i) {
        Eigen::Vector3f p; p.setZero();
        ok[i] = F.UnprojectStereo(i, p) ? 1 : 0;
        x3d[3 * i] = p(0); x3d[3 * i + 1] = p(1); x3d[3 * i + 2] = p(2);
    }
}

// MapPoint::ComputeDistinctiveDescriptors for a batch of map points: obs_start[n+1] offsets into desc (rows of 32 bytes);
// best[i] = index (within the point's observations) of the descriptor it keeps.
void ref_distinctive_descriptors(int n, const int* obs_start, const uint8_t* desc, int* best) {
    for (int i = 0; i < n; ++i) {
        const int m = obs_start[i + 1] - obs_start[i];
        best[i] = -1;
        if (m == 0) continue;
        std::vector<KeyFrame> kfs(m);
        MapPoint p;
        // std::map<KeyFrame*, ...> iterates in pointer order: the vector's storage is ascending, so observation j is visited j-th
        for (int j = 0; j < m; ++j) { kfs[j].mDescriptors = desc_mat(desc + 32 * (size_t)(obs_start[i] + j), 1); p.mObservations[&kfs[j]] = std::tuple<int, int>(0, -1); }
        p.ComputeDistinctiveDescriptors();
        for (int j = 0; j < m; ++j) if (memcmp(p.mDescriptor.data, desc + 32 * (size_t)(obs_start[i] + j), 32) == 0) { best[i] = j; break; }
    }
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:907-1146) with
// Pinhole::epipolarConstrain; key-frame arrays as orc_search_for_triangulation.  The reference derives the epipole and (inside
// epipolarConstrain, per pair) F12 from the two poses: both are returned so that the restatement can be fed the same numbers.
int ref_search_for_triangulation(const float T1w[7], const float T2w[7], float fx, float fy, float cx, float cy,
                                 int n1, const uint8_t* desc1, const void* keys1_, const uint8_t* has_mp1, const float* uright1,
                                 int nn1, const uint32_t* node_ids1, const int* node_start1, const int* node_feat1,
                                 int n2, const uint8_t* desc2, const void* keys2_, const uint8_t* has_mp2, const float* uright2,
                                 int nn2, const uint32_t* node_ids2, const int* node_start2, const int* node_feat2,
                                 int n_levels, const float* scale_factors2, const float* level_sigma2_2,
                                 int only_stereo, int coarse, int check_orientation, int32_t* match12, float* F12_out, float* ep_out) {
    Pinhole cam(std::vector<float>{fx, fy, cx, cy});
    MapPoint occupied;
    auto fill = [&](KeyFrame& kf, const float T[7], int n, const uint8_t* desc, const void* keys_, const uint8_t* has_mp, const float* uright,
                    int nn, const uint32_t* ids, const int* start, const int* feat) {
        const RefKp* k = (const RefKp*)keys_;
        kf.N = n; kf.NLeft = -1; kf.mpCamera = &cam; kf.mpCamera2 = nullptr;
        kf.mvKeysUn.resize(n); kf.mvpMapPoints.assign(n, nullptr); kf.mvuRight.assign(uright, uright + n); kf.mDescriptors = desc_mat(desc, n);
        for (int i = 0; i < n; ++i) {
            kf.mvKeysUn[i] = cv::KeyPoint(cv::Point2f(k[i].x, k[i].y), k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id);
            if (has_mp[i]) kf.mvpMapPoints[i] = &occupied;
        }
        kf.mvKeys = kf.mvKeysUn;
        for (int a = 0; a < nn; ++a) for (int j = start[a]; j < start[a + 1]; ++j) kf.mFeatVec.addFeature(ids[a], (unsigned)feat[j]);
        kf.SetPose(to_se3(T));
    };
    KeyFrame kf1, kf2;
    fill(kf1, T1w, n1, desc1, keys1_, has_mp1, uright1, nn1, node_ids1, node_start1, node_feat1);
    fill(kf2, T2w, n2, desc2, keys2_, has_mp2, uright2, nn2, node_ids2, node_start2, node_feat2);
    kf2.mvScaleFactors.assign(scale_factors2, scale_factors2 + n_levels); kf2.mvLevelSigma2.assign(level_sigma2_2, level_sigma2_2 + n_levels);
    kf1.mvScaleFactors = kf2.mvScaleFactors; kf1.mvLevelSigma2 = kf2.mvLevelSigma2;
    {   // what the function computes at its top (:914-931) and epipolarConstrain per pair (Pinhole.cpp:108-112)
        const Sophus::SE3f T12 = kf1.GetPose() * kf2.GetPoseInverse();
        const Eigen::Matrix3f R12 = T12.rotationMatrix(); const Eigen::Vector3f t12 = T12.translation();
        const Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12), K1 = cam.toK_(), K2 = cam.toK_();
        const Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F12_out[3 * r + c] = F12(r, c);
        const Eigen::Vector3f C2 = kf2.GetPose() * kf1.GetCameraCenter();
        const Eigen::Vector2f ep = cam.project(C2);
        ep_out[0] = ep(0); ep_out[1] = ep(1);
    }
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(0.6f, check_orientation != 0);
    const int nm = matcher.SearchForTriangulation(&kf1, &kf2, pairs, only_stereo != 0, coarse != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (const auto& pr : pairs) match12[pr.first] = (int32_t)pr.second;
    return nm;
}

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>&, th, bRight = false) (src/ORBmatcher.cc:1148-1330): the key frame's
// members come from the frame view (KeyFrame = copy of the Frame's keypoints, grid, scale tables; src/KeyFrame.cc ctor).  Every
// key-frame slot holds a placeholder map point without observations, so each fused point i ends in `slot.Replace(point i)`
// (:1313-1318): best_idx[i] = that slot, -1 when the point was not fused (no candidate with bestDist <= TH_LOW).  Ow_out = the camera
// centre the function uses (pKF->GetCameraCenter()).
int ref_fuse(const ref_frame_view* kfv, const float Tcw[7], int n, const uint8_t* valid, const float* xw, const float* normal,
             const float* mf_min_dist, const float* mf_max_dist, const uint8_t* mp_desc, float th, int* best_idx, float* Ow_out) {
    FrameHolder h(kfv);
    KeyFrame kf;
    kf.N = kfv->n; kf.NLeft = -1; kf.mpCamera = &h.cam; kf.mpCamera2 = nullptr;
    kf.mvKeysUn = h.F.mvKeysUn; kf.mvKeys = h.F.mvKeys; kf.mvuRight = h.F.mvuRight; kf.mDescriptors = h.F.mDescriptors;
    kf.fx = kfv->fx; kf.fy = kfv->fy; kf.cx = kfv->cx; kf.cy = kfv->cy; kf.mbf = kfv->bf;
    kf.mnMinX = (int)Frame::mnMinX; kf.mnMinY = (int)Frame::mnMinY; kf.mnMaxX = (int)Frame::mnMaxX; kf.mnMaxY = (int)Frame::mnMaxY;
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mnScaleLevels = h.F.mnScaleLevels; kf.mfLogScaleFactor = h.F.mfLogScaleFactor;
    kf.mvScaleFactors = h.F.mvScaleFactors; kf.mvLevelSigma2 = h.F.mvLevelSigma2; kf.mvInvLevelSigma2 = h.F.mvInvLevelSigma2;
    kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i) for (int j = 0; j < FRAME_GRID_ROWS; ++j) kf.mGrid[i][j] = h.F.mGrid[i][j];
    kf.SetPose(to_se3(Tcw));
    const Eigen::Vector3f Ow = kf.GetCameraCenter();
    Ow_out[0] = Ow(0); Ow_out[1] = Ow(1); Ow_out[2] = Ow(2);
    std::vector<MapPoint> slots(kfv->n);
    kf.mvpMapPoints.resize(kfv->n);
    for (int i = 0; i < kfv->n; ++i) { slots[i].nObs = 0; kf.mvpMapPoints[i] = &slots[i]; }
    std::vector<MapPoint> pts(n);
    std::vector<MapPoint*> vp(n, nullptr);
    for (int i = 0; i < n; ++i) {
        best_idx[i] = -1;
        if (!valid[i]) continue;
        pts[i].mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        pts[i].mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        pts[i].mfMinDistance = mf_min_dist[i]; pts[i].mfMaxDistance = mf_max_dist[i];
        pts[i].mDescriptor = desc_mat(mp_desc + 32 * i, 1);
        pts[i].nObs = 1;
        vp[i] = &pts[i];
    }
    ORBmatcher matcher(0.6f, true);
    const int nf = matcher.Fuse(&kf, vp, th, false);
    for (int s = 0; s < kfv->n; ++s)
        for (MapPoint* p : slots[s].replaced_by) best_idx[(int)(p - pts.data())] = s;
    return nf;
}

// Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1116-1499) over the reference's g2o (Schur-complement block solver, LM), arguments
// as orc_local_bundle_adjustment.  The key-frame graph is built so that the function finds: the non-fixed poses as the local key frames
// (the first one plays pKF, the others its covisible key frames), the fixed poses as the "fixed cameras" that observe local points.
// edge_erase[e] = the (key frame, map point) pair of edge e was put on vToErase.  Returns 0 (the LM iteration count is internal).
int ref_local_bundle_adjustment(int n_poses, const float* poses, const uint8_t* pose_fixed, int n_points, const float* points, int n_edges,
                                const int32_t* e_point, const int32_t* e_pose, const float* obs, const uint8_t* stereo,
                                const float* inv_sigma2, float fx, float fy, float cx, float cy, float bf, int iterations,
                                float* poses_out, float* points_out, uint8_t* edge_erase, double* final_chi2) {
    (void)iterations; if (final_chi2) *final_chi2 = 0;
    Pinhole cam(std::vector<float>{fx, fy, cx, cy});
    Map map;
    std::vector<KeyFrame> kfs(n_poses);
    std::vector<MapPoint> mps(n_points);
    for (int i = 0; i < n_poses; ++i) {
        KeyFrame& k = kfs[i];
        k.mnId = (unsigned long)i; k.mpMap = &map; k.mpCamera = &cam; k.mpCamera2 = nullptr; k.NLeft = -1;
        k.fx = fx; k.fy = fy; k.cx = cx; k.cy = cy; k.mbf = bf;
        k.SetPose(to_se3(poses + 7 * i));
    }
    for (int j = 0; j < n_points; ++j) {
        mps[j].mnId = (unsigned long)j; mps[j].mpMap = &map;
        mps[j].mWorldPos = Eigen::Vector3f(points[3 * j], points[3 * j + 1], points[3 * j + 2]);
    }
    for (int e = 0; e < n_edges; ++e) {          // one keypoint per observation, on its own pyramid level (so that mvInvLevelSigma2[octave] = inv_sigma2[e])
        KeyFrame& k = kfs[e_pose[e]]; MapPoint& p = mps[e_point[e]];
        const int idx = (int)k.mvKeysUn.size();
        cv::KeyPoint kp; kp.pt = cv::Point2f(obs[3 * e], obs[3 * e + 1]); kp.octave = idx;
        k.mvKeysUn.push_back(kp); k.mvuRight.push_back(stereo[e] ? obs[3 * e + 2] : -1.0f); k.mvInvLevelSigma2.push_back(inv_sigma2[e]);
        k.mvpMapPoints.push_back(&p);
        p.mObservations[&k] = std::tuple<int, int>(idx, -1);
    }
    KeyFrame* pKF = nullptr;
    for (int i = 0; i < n_poses; ++i) if (!pose_fixed[i]) { if (!pKF) pKF = &kfs[i]; else pKF->mvpOrderedConnectedKeyFrames.push_back(&kfs[i]); }
    for (int i = 0; i < n_poses; ++i) { std::memcpy(poses_out + 7 * i, poses + 7 * i, 7 * sizeof(float)); }
    std::memcpy(points_out, points, (size_t)n_points * 3 * sizeof(float));
    std::memset(edge_erase, 0, (size_t)n_edges);
    if (!pKF) return 0;
    pKF->mnId = (unsigned long)n_poses + 7;      // mnBALocalForKF markers compare against pKF->mnId: make it non-zero and unique
    int nf = 0, no = 0, nm = 0, ne = 0;
    Optimizer::LocalBundleAdjustment(pKF, nullptr, &map, nf, no, nm, ne);
    for (int i = 0; i < n_poses; ++i) {
        const Sophus::SE3f T = kfs[i].GetPose();
        poses_out[7 * i] = T.unit_quaternion().x(); poses_out[7 * i + 1] = T.unit_quaternion().y(); poses_out[7 * i + 2] = T.unit_quaternion().z();
        poses_out[7 * i + 3] = T.unit_quaternion().w();
        for (int c = 0; c < 3; ++c) poses_out[7 * i + 4 + c] = T.translation()(c);
    }
    for (int j = 0; j < n_points; ++j) for (int c = 0; c < 3; ++c) points_out[3 * j + c] = mps[j].mWorldPos(c);
    for (int e = 0; e < n_edges; ++e) {
        const KeyFrame& k = kfs[e_pose[e]];
        for (MapPoint* p : k.erased) if (p == &mps[e_point[e]]) edge_erase[e] = 1;
    }
    return 0;
}

}  // extern "C"
