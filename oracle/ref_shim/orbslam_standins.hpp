// ORACLE SUPPORT (test infrastructure, NOT product code): stand-in declarations of the ORB-SLAM3 classes around the hot path, so that
// the reference's OWN function bodies compile unmodified from /root/reference:
//   * src/ORBmatcher.cc (whole file), src/OptimizableTypes.cpp (whole file), src/CameraModels/Pinhole.cpp (whole file, over the real
//     include/CameraModels/GeometricCamera.h and Pinhole.h);
//   * member functions of Frame / MapPoint / Optimizer extracted verbatim at build time (oracle/extract_ref.py -> oracle/_ref/gen/*.inc,
//     never committed): Frame::AssignFeaturesToGrid, isInFrustum, GetFeaturesInArea, PosInGrid, ComputeStereoMatches,
//     ComputeStereoFromRGBD, UnprojectStereo (src/Frame.cc), MapPoint::PredictScale, ComputeDistinctiveDescriptors (src/MapPoint.cc),
//     Optimizer::PoseOptimization (src/Optimizer.cc:814-1114).
// The classes below carry exactly the members those bodies touch, with the reference's names, types and signatures
// (include/Frame.h, include/MapPoint.h, include/KeyFrame.h); the real headers are switched off by pre-defining their include guards
// because they pull in the whole system (Atlas, IMU, vocabulary, serialisation, viewer ...).
#pragma once
#define FRAME_H
#define MAPPOINT_H
#define KEYFRAME_H
#define MAP_H
#define CONVERTER_H
#define GEOMETRIC_TOOLS_H
#define TwoViewReconstruction_H
#ifndef RGBL_B200_SHIM
#define ORBEXTRACTOR_H
#endif
#define OPTIMIZER_H
#define G2OTYPES_H

#include <climits>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <sophus/se3.hpp>
#include <sophus/sim3.hpp>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48      // include/Frame.h:46-47
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {
using namespace std;            // the reference's headers do this (include/Frame.h etc.), and its sources rely on it

class TwoViewReconstruction {   // only constructed by Pinhole::ReconstructWithTwoViews (monocular initialisation, off the path)
public:
    TwoViewReconstruction(const Eigen::Matrix3f&, float = 1.0, int = 200) {}
    bool Reconstruct(const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const std::vector<int>&, Sophus::SE3f&,
                     std::vector<cv::Point3f>&, std::vector<bool>&) { abort(); }
};
}  // namespace ORB_SLAM3

#include "CameraModels/GeometricCamera.h"       // the reference's own (its Converter.h / GeometricTools.h includes are guarded off above)
#include "CameraModels/Pinhole.h"

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;
class Map;

#ifndef RGBL_B200_SHIM            // the binding test compiles shim/ORBextractor.h and shim/Optimizer_b200.cc instead
class ORBextractor {            // ComputeStereoMatches reads the two pyramids (include/ORBextractor.h:83)
public:
    std::vector<cv::Mat> mvImagePyramid;
};
#else
class ORBextractor;
#endif

class MapPoint {
public:
    // --- include/MapPoint.h:171-179 (tracking variables)
    float mTrackProjX = -1, mTrackProjY = -1, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
    long unsigned int mnId = 0;
    long unsigned int mnBALocalForKF = 0, mnFuseCandidateForKF = 0;
    static std::mutex mGlobalMutex;

    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    void SetWorldPos(const Eigen::Vector3f& p) { mWorldPos = p; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }                   // src/MapPoint.cc:411-415
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }          // src/MapPoint.cc:502-506
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }          // src/MapPoint.cc:508-512
    float GetMinDistance() { return mfMinDistance; }                            // the two getters the binding adds to include/MapPoint.h
    float GetMaxDistance() { return mfMaxDistance; }
    int PredictScale(const float& currentDist, KeyFrame* pKF);                 // bodies: src/MapPoint.cc, extracted
    int PredictScale(const float& currentDist, Frame* pF);
    void ComputeDistinctiveDescriptors();
    std::map<KeyFrame*, std::tuple<int, int>> GetObservations() { return mObservations; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? mObservations[pKF] : std::tuple<int, int>(-1, -1); }   // src/MapPoint.cc:411-418
    // bookkeeping the fusion / BA code calls after its search: recorded, not modelled
    void Replace(MapPoint* pMP) { mpReplaced = pMP; replaced_by.push_back(pMP); }
    void AddObservation(KeyFrame* pKF, int idx) { mObservations[pKF] = std::tuple<int, int>(idx, -1); ++nObs; }
    void EraseObservation(KeyFrame* pKF) { mObservations.erase(pKF); }
    void UpdateNormalAndDepth() {}
    Map* GetMap() { return mpMap; }
    Map* mpMap = nullptr;

    Eigen::Vector3f mWorldPos, mNormalVector;
    cv::Mat mDescriptor;
    bool mbBad = false;
    int nObs = 1;
    float mfMinDistance = 0, mfMaxDistance = 0;
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    MapPoint* mpReplaced = nullptr;
    std::vector<MapPoint*> replaced_by;          // every Replace() call, in order (the driver reads the fusion outcome from it)
    std::mutex mMutexPos, mMutexFeatures;
};

class Frame {
public:
    // --- the members the matcher / optimiser / frustum / grid / stereo bodies read, include/Frame.h
    int N = 0, Nleft = -1, Nright = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float mbf = 0, mb = 0, mThDepth = 0;
    int mnScaleLevels = 0;
    float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    static float fx, fy, cx, cy, invfx, invfy;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mDistCoef;

    // --- pose (include/Frame.h:150-160, src/Frame.cc:521-569)
    Sophus::SE3<float> mTcw, mTrl, mTlr;
    Eigen::Matrix<float, 3, 3> mRwc, mRcw;
    Eigen::Matrix<float, 3, 1> mOw, mtcw;
    bool mbHasPose = false;
    void SetPose(const Sophus::SE3<float>& Tcw);                // bodies: src/Frame.cc, extracted
    void UpdatePoseMatrices();
    inline Sophus::SE3<float> GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    Sophus::SE3f GetRelativePoseTlr() { return mTlr; }

    // --- extracted bodies (src/Frame.cc)
    void AssignFeaturesToGrid();
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
    bool isInFrustumChecks(MapPoint*, float, bool = false) { abort(); }          // fisheye stereo rigs only (Nleft != -1)
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1, const bool bRight = false) const;
    void ComputeStereoMatches();
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);
    bool UnprojectStereo(const int& i, Eigen::Vector3f& x3D);
};

class KeyFrame {
public:
    long unsigned int mnId = 0;
    long unsigned int mnBALocalForKF = 0, mnBAFixedForKF = 0;
    int N = 0, NLeft = -1, NRight = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0, mThDepth = 0;
    int mnScaleLevels = 0;
    float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;          // const int in the reference (include/KeyFrame.h)
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
    std::vector<MapPoint*> mvpMapPoints;
    bool mbBad = false;
    Sophus::SE3f mTcw, mTwc, mTrl, mTlr;

    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTwc; }
    Eigen::Vector3f GetCameraCenter() { return mTwc.translation(); }
    Eigen::Matrix3f GetRotation() { return mTcw.rotationMatrix(); }
    Eigen::Vector3f GetTranslation() { return mTcw.translation(); }
    Sophus::SE3f GetRightPose() { return mTrl * mTcw; }
    Sophus::SE3f GetRightPoseInverse() { return mTwc * mTlr; }
    Eigen::Vector3f GetRightCameraCenter() { return (mTwc * mTlr).translation(); }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    Sophus::SE3f GetRelativePoseTlr() { return mTlr; }
    void SetPose(const Sophus::SE3f& Tcw) { mTcw = Tcw; mTwc = Tcw.inverse(); }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    void EraseMapPointMatch(const int& idx) { mvpMapPoints[idx] = nullptr; }
    void EraseMapPointMatch(MapPoint* pMP) { erased.push_back(pMP); for (auto& q : mvpMapPoints) if (q == pMP) q = nullptr; }      // src/KeyFrame.cc:365-379
    std::vector<MapPoint*> erased;               // the driver reads the outcome of LocalBundleAdjustment from it
    Map* GetMap() { return mpMap; }
    Map* mpMap = nullptr;
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;
    bool isBad() { return mbBad; }
    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   // src/KeyFrame.cc:750-753
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;      // body: src/KeyFrame.cc, extracted
    bool UnprojectStereo(int, Eigen::Vector3f&) { abort(); }
};

class Map {
public:
    void IncreaseChangeIndex() {}
    long unsigned int GetMaxKFid() { return 0; }
    long unsigned int GetInitKFid() { return mnInitKFid; }
    bool IsInertial() { return false; }
    bool isImuInitialized() { return false; }
    long unsigned int mnInitKFid = ~0ul;
    std::set<long unsigned int> msOptKFs, msFixedKFs;
    std::mutex mMutexMapUpdate;
};

class Verbose {                 // include/System.h:47-75
public:
    enum eLevel { VERBOSITY_QUIET = 0, VERBOSITY_NORMAL = 1, VERBOSITY_VERBOSE = 2, VERBOSITY_VERY_VERBOSE = 3, VERBOSITY_DEBUG = 4 };
    static void PrintMess(std::string, eLevel) {}
};

class Optimizer {
public:
    static int PoseOptimization(Frame* pFrame);                  // body: src/Optimizer.cc:814-1114, extracted
    static void LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF, int& num_MPs, int& num_edges);   // :1116-1499
};

}  // namespace ORB_SLAM3
