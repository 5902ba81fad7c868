// Reference-side binding of librgbl_b200.so (glue only): helpers shared by the replacement translation units below.
//   shim/ORBextractor.h        replaces include/ORBextractor.h + src/ORBextractor.cc (the whole class)
//   shim/DepthModule_b200.cc   replaces ONE function of src/DepthModule.cc:   DepthModule::CalculateDepthFromPcd (:50-79)
//   shim/ORBmatcher_b200.cc    replaces FIVE functions of src/ORBmatcher.cc:  DescriptorDistance (:2058-2074), SearchByProjection
//                              (Frame&, const vector<MapPoint*>&, ...) (:43-213), SearchByBoW(KeyFrame*, Frame&, ...) (:223-425),
//                              SearchByProjection(Frame&, const Frame&, ...) (:1676-1887), SearchByProjection(Frame&, KeyFrame*, ...) (:1889-2010)
//   shim/Optimizer_b200.cc     replaces ONE function of src/Optimizer.cc:     Optimizer::PoseOptimization (:814-1114)
// Every replaced function keeps the reference's signature, so Frame.cc / Tracking.cc / System.cc compile and behave unchanged; the
// functions not listed (LocalMapping / LoopClosing matchers, the other optimisers) stay the reference's CPU code.  See INTEGRATION.md
// for the CMake lines.  tests/test_shim.py compiles these files against stand-ins of OpenCV / Eigen / Sophus and the ORB-SLAM3 class
// shells (oracle/ref_shim), links them with librgbl_b200.so and drives them with Frame-constructor- and Tracking-shaped caller code.
#pragma once
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "rgbl_b200.h"

namespace rgbl_shim {

// The device context of the tracking thread = the context of the (left) ORBextractor that extracted last.  The reference creates one
// ORBextractor per camera in the Tracking constructor (src/Tracking.cc:595-601); DepthModule, ORBmatcher and Optimizer calls of that
// thread run on the same device context.
inline rgbl_ctx*& shared_context() { static rgbl_ctx* ctx = nullptr; return ctx; }

inline void check(rgbl_ctx* ctx, int rc) {
    if (rc != RGBL_OK) throw std::runtime_error(std::string("librgbl_b200: ") + rgbl_last_error(ctx));
}
inline rgbl_ctx* need_context() {
    rgbl_ctx* c = shared_context();
    if (!c) throw std::runtime_error("librgbl_b200: no device context yet (ORBextractor::operator() creates it with the first image)");
    return c;
}

// Sophus::SE3f <-> (qx, qy, qz, qw, tx, ty, tz)
template <class SE3> inline void to_pose7(const SE3& T, float p[7]) {
    const auto& q = T.unit_quaternion(); const auto& t = T.translation();
    p[0] = q.x(); p[1] = q.y(); p[2] = q.z(); p[3] = q.w(); p[4] = t(0); p[5] = t(1); p[6] = t(2);
}

// the members of a Frame the device matchers read (include/rgbl_b200.h: rgbl_frame_view)
template <class FrameT> inline rgbl_frame_view make_view(const FrameT& F) {
    rgbl_frame_view v{};
    v.n = F.N;
    v.keys_un = reinterpret_cast<const rgbl_keypoint*>(F.mvKeysUn.data());
    v.uright = F.mvuRight.data();
    v.desc = F.mDescriptors.data;
    v.min_x = FrameT::mnMinX; v.max_x = FrameT::mnMaxX; v.min_y = FrameT::mnMinY; v.max_y = FrameT::mnMaxY;
    v.n_levels = F.mnScaleLevels; v.scale_factors = F.mvScaleFactors.data();
    v.fx = FrameT::fx; v.fy = FrameT::fy; v.cx = FrameT::cx; v.cy = FrameT::cy; v.bf = F.mbf;
    v.log_scale_factor = F.mfLogScaleFactor;
    return v;
}

}  // namespace rgbl_shim
