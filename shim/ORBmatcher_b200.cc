// Replacement of FIVE functions of the reference's src/ORBmatcher.cc, same signatures (include/ORBmatcher.h:40-69): the tracking-thread
// matchers.  Each gathers the MapPoint / Frame state it needs through the same getters the reference calls (GetWorldPos, GetDescriptor,
// Observations, isBad ...: they take the reference's mutexes), calls the C ABI and writes the matches into the same output the
// reference writes.  Remove the five definitions from src/ORBmatcher.cc and add this file; every other ORBmatcher method (the
// LocalMapping / LoopClosing matchers that run concurrently on other threads) stays the reference's CPU code.
#include "ORBmatcher.h"

#include <cstring>
#include <map>

#include "rgbl_shim_common.h"

namespace ORB_SLAM3 {

namespace {
inline void copy_desc(const cv::Mat& d, uint8_t* dst) { std::memcpy(dst, d.ptr<uint8_t>(0), 32); }

// CSR of a DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>): ascending node ids, features in vector order
struct FeatCsr { std::vector<uint32_t> ids; std::vector<int32_t> start, feat; };
inline FeatCsr to_csr(const DBoW2::FeatureVector& fv) {
    FeatCsr c; c.start.push_back(0);
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        c.ids.push_back((uint32_t)it->first);
        for (size_t k = 0; k < it->second.size(); ++k) c.feat.push_back((int32_t)it->second[k]);
        c.start.push_back((int32_t)c.feat.size());
    }
    return c;
}
template <class T> inline const T* ptr_or(const std::vector<T>& v, const T* dummy) { return v.empty() ? dummy : v.data(); }
}  // namespace

int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {             // src/ORBmatcher.cc:2058-2074
    return rgbl_descriptor_distance(a.ptr<uint8_t>(0), b.ptr<uint8_t>(0));
}

// src/ORBmatcher.cc:1676-1887 (frames with Nleft == -1)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
    if (CurrentFrame.Nleft != -1) throw std::runtime_error("librgbl_b200: two-camera frames keep the reference's SearchByProjection");
    const int NL = LastFrame.N, NC = CurrentFrame.N;
    std::vector<uint8_t> valid(NL > 0 ? NL : 1, 0), obs(NL > 0 ? NL : 1, 0), desc((size_t)(NL > 0 ? NL : 1) * 32, 0), state(NC > 0 ? NC : 1, 0);
    std::vector<float> xw((size_t)(NL > 0 ? NL : 1) * 3, 0.f), ang(NL > 0 ? NL : 1, 0.f);
    std::vector<int32_t> oct(NL > 0 ? NL : 1, 0), match(NC > 0 ? NC : 1, -1);
    for (int i = 0; i < NL; ++i) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        oct[i] = LastFrame.mvKeysUn[i].octave; ang[i] = LastFrame.mvKeysUn[i].angle;
        if (!pMP || LastFrame.mvbOutlier[i]) continue;                               // :1699-1701
        valid[i] = 1;
        const Eigen::Vector3f X = pMP->GetWorldPos();
        xw[3 * i] = X(0); xw[3 * i + 1] = X(1); xw[3 * i + 2] = X(2);
        copy_desc(pMP->GetDescriptor(), &desc[(size_t)32 * i]);
        obs[i] = pMP->Observations() > 0;
    }
    for (int i = 0; i < NC; ++i)
        if (MapPoint* p = CurrentFrame.mvpMapPoints[i]) state[i] = p->Observations() > 0 ? 1 : 2;       // :1761-1763
    const rgbl_frame_view fv = rgbl_shim::make_view(CurrentFrame);
    float Tc[7], Tl[7];
    rgbl_shim::to_pose7(CurrentFrame.GetPose(), Tc); rgbl_shim::to_pose7(LastFrame.GetPose(), Tl);
    int n = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_search_by_projection_last(ctx, &fv, Tc, Tl, NL, valid.data(), xw.data(), desc.data(), oct.data(), ang.data(), obs.data(),
                                                         th, bMono ? 1 : 0, mbCheckOrientation ? 1 : 0, state.data(), match.data(), &n));
    for (int i = 0; i < NC; ++i) {
        if (match[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[match[i]];
        else if (match[i] == -2) CurrentFrame.mvpMapPoints[i] = static_cast<MapPoint*>(NULL);           // rotation check, :1878
    }
    return n;
}

// src/ORBmatcher.cc:43-213
int ORBmatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints) {
    if (F.Nleft != -1) throw std::runtime_error("librgbl_b200: two-camera frames keep the reference's SearchByProjection");
    const int n = (int)vpMapPoints.size(), NC = F.N, nn = n > 0 ? n : 1;
    std::vector<uint8_t> in_view(nn, 0), obs(nn, 0), desc((size_t)nn * 32, 0), state(NC > 0 ? NC : 1, 0);
    std::vector<float> px(nn, 0.f), py(nn, 0.f), pxr(nn, 0.f), depth(nn, 0.f), vc(nn, 0.f);
    std::vector<int32_t> level(nn, 0), match(NC > 0 ? NC : 1, -1);
    for (int i = 0; i < n; ++i) {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP->mbTrackInView || pMP->isBad()) continue;                            // :53-60
        in_view[i] = 1;
        px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR; depth[i] = pMP->mTrackDepth;
        level[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
        copy_desc(pMP->GetDescriptor(), &desc[(size_t)32 * i]);
        obs[i] = pMP->Observations() > 0;
    }
    for (int i = 0; i < NC; ++i)
        if (MapPoint* p = F.mvpMapPoints[i]) state[i] = p->Observations() > 0 ? 1 : 2;                  // :97-99
    const rgbl_frame_view fv = rgbl_shim::make_view(F);
    int nm = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_search_by_projection_local(ctx, &fv, n, in_view.data(), px.data(), py.data(), pxr.data(), depth.data(), level.data(), vc.data(),
                                                          desc.data(), obs.data(), th, mfNNratio, bFarPoints ? 1 : 0, thFarPoints, state.data(), match.data(), &nm));
    for (int i = 0; i < NC; ++i)
        if (match[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[match[i]];
    return nm;
}

// src/ORBmatcher.cc:223-425
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
    if (F.Nleft != -1) throw std::runtime_error("librgbl_b200: two-camera frames keep the reference's SearchByBoW");
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const int nk = (int)vpMapPointsKF.size(), nn = nk > 0 ? nk : 1;
    std::vector<uint8_t> kvalid(nn, 0);
    std::vector<float> kang(nn, 0.f), fang(F.N > 0 ? F.N : 1, 0.f);
    for (int i = 0; i < nk; ++i) {
        kang[i] = pKF->mvKeysUn[i].angle;
        MapPoint* p = vpMapPointsKF[i];
        kvalid[i] = (p && !p->isBad()) ? 1 : 0;                                       // :253-259
    }
    for (int i = 0; i < F.N; ++i) fang[i] = F.mvKeysUn[i].angle;
    const FeatCsr kc = to_csr(pKF->mFeatVec), fc = to_csr(F.mFeatVec);
    std::vector<int32_t> match(F.N > 0 ? F.N : 1, -1);
    int nm = 0;
    const uint32_t du = 0; const int32_t di = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_search_by_bow(ctx, nk, pKF->mDescriptors.ptr<uint8_t>(0), kang.data(), kvalid.data(), (int)kc.ids.size(), ptr_or(kc.ids, &du),
                                             kc.start.data(), ptr_or(kc.feat, &di), F.N, F.mDescriptors.ptr<uint8_t>(0), fang.data(), (int)fc.ids.size(),
                                             ptr_or(fc.ids, &du), fc.start.data(), ptr_or(fc.feat, &di), mfNNratio, mbCheckOrientation ? 1 : 0, match.data(), &nm));
    for (int i = 0; i < F.N; ++i)
        if (match[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[match[i]];
    return nm;
}

// src/ORBmatcher.cc:1889-2010
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist) {
    if (CurrentFrame.Nleft != -1) throw std::runtime_error("librgbl_b200: two-camera frames keep the reference's SearchByProjection");
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n = (int)vpMPs.size(), nn = n > 0 ? n : 1, NC = CurrentFrame.N;
    std::vector<uint8_t> valid(nn, 0), desc((size_t)nn * 32, 0), occ(NC > 0 ? NC : 1, 0);
    std::vector<float> xw((size_t)nn * 3, 0.f), ang(nn, 0.f), mn(nn, 0.f), mx(nn, 0.f);
    for (int i = 0; i < n; ++i) {
        ang[i] = pKF->mvKeysUn[i].angle;
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;                // :1906-1910
        valid[i] = 1;
        const Eigen::Vector3f X = pMP->GetWorldPos();
        xw[3 * i] = X(0); xw[3 * i + 1] = X(1); xw[3 * i + 2] = X(2);
        copy_desc(pMP->GetDescriptor(), &desc[(size_t)32 * i]);
        // the ABI takes mfMinDistance / mfMaxDistance themselves (MapPoint::PredictScale needs the raw mfMaxDistance, and 0.8f * / 1.2f * of
        // them cannot be undone exactly): two inline getters added to include/MapPoint.h, see INTEGRATION.md
        mn[i] = pMP->GetMinDistance(); mx[i] = pMP->GetMaxDistance();
    }
    for (int i = 0; i < NC; ++i) occ[i] = CurrentFrame.mvpMapPoints[i] ? 1 : 0;         // :1950-1951
    const rgbl_frame_view fv = rgbl_shim::make_view(CurrentFrame);
    float Tc[7];
    rgbl_shim::to_pose7(CurrentFrame.GetPose(), Tc);
    std::vector<int32_t> match(NC > 0 ? NC : 1, -1);
    int nm = 0;
    rgbl_ctx* ctx = rgbl_shim::need_context();
    rgbl_shim::check(ctx, rgbl_search_by_projection_reloc(ctx, &fv, Tc, n, valid.data(), xw.data(), desc.data(), ang.data(), mn.data(), mx.data(), th, ORBdist,
                                                          mbCheckOrientation ? 1 : 0, occ.data(), match.data(), &nm));
    for (int i = 0; i < NC; ++i) {
        if (match[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[match[i]];
        else if (match[i] == -2) CurrentFrame.mvpMapPoints[i] = NULL;
    }
    return nm;
}

}  // namespace ORB_SLAM3
