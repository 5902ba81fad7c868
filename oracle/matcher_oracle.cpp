// ORACLE (test infrastructure, NOT product code): CPU restatement of the tracking-thread matchers.
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:475-506, 815-825, 747-813
//   ORBmatcher::DescriptorDistance                                src/ORBmatcher.cc:2058-2074
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, mono) src/ORBmatcher.cc:1676-1887  (Nleft == -1 path)
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ..) src/ORBmatcher.cc:43-213      (Nleft == -1 path)
//   ORBmatcher::ComputeThreeMaxima                                 src/ORBmatcher.cc:2012-2053
//   Frame::isInFrustum + MapPoint::PredictScale                    src/Frame.cc:602-664, src/MapPoint.cc:531-545
// MapPoint objects are replaced by flat per-point arrays (what the C-ABI shim gathers).  Float32
// arithmetic follows the reference expression by expression (Sophus SE3f point action, Pinhole::project),
// canonical no-FMA semantics.  Only tests/, smoke() and bench.py's CPU legs may load this library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int GRID_COLS = 64, GRID_ROWS = 48;      // include/Frame.h:46-47
constexpr int TH_HIGH = 100, HISTO_LENGTH = 30;    // src/ORBmatcher.cc:35-37

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };

struct FrameView {          // the members of Frame the matchers read (Nleft == -1)
    int n;
    const KeyPoint* keys_un;    // mvKeysUn
    const float* uright;        // mvuRight
    const uint8_t* desc;        // mDescriptors, n x 32
    float min_x, max_x, min_y, max_y;     // mnMinX ...
    float inv_w, inv_h;         // mfGridElementWidthInv / HeightInv
    int n_levels;
    const float* scale_factors; // mvScaleFactors
    float fx, fy, cx, cy, bf, mb;
    float log_scale_factor;
};

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    void build(const FrameView& f) {
        for (int i = 0; i < f.n; ++i) {
            const int px = (int)std::round((f.keys_un[i].x - f.min_x) * f.inv_w);
            const int py = (int)std::round((f.keys_un[i].y - f.min_y) * f.inv_h);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
};

void features_in_area(const FrameView& f, const Grid& g, float x, float y, float r, int min_level, int max_level,
                      std::vector<int>& out) {
    out.clear();
    const int min_cx = std::max(0, (int)std::floor((x - f.min_x - r) * f.inv_w));
    if (min_cx >= GRID_COLS) return;
    const int max_cx = std::min(GRID_COLS - 1, (int)std::ceil((x - f.min_x + r) * f.inv_w));
    if (max_cx < 0) return;
    const int min_cy = std::max(0, (int)std::floor((y - f.min_y - r) * f.inv_h));
    if (min_cy >= GRID_ROWS) return;
    const int max_cy = std::min(GRID_ROWS - 1, (int)std::ceil((y - f.min_y + r) * f.inv_h));
    if (max_cy < 0) return;
    const bool check_levels = (min_level > 0) || (max_level >= 0);
    for (int ix = min_cx; ix <= max_cx; ++ix)
        for (int iy = min_cy; iy <= max_cy; ++iy)
            for (int idx : g.cell[ix][iy]) {
                const KeyPoint& kp = f.keys_un[idx];
                if (check_levels) {
                    if (kp.octave < min_level) continue;
                    if (max_level >= 0 && kp.octave > max_level) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(idx);
            }
}

int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t pa, pb;
        std::memcpy(&pa, a + 4 * i, 4); std::memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// Sophus::SE3f (unit quaternion x,y,z,w + translation) acting on a point: so3.hpp:358-366, se3.hpp:321-324
struct Pose { float qx, qy, qz, qw, tx, ty, tz; };

inline void rotate(const Pose& T, const float p[3], float out[3]) {
    float uv[3] = {T.qy * p[2] - T.qz * p[1], T.qz * p[0] - T.qx * p[2], T.qx * p[1] - T.qy * p[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    const float c[3] = {T.qy * uv[2] - T.qz * uv[1], T.qz * uv[0] - T.qx * uv[2], T.qx * uv[1] - T.qy * uv[0]};
    for (int i = 0; i < 3; ++i) out[i] = (p[i] + T.qw * uv[i]) + c[i];
}
inline void transform(const Pose& T, const float p[3], float out[3]) {
    float r[3];
    rotate(T, p, r);
    out[0] = r[0] + T.tx; out[1] = r[1] + T.ty; out[2] = r[2] + T.tz;
}
// Eigen 3.3 (Ubuntu 20.04's libeigen3-dev, the reference's Dockerfile) adds the terms of a fixed-size dot product / squared norm /
// matrix-product coefficient with its unrolled scalar reduction (Core/Redux.h redux_novec_unroller: halves, recursively):
inline float sum3(float a, float b, float c) { return a + (b + c); }
inline float sum4(float a, float b, float c, float d) { return (a + b) + (c + d); }
// SO3f::inverse() = SO3f(unit_quaternion().conjugate()): the quaternion constructor normalises (so3.hpp:229-231, 481-487, 297-303)
inline Pose inverse_rotation(const Pose& T) {
    Pose inv = T; inv.qx = -T.qx; inv.qy = -T.qy; inv.qz = -T.qz;
    const float length = std::sqrt(sum4(inv.qx * inv.qx, inv.qy * inv.qy, inv.qz * inv.qz, inv.qw * inv.qw));
    inv.qx /= length; inv.qy /= length; inv.qz /= length; inv.qw /= length;
    return inv;
}
// Tcw.inverse().translation() = invR * (translation() * -1)   (se3.hpp:208-211)
inline void inverse_translation(const Pose& T, float out[3]) {
    const Pose inv = inverse_rotation(T);
    const float nt[3] = {T.tx * -1.f, T.ty * -1.f, T.tz * -1.f};
    rotate(inv, nt, out);
}

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

}  // namespace

extern "C" {

struct orc_frame_view {
    int n;
    const void* keys_un; const float* uright; const uint8_t* desc;
    float min_x, max_x, min_y, max_y;
    int n_levels; const float* scale_factors;
    float fx, fy, cx, cy, bf;
    float log_scale_factor;
};

static FrameView to_view(const orc_frame_view* v) {
    FrameView f;
    f.n = v->n; f.keys_un = (const KeyPoint*)v->keys_un; f.uright = v->uright; f.desc = v->desc;
    f.min_x = v->min_x; f.max_x = v->max_x; f.min_y = v->min_y; f.max_y = v->max_y;
    f.inv_w = static_cast<float>(GRID_COLS) / static_cast<float>(v->max_x - v->min_x);     // src/Frame.cc:351-352
    f.inv_h = static_cast<float>(GRID_ROWS) / static_cast<float>(v->max_y - v->min_y);
    f.n_levels = v->n_levels; f.scale_factors = v->scale_factors;
    f.fx = v->fx; f.fy = v->fy; f.cx = v->cx; f.cy = v->cy; f.bf = v->bf;
    f.mb = v->bf / v->fx;                                                                    // src/Frame.cc:360
    f.log_scale_factor = v->log_scale_factor;
    return f;
}

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

int orc_features_in_area(const orc_frame_view* fv, float x, float y, float r, int min_level, int max_level, int* out, int cap) {
    FrameView f = to_view(fv);
    Grid g; g.build(f);
    std::vector<int> v;
    features_in_area(f, g, x, y, r, min_level, max_level, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono).  Last-frame map points are given as flat arrays over
// i = 0..n_last-1: valid[i] = (mvpMapPoints[i] != NULL && !mvbOutlier[i]), xw (3 floats), desc (32 B, GetDescriptor()),
// octave[i] and angle[i] of the last-frame keypoint, obs_pos[i] = (Observations() > 0).
// cur_state[i2] on entry: 0 free, 1 holds a point with Observations()>0, 2 holds a point with 0 observations.
// match[i2] on exit: >= 0 index i assigned, -1 untouched, -2 cleared by the rotation check.  Returns nmatches.
int orc_search_by_projection_last(const orc_frame_view* cur, const float cur_pose[7], const float last_pose[7],
                                  int n_last, const uint8_t* valid, const float* xw, const uint8_t* mp_desc,
                                  const int* last_octave, const float* last_angle, const uint8_t* obs_pos,
                                  float th, int mono, int check_orientation, const uint8_t* cur_state, int* match) {
    FrameView F = to_view(cur);
    Grid g; g.build(F);
    Pose Tcw = {cur_pose[0], cur_pose[1], cur_pose[2], cur_pose[3], cur_pose[4], cur_pose[5], cur_pose[6]};
    Pose Tlw = {last_pose[0], last_pose[1], last_pose[2], last_pose[3], last_pose[4], last_pose[5], last_pose[6]};
    int nmatches = 0;
    std::vector<int> rot_hist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    float twc[3], tlc[3];
    inverse_translation(Tcw, twc);
    transform(Tlw, twc, tlc);
    const bool forward = tlc[2] > F.mb && !mono;
    const bool backward = -tlc[2] > F.mb && !mono;
    std::vector<uint8_t> state(cur_state, cur_state + F.n);
    for (int i = 0; i < F.n; ++i) match[i] = -1;
    std::vector<int> cand;
    for (int i = 0; i < n_last; ++i) {
        if (!valid[i]) continue;
        float xc[3];
        transform(Tcw, xw + 3 * i, xc);
        const float invzc = 1.0 / xc[2];                 // double division, then float (src/ORBmatcher.cc:1709)
        if (invzc < 0) continue;
        const float u = F.fx * xc[0] / xc[2] + F.cx, v = F.fy * xc[1] / xc[2] + F.cy;    // Pinhole::project (float)
        if (u < F.min_x || u > F.max_x) continue;
        if (v < F.min_y || v > F.max_y) continue;
        const int oct = last_octave[i];
        const float radius = th * F.scale_factors[oct];
        if (forward) features_in_area(F, g, u, v, radius, oct, -1, cand);
        else if (backward) features_in_area(F, g, u, v, radius, 0, oct, cand);
        else features_in_area(F, g, u, v, radius, oct - 1, oct + 1, cand);
        if (cand.empty()) continue;
        int best = 256, best_idx = -1;
        for (int i2 : cand) {
            if (state[i2] == 1) continue;
            if (F.uright[i2] > 0) {
                const float ur = u - F.bf * invzc;
                const float er = std::fabs(ur - F.uright[i2]);
                if (er > radius) continue;
            }
            const int d = descriptor_distance(mp_desc + 32 * i, F.desc + 32 * i2);
            if (d < best) { best = d; best_idx = i2; }
        }
        if (best <= TH_HIGH) {
            match[best_idx] = i;
            state[best_idx] = obs_pos[i] ? 1 : 2;
            ++nmatches;
            if (check_orientation) {
                float rot = last_angle[i] - F.keys_un[best_idx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rot_hist[bin].push_back(best_idx);
            }
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rot_hist, HISTO_LENGTH, i1, i2, i3);
        for (int b = 0; b < HISTO_LENGTH; ++b)
            if (b != i1 && b != i2 && b != i3)
                for (int idx : rot_hist[b]) { match[idx] = -2; --nmatches; }
    }
    return nmatches;
}

// Frame::isInFrustum for a list of map points (Nleft == -1).  Rcw row-major 3x3, tcw, Ow (= mOw), all float.
// Outputs per point: in_view, proj_x, proj_y, proj_xr, track_depth, scale_level, view_cos.
void orc_is_in_frustum(const orc_frame_view* fv, const float* Rcw, const float* tcw, const float* Ow, int n,
                       const float* xw, const float* normal, const float* mf_min_dist, const float* mf_max_dist,
                       float cos_limit, uint8_t* in_view, float* px, float* py, float* pxr, float* depth,
                       int* level, float* view_cos) {
    FrameView F = to_view(fv);
    for (int i = 0; i < n; ++i) {
        in_view[i] = 0; px[i] = -1; py[i] = -1; pxr[i] = 0; depth[i] = 0; level[i] = 0; view_cos[i] = 0;
        const float* P = xw + 3 * i;
        float Pc[3];
        for (int r = 0; r < 3; ++r) Pc[r] = sum3(Rcw[3 * r] * P[0], Rcw[3 * r + 1] * P[1], Rcw[3 * r + 2] * P[2]) + tcw[r];     // mRcw * P + mtcw
        const float pc_dist = std::sqrt(sum3(Pc[0] * Pc[0], Pc[1] * Pc[1], Pc[2] * Pc[2]));
        const float z = Pc[2];
        const float invz = 1.0f / z;
        if (z < 0.0f) continue;
        const float u = F.fx * Pc[0] / Pc[2] + F.cx, v = F.fy * Pc[1] / Pc[2] + F.cy;
        if (u < F.min_x || u > F.max_x) continue;
        if (v < F.min_y || v > F.max_y) continue;
        px[i] = u; py[i] = v;
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
        // GetMin/MaxDistanceInvariance: 0.8f*mfMinDistance, 1.2f*mfMaxDistance (src/MapPoint.cc:502-512)
        if (dist < 0.8f * mf_min_dist[i] || dist > 1.2f * mf_max_dist[i]) continue;
        const float* Pn = normal + 3 * i;
        const float vc = sum3(PO[0] * Pn[0], PO[1] * Pn[1], PO[2] * Pn[2]) / dist;
        if (vc < cos_limit) continue;
        const float mf_max = mf_max_dist[i];      // MapPoint::PredictScale uses mfMaxDistance itself (src/MapPoint.cc:531-545)
        const float ratio = mf_max / dist;
        int ns = (int)std::ceil(std::log(ratio) / F.log_scale_factor);
        if (ns < 0) ns = 0; else if (ns >= F.n_levels) ns = F.n_levels - 1;
        in_view[i] = 1; pxr[i] = u - F.bf * invz; depth[i] = pc_dist; level[i] = ns; view_cos[i] = vc;
    }
}

// SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints), Nleft == -1.  Per map point: in_view
// (mbTrackInView && !isBad()), proj_x/y/xr, track_depth, level (mnTrackScaleLevel), view_cos, desc, obs_pos.
int orc_search_by_projection_local(const orc_frame_view* fv, int n, const uint8_t* in_view, const float* px,
                                   const float* py, const float* pxr, const float* track_depth, const int* level,
                                   const float* view_cos, const uint8_t* mp_desc, const uint8_t* obs_pos, float th,
                                   float nn_ratio, int far_points, float th_far, const uint8_t* cur_state, int* match) {
    FrameView F = to_view(fv);
    Grid g; g.build(F);
    int nmatches = 0;
    const bool use_factor = th != 1.0;
    std::vector<uint8_t> state(cur_state, cur_state + F.n);
    for (int i = 0; i < F.n; ++i) match[i] = -1;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        if (!in_view[i]) continue;
        if (far_points && track_depth[i] > th_far) continue;
        const int pl = level[i];
        float r = (view_cos[i] > 0.998) ? 2.5f : 4.0f;          // RadiusByViewingCos, :215-221 (double literal compare)
        if (use_factor) r *= th;
        features_in_area(F, g, px[i], py[i], r * F.scale_factors[pl], pl - 1, pl, cand);
        if (cand.empty()) continue;
        int best = 256, best_level = -1, best2 = 256, best_level2 = -1, best_idx = -1;
        for (int idx : cand) {
            if (state[idx] == 1) continue;
            if (F.uright[idx] > 0) {
                const float er = std::fabs(pxr[i] - F.uright[idx]);
                if (er > r * F.scale_factors[pl]) continue;
            }
            const int d = descriptor_distance(mp_desc + 32 * i, F.desc + 32 * idx);
            if (d < best) { best2 = best; best = d; best_level2 = best_level; best_level = F.keys_un[idx].octave; best_idx = idx; }
            else if (d < best2) { best_level2 = F.keys_un[idx].octave; best2 = d; }
        }
        if (best <= TH_HIGH) {
            if (best_level == best_level2 && best > nn_ratio * best2) continue;
            if (best_level != best_level2 || best <= nn_ratio * best2) {
                match[best_idx] = i;
                state[best_idx] = obs_pos[i] ? 1 : 2;
                ++nmatches;
            }
        }
    }
    return nmatches;
}


// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), Nleft == -1 path (src/ORBmatcher.cc:223-425).
// The two DBoW2::FeatureVector maps are given as CSR: ascending node ids, per node the feature indices in vector order.
// kf_valid[i] = (vpMapPointsKF[i] != NULL && !isBad()).  match[idxF] = KF feature index (the caller maps it to the
// MapPoint*) or -1.  Returns nmatches.
int orc_search_by_bow(int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                      int n_nodes_kf, const uint32_t* kf_node_ids, const int* kf_node_start, const int* kf_node_feat,
                      int n_f, const uint8_t* f_desc, const float* f_angle,
                      int n_nodes_f, const uint32_t* f_node_ids, const int* f_node_start, const int* f_node_feat,
                      float nn_ratio, int check_orientation, int* match) {
    constexpr int TH_LOW = 50;
    (void)n_kf;
    for (int i = 0; i < n_f; ++i) match[i] = -1;
    int nmatches = 0;
    std::vector<int> rot_hist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < n_nodes_kf && b < n_nodes_f) {
        if (kf_node_ids[a] == f_node_ids[b]) {
            for (int ik = kf_node_start[a]; ik < kf_node_start[a + 1]; ++ik) {
                const int ikf = kf_node_feat[ik];
                if (!kf_valid[ikf]) continue;
                int best1 = 256, best_idx = -1, best2 = 256;
                for (int jf = f_node_start[b]; jf < f_node_start[b + 1]; ++jf) {
                    const int idf = f_node_feat[jf];
                    if (match[idf] >= 0) continue;
                    const int d = descriptor_distance(kf_desc + 32 * ikf, f_desc + 32 * idf);
                    if (d < best1) { best2 = best1; best1 = d; best_idx = idf; }
                    else if (d < best2) best2 = d;
                }
                if (best1 <= TH_LOW) {
                    if (static_cast<float>(best1) < nn_ratio * static_cast<float>(best2)) {
                        match[best_idx] = ikf;
                        if (check_orientation) {
                            float rot = kf_angle[ikf] - f_angle[best_idx];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)std::round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rot_hist[bin].push_back(best_idx);
                        }
                        ++nmatches;
                    }
                }
            }
            ++a; ++b;
        } else if (kf_node_ids[a] < f_node_ids[b]) {
            while (a < n_nodes_kf && kf_node_ids[a] < f_node_ids[b]) ++a;      // lower_bound
        } else {
            while (b < n_nodes_f && f_node_ids[b] < kf_node_ids[a]) ++b;
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rot_hist, HISTO_LENGTH, i1, i2, i3);
        for (int k = 0; k < HISTO_LENGTH; ++k) {
            if (k == i1 || k == i2 || k == i3) continue;
            for (int idx : rot_hist[k]) { match[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>&, th, bRight = false) (src/ORBmatcher.cc:1148-1330), the search part:
// per map point (valid[i] = pMP && !isBad() && !IsInKeyFrame(pKF)) the best key-frame feature in the radius (index, Hamming
// distance) or (-1, 256); the Replace / AddObservation bookkeeping that follows (:1306-1325) consumes best_dist <= TH_LOW on the
// host.  The key frame's members the search reads are those of the frame view (KeyFrame::GetFeaturesInArea = Frame's,
// src/KeyFrame.cc:704-748; IsInImage :750-753; mvInvLevelSigma2 = 1 / scale^2).
void orc_fuse_search(const orc_frame_view* kf, const float Tcw_[7], const float Ow[3], int n, const uint8_t* valid, const float* xw,
                     const float* normal, const float* mf_min_dist, const float* mf_max_dist, const uint8_t* mp_desc, float th,
                     int* best_idx, int* best_dist) {
    FrameView F = to_view(kf);
    Grid g; g.build(F);
    const Pose Tcw = {Tcw_[0], Tcw_[1], Tcw_[2], Tcw_[3], Tcw_[4], Tcw_[5], Tcw_[6]};
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (!valid[i]) continue;
        const float* P = xw + 3 * i;
        float Pc[3];
        transform(Tcw, P, Pc);
        if (Pc[2] < 0.0f) continue;
        const float invz = 1 / Pc[2];
        const float u = F.fx * Pc[0] / Pc[2] + F.cx, v = F.fy * Pc[1] / Pc[2] + F.cy;      // Pinhole::project(Vector3f)
        if (!(u >= F.min_x && u < F.max_x && v >= F.min_y && v < F.max_y)) continue;        // KeyFrame::IsInImage
        const float ur = u - F.bf * invz;
        const float max_d = 1.2f * mf_max_dist[i], min_d = 0.8f * mf_min_dist[i];           // Get{Max,Min}DistanceInvariance
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist3D = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
        if (dist3D < min_d || dist3D > max_d) continue;
        const float* Pn = normal + 3 * i;
        if (sum3(PO[0] * Pn[0], PO[1] * Pn[1], PO[2] * Pn[2]) < 0.5 * dist3D) continue;
        const float ratio = mf_max_dist[i] / dist3D;                                        // MapPoint::PredictScale(dist, KeyFrame*)
        int level = (int)std::ceil(std::log(ratio) / F.log_scale_factor);
        if (level < 0) level = 0; else if (level >= F.n_levels) level = F.n_levels - 1;
        const float radius = th * F.scale_factors[level];
        features_in_area(F, g, u, v, radius, -1, -1, cand);
        int bd = 256, bi = -1;
        for (int idx : cand) {
            const KeyPoint& kp = F.keys_un[idx];
            const int kl = kp.octave;
            if (kl < level - 1 || kl > level) continue;
            const float inv_sigma2 = 1.0f / (F.scale_factors[kl] * F.scale_factors[kl]);
            if (F.uright[idx] >= 0) {
                const float ex = u - kp.x, ey = v - kp.y, er = ur - F.uright[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_sigma2 > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_sigma2 > 5.99) continue;
            }
            const int d = descriptor_distance(mp_desc + 32 * (size_t)i, F.desc + 32 * (size_t)idx);
            if (d < bd) { bd = d; bi = idx; }
        }
        best_idx[i] = bi; best_dist[i] = bd;
    }
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame*, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
// (src/ORBmatcher.cc:1889-2010), the relocalisation refinement.  Per key-frame map point i: valid[i] = (pMP && !isBad() &&
// !sAlreadyFound.count(pMP)), xw, descriptor, the key frame's keypoint angle, mfMinDistance / mfMaxDistance.
// cur_occupied[i2] != 0 <=> CurrentFrame.mvpMapPoints[i2] != NULL.  match[i2]: >= 0 index i, -1 untouched, -2 cleared.
int orc_search_by_projection_reloc(const orc_frame_view* cur, const float cur_pose[7], int n, const uint8_t* valid,
                                   const float* xw, const uint8_t* mp_desc, const float* kf_angle, const float* mf_min_dist,
                                   const float* mf_max_dist, float th, int orb_dist, int check_orientation,
                                   const uint8_t* cur_occupied, int* match) {
    FrameView F = to_view(cur);
    Grid g; g.build(F);
    Pose Tcw = {cur_pose[0], cur_pose[1], cur_pose[2], cur_pose[3], cur_pose[4], cur_pose[5], cur_pose[6]};
    float Ow[3];
    inverse_translation(Tcw, Ow);
    int nmatches = 0;
    std::vector<int> rot_hist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<uint8_t> occ(cur_occupied, cur_occupied + F.n);
    for (int i = 0; i < F.n; ++i) match[i] = -1;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        float xc[3];
        transform(Tcw, xw + 3 * i, xc);
        const float u = F.fx * xc[0] / xc[2] + F.cx, v = F.fy * xc[1] / xc[2] + F.cy;
        if (u < F.min_x || u > F.max_x) continue;
        if (v < F.min_y || v > F.max_y) continue;
        const float PO[3] = {xw[3 * i] - Ow[0], xw[3 * i + 1] - Ow[1], xw[3 * i + 2] - Ow[2]};
        const float dist3D = std::sqrt(sum3(PO[0] * PO[0], PO[1] * PO[1], PO[2] * PO[2]));
        if (dist3D < 0.8f * mf_min_dist[i] || dist3D > 1.2f * mf_max_dist[i]) continue;
        const float ratio = mf_max_dist[i] / dist3D;
        int pl = (int)std::ceil(std::log(ratio) / F.log_scale_factor);
        if (pl < 0) pl = 0; else if (pl >= F.n_levels) pl = F.n_levels - 1;
        const float radius = th * F.scale_factors[pl];
        features_in_area(F, g, u, v, radius, pl - 1, pl + 1, cand);
        if (cand.empty()) continue;
        int best = 256, best_idx = -1;
        for (int i2 : cand) {
            if (occ[i2]) continue;
            const int d = descriptor_distance(mp_desc + 32 * i, F.desc + 32 * i2);
            if (d < best) { best = d; best_idx = i2; }
        }
        if (best <= orb_dist) {
            match[best_idx] = i; occ[best_idx] = 1; ++nmatches;
            if (check_orientation) {
                float rot = kf_angle[i] - F.keys_un[best_idx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rot_hist[bin].push_back(best_idx);
            }
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rot_hist, HISTO_LENGTH, i1, i2, i3);
        for (int b = 0; b < HISTO_LENGTH; ++b)
            if (b != i1 && b != i2 && b != i3)
                for (int idx : rot_hist[b]) { match[idx] = -2; --nmatches; }
    }
    return nmatches;
}

}  // extern "C"
