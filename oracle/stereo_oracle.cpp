// ORACLE (test infrastructure, NOT product code): CPU restatement of Frame::ComputeStereoMatches
// (/root/reference/src/Frame.cc:901-1071): row-table candidates, Hamming best match, 11x11 SAD refinement sliding +-5 px on
// the pyramid level of the left keypoint (cv::norm(IL, IR, NORM_L1) on CV_8U = integer sum of absolute differences),
// parabola sub-pixel fit, median-based outlier removal.  Level images are the un-padded planes (all windows stay inside
// them, see the bounds checks at :1006-1009).  Only tests/, smoke() and bench.py's CPU legs may load this library.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {
struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };

int hamming(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y; std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        d += __builtin_popcount(x ^ y);
    }
    return d;
}
}  // namespace

extern "C" {

// levels_l / levels_r: pointers to the un-padded level images (row stride = width), level sizes in lw/lh.
void orc_stereo_matches(int n_l, const void* kps_l_, const uint8_t* desc_l, int n_r, const void* kps_r_, const uint8_t* desc_r,
                        int n_levels, const float* scale, const float* inv_scale, const uint8_t* const* levels_l,
                        const uint8_t* const* levels_r, const int* lw, const int* lh, float mb, float mbf, float* depth, float* uright) {
    const KeyPoint* kl = (const KeyPoint*)kps_l_; const KeyPoint* kr = (const KeyPoint*)kps_r_;
    (void)n_levels;
    for (int i = 0; i < n_l; ++i) { depth[i] = -1.f; uright[i] = -1.f; }
    const int th_orb = (100 + 50) / 2;
    const int n_rows = lh[0];
    std::vector<std::vector<int>> rows(n_rows);
    for (int ir = 0; ir < n_r; ++ir) {
        const float r = 2.0f * scale[kr[ir].octave];
        const int maxr = (int)std::ceil(kr[ir].y + r), minr = (int)std::floor(kr[ir].y - r);
        for (int y = minr; y <= maxr; ++y) if (y >= 0 && y < n_rows) rows[y].push_back(ir);
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> dist_idx;
    for (int il = 0; il < n_l; ++il) {
        const KeyPoint& kp = kl[il];
        const int lvl = kp.octave;
        const float vL = kp.y, uL = kp.x;
        const std::vector<int>& cand = rows[(int)vL];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int best = 100; int best_r = 0;
        for (int ir : cand) {
            if (kr[ir].octave < lvl - 1 || kr[ir].octave > lvl + 1) continue;
            const float uR = kr[ir].x;
            if (uR >= minU && uR <= maxU) {
                const int d = hamming(desc_l + 32 * il, desc_r + 32 * ir);
                if (d < best) { best = d; best_r = ir; }
            }
        }
        if (best < th_orb) {
            const float uR0 = kr[best_r].x;
            const float sf = inv_scale[lvl];
            const float su = std::round(kp.x * sf), sv = std::round(kp.y * sf), sur0 = std::round(uR0 * sf);
            const int w = 5, L = 5;
            const float iniu = sur0 + L - w, endu = sur0 + L + w + 1;
            if (iniu < 0 || endu >= lw[lvl]) continue;
            const uint8_t* IL = levels_l[lvl]; const uint8_t* IR = levels_r[lvl];
            const int W = lw[lvl];
            int best_sad = INT_MAX, best_inc = 0;
            float dists[2 * 5 + 1];
            for (int inc = -L; inc <= L; ++inc) {
                int sad = 0;
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        const int a = IL[(size_t)((int)sv + dy) * W + (int)su + dx];
                        const int b = IR[(size_t)((int)sv + dy) * W + (int)sur0 + inc + dx];
                        sad += std::abs(a - b);
                    }
                const float dist = (float)sad;
                if (dist < best_sad) { best_sad = (int)dist; best_inc = inc; }
                dists[L + inc] = dist;
            }
            if (best_inc == -L || best_inc == L) continue;
            const float d1 = dists[L + best_inc - 1], d2 = dists[L + best_inc], d3 = dists[L + best_inc + 1];
            const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
            if (deltaR < -1 || deltaR > 1) continue;
            float best_ur = scale[lvl] * ((float)sur0 + (float)best_inc + deltaR);
            float disparity = (uL - best_ur);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; best_ur = uL - 0.01; }
                depth[il] = mbf / disparity;
                uright[il] = best_ur;
                dist_idx.push_back(std::make_pair(best_sad, il));
            }
        }
    }
    if (dist_idx.empty()) return;
    std::sort(dist_idx.begin(), dist_idx.end());
    const float median = dist_idx[dist_idx.size() / 2].first;
    const float th = 1.5f * 1.4f * median;
    for (int i = (int)dist_idx.size() - 1; i >= 0; --i) {
        if (dist_idx[i].first < th) break;
        uright[dist_idx[i].second] = -1; depth[dist_idx[i].second] = -1;
    }
}

}  // extern "C"
