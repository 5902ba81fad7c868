// ORACLE (test infrastructure, NOT product code): CPU restatement of two LocalMapping-thread kernels of the reference
//   MapPoint::ComputeDistinctiveDescriptors            src/MapPoint.cc:329-403
//   ORBmatcher::SearchForTriangulation                 src/ORBmatcher.cc:907-1146 (Nleft == -1, no second camera), with
//     Pinhole::epipolarConstrain                       src/CameraModels/Pinhole.cpp:107-129 (F12 is per key-frame pair: the caller passes it)
//     ORBmatcher::ComputeThreeMaxima                   src/ORBmatcher.cc:2012-2053
// PARITY UNPINNED (no reference vectors; restated from the sources).  Only tests/ and bench.py's CPU legs load this library.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };
constexpr int TH_LOW = 50, HISTO_LENGTH = 30;

inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 4; ++i) { uint64_t x, y; std::memcpy(&x, a + 8 * i, 8); std::memcpy(&y, b + 8 * i, 8); d += __builtin_popcountll(x ^ y); }
    return d;
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

// observations of point p: descriptors desc[obs_start[p] .. obs_start[p+1]).  best[p] = index (within the point's list) of the
// descriptor with the least median distance to the others (first minimum), -1 for a point without observations.
void orc_distinctive_descriptors(int n_points, const int32_t* obs_start, const uint8_t* desc, int32_t* best) {
    for (int p = 0; p < n_points; ++p) {
        const int b = obs_start[p], N = obs_start[p + 1] - b;
        best[p] = -1;
        if (N <= 0) continue;
        std::vector<float> D((size_t)N * N);
        for (int i = 0; i < N; ++i) {
            D[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; ++j) {
                const int d = descriptor_distance(desc + 32 * (size_t)(b + i), desc + 32 * (size_t)(b + j));
                D[(size_t)i * N + j] = (float)d; D[(size_t)j * N + i] = (float)d;
            }
        }
        int best_median = INT_MAX, best_idx = 0;
        for (int i = 0; i < N; ++i) {
            std::vector<int> v(D.begin() + (size_t)i * N, D.begin() + (size_t)(i + 1) * N);
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < best_median) { best_median = median; best_idx = i; }
        }
        best[p] = best_idx;
    }
}

// match12[idx1] = idx2 or -1; returns nmatches.  has_mp*: the feature already has a map point; uright*: mvuRight (>= 0 <=> stereo).
// Feature vectors as CSR (ascending node ids, start[n + 1], features in vector order).  F12 row-major, ep = epipole in image 2,
// scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2.
int orc_search_for_triangulation(int n1, const uint8_t* desc1, const void* keys1_, const uint8_t* has_mp1, const float* uright1,
                                 int nn1, const uint32_t* node_ids1, const int* node_start1, const int* node_feat1,
                                 int n2, const uint8_t* desc2, const void* keys2_, const uint8_t* has_mp2, const float* uright2,
                                 int nn2, const uint32_t* node_ids2, const int* node_start2, const int* node_feat2,
                                 const float* F12, const float* ep, const float* scale_factors2, const float* level_sigma2_2,
                                 int only_stereo, int coarse, int check_orientation, int32_t* match12) {
    const KeyPoint* keys1 = (const KeyPoint*)keys1_; const KeyPoint* keys2 = (const KeyPoint*)keys2_;
    (void)n2;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    int nmatches = 0;
    std::vector<int> rot_hist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (node_ids1[a] == node_ids2[b]) {
            for (int i1 = node_start1[a]; i1 < node_start1[a + 1]; ++i1) {
                const int idx1 = node_feat1[i1];
                if (has_mp1[idx1]) continue;
                const bool stereo1 = uright1[idx1] >= 0;
                if (only_stereo && !stereo1) continue;
                const KeyPoint& kp1 = keys1[idx1];
                int best_dist = TH_LOW, best_idx2 = -1;
                for (int i2 = node_start2[b]; i2 < node_start2[b + 1]; ++i2) {
                    const int idx2 = node_feat2[i2];
                    if (has_mp2[idx2]) continue;                       // vbMatched2 is never set in this version of the reference
                    const bool stereo2 = uright2[idx2] >= 0;
                    if (only_stereo && !stereo2) continue;
                    const int dist = descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > best_dist) continue;
                    const KeyPoint& kp2 = keys2[idx2];
                    if (!stereo1 && !stereo2) {
                        const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[kp2.octave]) continue;
                    }
                    bool ok = coarse != 0;
                    if (!ok) {                                          // Pinhole::epipolarConstrain
                        const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                        const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                        const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                        const float num = la * kp2.x + lb * kp2.y + lc;
                        const float den = la * la + lb * lb;
                        if (den != 0) { const float dsqr = num * num / den; ok = dsqr < 3.84 * level_sigma2_2[kp2.octave]; }
                    }
                    if (ok) { best_idx2 = idx2; best_dist = dist; }
                }
                if (best_idx2 >= 0) {
                    match12[idx1] = best_idx2;
                    ++nmatches;
                    if (check_orientation) {
                        float rot = kp1.angle - keys2[best_idx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rot_hist[bin].push_back(idx1);
                    }
                }
            }
            ++a; ++b;
        } else if (node_ids1[a] < node_ids2[b]) {
            while (a < nn1 && node_ids1[a] < node_ids2[b]) ++a;
        } else {
            while (b < nn2 && node_ids2[b] < node_ids1[a]) ++b;
        }
    }
    if (check_orientation) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rot_hist, HISTO_LENGTH, i1, i2, i3);
        for (int k = 0; k < HISTO_LENGTH; ++k) {
            if (k == i1 || k == i2 || k == i3) continue;
            for (int idx : rot_hist[k]) { match12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

}  // extern "C"
