// Frame::ComputeStereoMatches (src/Frame.cc:901-1071) for two frame slots (left, right) of one batched extraction.
// One warp per left keypoint: candidate test against every right keypoint (row band of +-2*scale around the right
// keypoint, octave +-1, disparity range), Hamming best (first strictly smaller in right-keypoint order), then the
// 11x11 SAD over 11 shifts on the resident pyramid level, parabola fit.  A second one-CTA kernel does the median-based
// rejection (rank selection instead of the reference's sort: same element at index size/2).
#include <climits>

#include "rgbl_device.cuh"
#include "rgbl_kernels.h"

namespace rgbl {

__global__ void __launch_bounds__(256) stereo_match_kernel(const uint8_t* __restrict__ pyr, size_t frame_stride, int slot_l, int slot_r,
                                                           const LevelGeom* __restrict__ levels, StereoFrameDev L, StereoFrameDev R,
                                                           float mb, float mbf, int n_rows, float* __restrict__ depth,
                                                           float* __restrict__ uright, int* __restrict__ sad_out) {
    const int il = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n_l = *L.n, n_r = *R.n;
    if (il >= n_l) return;
    const rgbl_keypoint kp = L.keys[il];
    const int lvl = kp.octave;
    const float vL = kp.y, uL = kp.x;
    const int row = (int)vL;
    const float maxD = __fdiv_rn(mbf, mb), minU = __fsub_rn(uL, maxD), maxU = uL;       // minD = 0
    float out_d = -1.f, out_u = -1.f; int out_sad = -1;
    const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(L.desc + (size_t)il * 32)), d1 = __ldg(reinterpret_cast<const uint4*>(L.desc + (size_t)il * 32) + 1);
    unsigned best = 0xffffffffu;                    // dist << 16 | iR  (first strictly smaller in iR order = min of this key)
    if (!(maxU < 0) && row >= 0 && row < n_rows) {
        for (int ir = lane; ir < n_r; ir += 32) {
            const rgbl_keypoint kr = R.keys[ir];
            const float r = __fmul_rn(2.0f, L.scale[kr.octave]);
            const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));
            if (row < minr || row > maxr) continue;
            if (kr.octave < lvl - 1 || kr.octave > lvl + 1) continue;
            if (!(kr.x >= minU && kr.x <= maxU)) continue;
            const uint8_t* b = R.desc + (size_t)ir * 32;
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b)), b1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
            const int d = __popc(d0.x ^ b0.x) + __popc(d0.y ^ b0.y) + __popc(d0.z ^ b0.z) + __popc(d0.w ^ b0.w) +
                          __popc(d1.x ^ b1.x) + __popc(d1.y ^ b1.y) + __popc(d1.z ^ b1.z) + __popc(d1.w ^ b1.w);
            if (d < 100) best = min(best, ((unsigned)d << 16) | (unsigned)ir);      // bestDist starts at TH_HIGH, strict <
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (best != 0xffffffffu && (int)(best >> 16) < 75) {                            // thOrbDist = (TH_HIGH + TH_LOW) / 2
        const int ir = (int)(best & 0xffffu);
        const float uR0 = R.keys[ir].x;
        const float sf = L.inv_scale[lvl];
        const float su = roundf(__fmul_rn(kp.x, sf)), sv = roundf(__fmul_rn(kp.y, sf)), sur0 = roundf(__fmul_rn(uR0, sf));
        const LevelGeom lg = levels[lvl];
        const float iniu = sur0, endu = __fadd_rn(sur0, 11.0f);                       // scaleduR0 + L - w, scaleduR0 + L + w + 1
        if (!(iniu < 0 || endu >= (float)lg.w)) {
            const uint8_t* IL = pyr + (size_t)slot_l * frame_stride + lg.off;
            const uint8_t* IR = pyr + (size_t)slot_r * frame_stride + lg.off;
            const int cu = (int)su, cv = (int)sv, cr = (int)sur0;
            int sad = INT_MAX;
            if (lane < 11) {
                const int inc = lane - 5;
                sad = 0;
                for (int dy = -5; dy <= 5; ++dy) {
                    const uint8_t* a = IL + (size_t)(cv + dy) * lg.pitch + cu - 5;
                    const uint8_t* b = IR + (size_t)(cv + dy) * lg.pitch + cr + inc - 5;
#pragma unroll
                    for (int dx = 0; dx < 11; ++dx) sad += abs((int)__ldg(a + dx) - (int)__ldg(b + dx));
                }
            }
            // best shift: first strictly smaller in inc order = min of (sad << 8 | lane)
            unsigned long long key = (lane < 11) ? (((unsigned long long)(unsigned)sad << 8) | (unsigned)lane) : ~0ull;
            unsigned lo = (unsigned)(key & 0xffffffffu), hi = (unsigned)(key >> 32);
            // 40-bit key min via two-step reduce (hi then lo among the hi-minimal lanes)
            const unsigned hmin = __reduce_min_sync(0xffffffffu, hi);
            const unsigned lmin = __reduce_min_sync(0xffffffffu, (hi == hmin) ? lo : 0xffffffffu);
            const int best_lane = (int)(lmin & 0xffu);
            const int best_inc = best_lane - 5;
            const int best_sad = (int)((((unsigned long long)hmin << 32) | lmin) >> 8);
            if (best_inc != -5 && best_inc != 5) {
                const float dist1 = (float)__shfl_sync(0xffffffffu, sad, best_lane - 1);
                const float dist2 = (float)__shfl_sync(0xffffffffu, sad, best_lane);
                const float dist3 = (float)__shfl_sync(0xffffffffu, sad, best_lane + 1);
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
                const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), den);
                if (!(deltaR < -1.f || deltaR > 1.f)) {
                    float best_ur = __fmul_rn(L.scale[lvl], __fadd_rn(__fadd_rn(sur0, (float)best_inc), deltaR));
                    float disparity = __fsub_rn(uL, best_ur);
                    if (disparity >= 0.f && disparity < maxD) {
                        if (disparity <= 0.f) { disparity = (float)0.01; best_ur = (float)((double)uL - 0.01); }
                        out_d = __fdiv_rn(mbf, disparity); out_u = best_ur; out_sad = best_sad;
                    }
                }
            } else {
                // keep the shuffles convergent for the whole warp
                (void)__shfl_sync(0xffffffffu, sad, 0); (void)__shfl_sync(0xffffffffu, sad, 0); (void)__shfl_sync(0xffffffffu, sad, 0);
            }
        }
    }
    if (lane == 0) { depth[il] = out_d; uright[il] = out_u; sad_out[il] = out_sad; }
}

// median-based rejection (src/Frame.cc:1057-1070): element at index size/2 of the (dist, iL)-sorted list
__global__ void __launch_bounds__(1024) stereo_median_kernel(const int* __restrict__ n_ptr, const int* __restrict__ sad,
                                                             float* __restrict__ depth, float* __restrict__ uright) {
    __shared__ int s_count, s_median;
    const int tid = threadIdx.x, n = *n_ptr;
    if (tid == 0) { s_count = 0; s_median = -1; }
    __syncthreads();
    int local = 0;
    for (int i = tid; i < n; i += 1024) local += (sad[i] >= 0);
    if (local) atomicAdd(&s_count, local);
    __syncthreads();
    const int m = s_count;
    if (m == 0) return;
    const int target = m / 2;
    for (int i = tid; i < n; i += 1024) {
        const int si = sad[i];
        if (si < 0) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) { const int sj = sad[j]; if (sj >= 0 && (sj < si || (sj == si && j < i))) ++rank; }
        if (rank == target) s_median = si;
    }
    __syncthreads();
    const float th = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_median);
    for (int i = tid; i < n; i += 1024)
        if (sad[i] >= 0 && !((float)sad[i] < th)) { depth[i] = -1.f; uright[i] = -1.f; }
}

void launch_stereo_matches(cudaStream_t st, const uint8_t* pyr, size_t frame_stride, int slot_l, int slot_r, const LevelGeom* d_levels,
                           const StereoFrameDev& L, const StereoFrameDev& R, float mb, float mbf, int n_rows, int cap, float* depth,
                           float* uright, int* sad) {
    stereo_match_kernel<<<(cap + 7) / 8, 256, 0, st>>>(pyr, frame_stride, slot_l, slot_r, d_levels, L, R, mb, mbf, n_rows, depth, uright, sad);
    stereo_median_kernel<<<1, 1024, 0, st>>>(L.n, sad, depth, uright);
}

}  // namespace rgbl
