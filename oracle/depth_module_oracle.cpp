// ORACLE (test infrastructure, NOT product code): CPU restatement of the reference's DepthModule
// hot path (/root/reference/src/DepthModule.cc:50-274) with the OpenCV primitives it calls
// (gemm via Mat*Mat, mul, 1/Mat, threshold, dilate, filter2D, distanceTransform, minMaxLoc)
// restated in closed form and pinned against cv2 4.13.0 in tests/test_oracle_vs_cv2.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs load it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };

// DepthModule::ProjectPointcloudToImage, DepthModule.cc:106-139.
// pts: 4 x n planar rows (x, y, z, 1) as built by rgbl_kitti.cc:168-177.  P: 3x4 row-major float.
// cv::gemm for CV_32F accumulates each dot product in double and rounds once; the normalisation is
// row.mul(1/row2): a float reciprocal followed by a float multiply (two roundings).
void project(const float* pts, int n, const float P[12], int W, int H, float min_d, float max_d,
             float* raw /* H x W, zero-filled here */) {
    std::fill(raw, raw + (size_t)W * H, 0.f);
    const float* X = pts; const float* Y = pts + n; const float* Z = pts + 2 * (size_t)n; const float* O = pts + 3 * (size_t)n;
    for (int i = 0; i < n; ++i) {
        float q[3];
        for (int r = 0; r < 3; ++r) {
            double acc = (double)P[4 * r] * X[i] + (double)P[4 * r + 1] * Y[i] + (double)P[4 * r + 2] * Z[i] + (double)P[4 * r + 3] * O[i];
            q[r] = (float)acc;
        }
        float inv = 1.0f / q[2];
        float u = q[0] * inv, v = q[1] * inv, d = q[2];
        if (u > 0 && v > 0 && u < W && v < H)
            if (d > min_d && d < max_d) raw[(size_t)(int)v * W + (int)u] = d;   // later points overwrite
    }
}

// cv::dilate on CV_32F with an arbitrary 0/1 structuring element, anchor at the centre, default
// border (out-of-image taps ignored).
void dilate_f32(const float* src, int W, int H, const uint8_t* mask, int ku, int kv, float* dst) {
    const int ax = ku / 2, ay = kv / 2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float m = -std::numeric_limits<float>::max();
            for (int j = 0; j < kv; ++j) {
                int yy = y + j - ay;
                if (yy < 0 || yy >= H) continue;
                for (int i = 0; i < ku; ++i) {
                    if (!mask[j * ku + i]) continue;
                    int xx = x + i - ax;
                    if (xx < 0 || xx >= W) continue;
                    m = std::max(m, src[(size_t)yy * W + xx]);
                }
            }
            dst[(size_t)y * W + x] = m;
        }
}

// DepthModule::Upsample_InverseDilation, DepthModule.cc:230-274 (the five Mat operations in order).
void inverse_dilation(const float* raw, int W, int H, float max_dist, float scale, const uint8_t* mask,
                      int ku, int kv, float* out) {
    const float M = max_dist * scale;
    const float thr = M - 1;
    std::vector<float> t((size_t)W * H), d((size_t)W * H);
    for (size_t i = 0; i < t.size(); ++i) { float v = M - raw[i]; t[i] = (v > thr) ? 0.f : v; }
    dilate_f32(t.data(), W, H, mask, ku, kv, d.data());
    for (size_t i = 0; i < t.size(); ++i) { float v = M - d[i]; out[i] = (v > thr) ? 0.f : v; }
}

// DepthModule::GetFeatureDepthFromDepthMap, DepthModule.cc:82-104.
void gather(const float* map, int W, const KeyPoint* k, const KeyPoint* ku, int n, float bf, float* depth, float* uright) {
    for (int i = 0; i < n; ++i) {
        depth[i] = -1.f; uright[i] = -1.f;
        float d = map[(size_t)(int)k[i].y * W + (int)k[i].x];    // Mat::at<float>(float v, float u): truncation
        if (d > 0) { depth[i] = d; uright[i] = ku[i].x - bf / d; }
    }
}

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

// DepthModule::Upsample_AverageFiltering, DepthModule.cc:200-228.
// cv::filter2D on CV_32F with the k x k kernel of (float)1/k^2 and BORDER_REFLECT_101: OpenCV's direct engine accumulates
// acc = fma(kernel_tap, src, acc) over the taps in row-major order starting from delta = 0 (its AVX2/FMA dispatch; probed
// against cv2 4.13.0 on an AVX2 host: 0 mismatches with fused multiply-add, 39 560 without).  The count image sums exact
// 0/1 values.  Processed = Filtered .* (k^2 ./ Count); empty windows give 0 * inf = NaN, which fails d > 0 in the gather.
void average_filter(const float* raw, int W, int H, int k, float* out) {
    const int a = k / 2;
    const float kv = 1.0f / (float)(k * k);   // Mat::ones(CV_32F) / k^2
    const float k2 = (float)(k * k);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float s = 0.f, c = 0.f;
            for (int j = 0; j < k; ++j) {
                const int yy = reflect101(y + j - a, H);
                for (int i = 0; i < k; ++i) {
                    const int xx = reflect101(x + i - a, W);
                    const float v = raw[(size_t)yy * W + xx];
                    s = std::fmaf(kv, v, s);
                    c += (v > 0.f) ? 1.f : 0.f;
                }
            }
            out[(size_t)y * W + x] = s * (k2 / c);
        }
}

// cv::distanceTransform(DIST_L2, DIST_MASK_5) value at one pixel: OpenCV runs the 5x5 chamfer in 16.16 fixed point
// (a = 1, b = 1.4, c = 2.1969 -> 65536, 91750, 143976), so the two-pass result equals the closed-form chamfer cost to the
// nearest source pixel; returned as float(fixed * 2^-16) like the library.  Sources = pixels whose rounded depth is 0.
inline unsigned chamfer5_fixed(int dx, int dy) {
    const unsigned A = 65536u, B = 91750u, Cc = 143976u;
    dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
    if (dx < dy) { const int t = dx; dx = dy; dy = t; }
    if (dx >= 2 * dy) return (unsigned)(dx - 2 * dy) * A + (unsigned)dy * Cc;
    return (unsigned)(2 * dy - dx) * B + (unsigned)(dx - dy) * Cc;
}

// DepthModule::Upsample_NearestNeighbor_Pixel, DepthModule.cc:145-198 (R = SearchRadius).
void nearest_neighbor_pixel(const float* raw, int W, int H, const KeyPoint* k, const KeyPoint* ku, int n, float bf, float Rf,
                            float* depth, float* uright) {
    const int R = (int)Rf;
    for (int i = 0; i < n; ++i) {
        depth[i] = -1.f; uright[i] = -1.f;
        const int u = (int)k[i].x, v = (int)k[i].y;
        // distance (fixed point) to the nearest pixel that holds a depth: only values below R+1 matter
        unsigned best = 0xffffffffu;
        const int win = R + 1;
        for (int yy = std::max(0, v - win); yy <= std::min(H - 1, v + win); ++yy)
            for (int xx = std::max(0, u - win); xx <= std::min(W - 1, u + win); ++xx) {
                const float d = raw[(size_t)yy * W + xx];
                if ((int)lrintf(d) > 0 || d >= 255.5f) best = std::min(best, chamfer5_fixed(xx - u, yy - v));   // convertTo(CV_8U) != 0
            }
        if (best == 0xffffffffu) continue;            // farther than R+1: searchradius >= R
        const float dist = (float)(best * (1.0 / 65536.0));
        int sr = (int)dist;
        float d = 0.f;
        if (sr >= 0 && sr < Rf) {
            ++sr;
            const int bx = (int)(k[i].x + Rf - (float)sr) - R, by = (int)(k[i].y + Rf - (float)sr) - R;   // Rect in padded coords -> image
            for (int yy = by; yy < by + 2 * sr; ++yy)
                for (int xx = bx; xx < bx + 2 * sr; ++xx) {
                    const float val = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? raw[(size_t)yy * W + xx] : 0.f;
                    d = std::max(d, val);
                }
        }
        if (d > 0) { depth[i] = d; uright[i] = ku[i].x - bf / d; }
    }
}

}  // namespace

extern "C" {

void orc_depth_project(const float* pts4xn, int n, const float* P, int W, int H, float min_d, float max_d, float* raw) {
    project(pts4xn, n, P, W, H, min_d, max_d, raw);
}
void orc_depth_inverse_dilation(const float* raw, int W, int H, float max_dist, float scale, const uint8_t* mask, int ku, int kv, float* out) {
    inverse_dilation(raw, W, H, max_dist, scale, mask, ku, kv, out);
}
void orc_depth_average_filter(const float* raw, int W, int H, int k, float* out) { average_filter(raw, W, H, k, out); }
void orc_depth_nearest_neighbor_pixel(const float* raw, int W, int H, const void* kps, const void* kps_un, int n, float bf, float R, float* depth, float* uright) {
    nearest_neighbor_pixel(raw, W, H, (const KeyPoint*)kps, (const KeyPoint*)kps_un, n, bf, R, depth, uright);
}
void orc_depth_gather(const float* map, int W, const void* kps, const void* kps_un, int n, float bf, float* depth, float* uright) {
    gather(map, W, (const KeyPoint*)kps, (const KeyPoint*)kps_un, n, bf, depth, uright);
}
// DepthModule::CalculateDepthFromPcd for the InverseDilation method, DepthModule.cc:50-79.
void orc_depth_from_pcd(const float* pts4xn, int n, const float* P, int W, int H, float min_d, float max_d,
                        const uint8_t* mask, int ku, int kv, float bf, const void* kps, const void* kps_un, int n_kp,
                        float* depth, float* uright, float* raw, float* processed) {
    project(pts4xn, n, P, W, H, min_d, max_d, raw);
    inverse_dilation(raw, W, H, max_d, 1.0f, mask, ku, kv, processed);
    gather(processed, W, (const KeyPoint*)kps, (const KeyPoint*)kps_un, n_kp, bf, depth, uright);
}

}  // extern "C"
