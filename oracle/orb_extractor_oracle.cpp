// ORACLE (test infrastructure, NOT product code): CPU restatement of the reference's
// ORBextractor for parity checks. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// Follows /root/reference/src/ORBextractor.cc line by line (cited per function) and restates
// the OpenCV (un-vendored, >=4.4; cv2 4.13.0 probed) primitives it calls in closed form:
// cv::resize(INTER_LINEAR, 8U), cv::FAST(TYPE_9_16, nms), cv::GaussianBlur(7x7, s=2, 8U),
// cv::fastAtan2.  Each primitive is pinned against cv2 in tests/test_oracle_vs_cv2.py and
// against the committed fixtures in tests/golden/.
//
// Canonical float semantics (SURVEY.md §7 hard part 4): IEEE float32, no FMA contraction
// (build with -ffp-contract=off, no -march=native), glibc cosf/sinf.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>

namespace {

const int8_t kPattern[1024] = {
#include "orb_pattern_31.inc"
};

constexpr int kHalfPatch = 15;   // ORBextractor.cc:72
constexpr int kEdge = 19;        // ORBextractor.cc:73
constexpr int kPatch = 31;       // ORBextractor.cc:71

inline int round_half_even(float v) { return (int)lrintf(v); }   // cvRound (SSE cvtss2si)
inline int round_half_even_d(double v) { return (int)lrint(v); }

inline int reflect101(int p, int n) {
    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// cv::resize(src, dst, sz, 0, 0, INTER_LINEAR) for CV_8UC1, non-integer scale (ORBextractor.cc:1183).
// Fixed point: 11-bit coefficients, horizontal pass int32, vertical pass the 8U VResizeLinear form.
// ---------------------------------------------------------------------------------------------
struct LinCoef { int s; int c0, c1; };

std::vector<LinCoef> linear_coefs(int src_n, int dst_n) {
    std::vector<LinCoef> t(dst_n);
    double scale = (double)src_n / (double)dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src_n - 1) { s = src_n - 1; f = 0.f; }
        float w1 = f * 2048.f, w0 = (1.f - f) * 2048.f;
        t[d].s = s;
        t[d].c0 = round_half_even(w0);
        t[d].c1 = round_half_even(w1);
    }
    return t;
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride) {
    std::vector<LinCoef> cx = linear_coefs(sw, dw), cy = linear_coefs(sh, dh);
    std::vector<int> row0(dw), row1(dw);
    for (int y = 0; y < dh; ++y) {
        int sy0 = cy[y].s, sy1 = std::min(sy0 + 1, sh - 1);
        const uint8_t* p0 = src + (size_t)sy0 * sstride;
        const uint8_t* p1 = src + (size_t)sy1 * sstride;
        for (int x = 0; x < dw; ++x) {
            int s0 = cx[x].s, s1 = std::min(s0 + 1, sw - 1);
            row0[x] = p0[s0] * cx[x].c0 + p0[s1] * cx[x].c1;
            row1[x] = p1[s0] * cx[x].c0 + p1[s1] * cx[x].c1;
        }
        int b0 = cy[y].c0, b1 = cy[y].c1;
        for (int x = 0; x < dw; ++x) {
            int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)y * dstride + x] = (uint8_t)v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) on a standalone CV_8UC1 image
// (ORBextractor.cc:1132-1133; the clone makes the level's own edge the reflection axis).
// OpenCV's 8U path: 8.8 fixed-point kernel {18,34,48,56,48,34,18}/256, exact sums, one rounding.
// ---------------------------------------------------------------------------------------------
const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

void gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uint8_t> prow(w + 6);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = src + (size_t)y * sstride;
        for (int k = 0; k < 3; ++k) { prow[k] = r[reflect101(k - 3, w)]; prow[w + 3 + k] = r[reflect101(w + k, w)]; }
        memcpy(&prow[3], r, w);
        uint16_t* ho = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &prow[x];
            ho[x] = (uint16_t)(kGauss7[0] * (p[0] + p[6]) + kGauss7[1] * (p[1] + p[5]) + kGauss7[2] * (p[2] + p[4]) + kGauss7[3] * p[3]);
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* rr[7];
        for (int k = 0; k < 7; ++k) rr[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        uint8_t* o = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            const uint32_t acc = (uint32_t)kGauss7[0] * (rr[0][x] + rr[6][x]) + (uint32_t)kGauss7[1] * (rr[1][x] + rr[5][x]) +
                                 (uint32_t)kGauss7[2] * (rr[2][x] + rr[4][x]) + (uint32_t)kGauss7[3] * rr[3][x];
            o[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(win, kps, th, true)  TYPE_9_16 (called at ORBextractor.cc:826 and :845).
// score = (largest t such that the pixel is still a corner at threshold t); NMS strict '>' over
// the 8 neighbours, pixels outside [3,w-3)x[3,h-3) of THIS window count as 0.  Row-major output.
// ---------------------------------------------------------------------------------------------
const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                          {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// returns K = max over the 16 contiguous 9-arcs of min(d) for bright and dark; corner@th iff K>th;
// cv cornerScore == max(th, K) - 1 (== K-1 for a corner).
inline int fast_arc_strength(const uint8_t* p, int stride) {
    int d[16];
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[kRing[k][1] * stride + kRing[k][0]];
    int best = -256;
    for (int s = 0; s < 16; ++s) {
        int mn = 256, mx = -256;
        for (int j = 0; j < 9; ++j) {
            int e = d[(s + j) & 15];
            mn = std::min(mn, e);
            mx = std::max(mx, e);
        }
        best = std::max(best, mn);    // bright arc: all d > t
        best = std::max(best, -mx);   // dark arc: all d < -t
    }
    return best;
}

struct Cand { int x, y, score; };

void fast_window(const uint8_t* win, int w, int h, int stride, int th, std::vector<Cand>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t* p = win + (size_t)y * stride + x;
            // cheap necessary condition (OpenCV's own high-speed test has the same role): any 9-arc covers
            // at least two of the four compass ring pixels, so a corner at th has >= 2 of them beyond +-th
            const int v = p[0], c0 = p[3 * stride], c4 = p[3], c8 = p[-3 * stride], c12 = p[-3];
            const int nb = (v - c0 > th) + (v - c4 > th) + (v - c8 > th) + (v - c12 > th);
            const int nd = (c0 - v > th) + (c4 - v > th) + (c8 - v > th) + (c12 - v > th);
            if (nb < 2 && nd < 2) continue;
            int K = fast_arc_strength(p, stride);
            if (K > th) sc[(size_t)y * w + x] = K - 1;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (s == 0) continue;
            const int* c = &sc[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                s > c[w - 1] && s > c[w] && s > c[w + 1])
                out.push_back({x, y, s});
        }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x), scalar float32 path, degrees (called at ORBextractor.cc:102).
// ---------------------------------------------------------------------------------------------
float fast_atan2_deg(float y, float x) {
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k,
                p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.2204460492503131e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.2204460492503131e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

struct KeyPoint {      // cv::KeyPoint memory layout (28 bytes)
    float x, y, size, angle, response;
    int octave, class_id;
};

// ---------------------------------------------------------------------------------------------
// Quad-tree distribution, ORBextractor.cc:480-536 (DivideNode), :538-553 (compareNodes),
// :555-779 (DistributeOctTree).  std::list / std::sort are used so that libstdc++'s container
// and introsort behaviour (tie order) is inherited rather than emulated.
// ---------------------------------------------------------------------------------------------
struct QNode {
    int ulx, uly, urx, ury, blx, bly, brx, bry;
    std::vector<KeyPoint> keys;
    bool leaf = false;
    std::list<QNode>::iterator self;
};

void split_node(const QNode& n, QNode c[4]) {
    const int hx = (int)std::ceil(static_cast<float>(n.urx - n.ulx) / 2);
    const int hy = (int)std::ceil(static_cast<float>(n.bry - n.uly) / 2);
    c[0].ulx = n.ulx; c[0].uly = n.uly; c[0].urx = n.ulx + hx; c[0].ury = n.uly;
    c[0].blx = n.ulx; c[0].bly = n.uly + hy; c[0].brx = n.ulx + hx; c[0].bry = n.uly + hy;
    c[1].ulx = c[0].urx; c[1].uly = c[0].ury; c[1].urx = n.urx; c[1].ury = n.ury;
    c[1].blx = c[0].brx; c[1].bly = c[0].bry; c[1].brx = n.urx; c[1].bry = n.uly + hy;
    c[2].ulx = c[0].blx; c[2].uly = c[0].bly; c[2].urx = c[0].brx; c[2].ury = c[0].bry;
    c[2].blx = n.blx; c[2].bly = n.bly; c[2].brx = c[0].brx; c[2].bry = n.bly;
    c[3].ulx = c[2].urx; c[3].uly = c[2].ury; c[3].urx = c[1].brx; c[3].ury = c[1].bry;
    c[3].blx = c[2].brx; c[3].bly = c[2].bry; c[3].brx = n.brx; c[3].bry = n.bry;
    for (const KeyPoint& kp : n.keys) {
        int q;
        if (kp.x < c[0].urx) q = (kp.y < c[0].bry) ? 0 : 2;
        else q = (kp.y < c[0].bry) ? 1 : 3;
        c[q].keys.push_back(kp);
    }
    for (int q = 0; q < 4; ++q) c[q].leaf = (c[q].keys.size() == 1);
}

typedef std::pair<int, QNode*> SizedNode;
bool sized_less(SizedNode& a, SizedNode& b) {
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ulx < b.second->ulx;
}

std::vector<KeyPoint> distribute_quadtree(const std::vector<KeyPoint>& in, int minX, int maxX,
                                          int minY, int maxY, int N) {
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<QNode> nodes;
    std::vector<QNode*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        QNode r;
        r.ulx = (int)(hX * static_cast<float>(i)); r.uly = 0;
        r.urx = (int)(hX * static_cast<float>(i + 1)); r.ury = 0;
        r.blx = r.ulx; r.bly = maxY - minY;
        r.brx = r.urx; r.bry = maxY - minY;
        nodes.push_back(r);
        roots[i] = &nodes.back();
    }
    for (const KeyPoint& kp : in) roots[(int)(kp.x / hX)]->keys.push_back(kp);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->leaf = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    std::vector<SizedNode> open;
    auto push_children = [&](QNode c[4], int& nExpand) {
        for (int q = 0; q < 4; ++q) {
            if (c[q].keys.empty()) continue;
            nodes.push_front(c[q]);
            if (c[q].keys.size() > 1) {
                ++nExpand;
                open.push_back(std::make_pair((int)c[q].keys.size(), &nodes.front()));
                nodes.front().self = nodes.begin();
            }
        }
    };
    bool done = false;
    while (!done) {
        int prev = (int)nodes.size();
        int nExpand = 0;
        open.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            QNode c[4];
            split_node(*it, c);
            push_children(c, nExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev) {
            done = true;
        } else if ((int)nodes.size() + nExpand * 3 > N) {
            while (!done) {
                prev = (int)nodes.size();
                std::vector<SizedNode> work = open;
                open.clear();
                std::sort(work.begin(), work.end(), sized_less);
                for (int j = (int)work.size() - 1; j >= 0; --j) {
                    QNode c[4];
                    int dummy = 0;
                    split_node(*work[j].second, c);
                    push_children(c, dummy);
                    nodes.erase(work[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev) done = true;
            }
        }
    }
    std::vector<KeyPoint> out;
    out.reserve(nodes.size());
    for (QNode& n : nodes) {
        const KeyPoint* best = &n.keys[0];
        float r = best->response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].response > r) { best = &n.keys[k]; r = n.keys[k].response; }
        out.push_back(*best);
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Extractor object state (ORBextractor::ORBextractor, ORBextractor.cc:409-469).
// ---------------------------------------------------------------------------------------------
struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactor;
    std::vector<float> scale, invScale;
    std::vector<int> quota;
    int umax[kHalfPatch + 1];
    // pyramid (un-padded level images; the 19 px border is not read by this path)
    std::vector<std::vector<uint8_t>> lvl;
    std::vector<int> lw, lh;
    std::vector<std::vector<Cand>> cands;         // per level, after cell offset (rel. to minBorder)
    std::vector<std::vector<KeyPoint>> kps;       // per level, level coordinates, with angle

    void init(int nf, float sf, int nl, int ini, int mn) {
        nfeatures = nf; scaleFactor = sf; nlevels = nl; iniTh = ini; minTh = mn;
        scale.assign(nl, 1.f); invScale.assign(nl, 1.f);
        for (int i = 1; i < nl; ++i) scale[i] = scale[i - 1] * sf;
        for (int i = 0; i < nl; ++i) invScale[i] = 1.0f / scale[i];
        quota.assign(nl, 0);
        float factor = 1.0f / sf;
        float per = nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
        int sum = 0;
        for (int l = 0; l < nl - 1; ++l) {
            quota[l] = round_half_even(per);
            sum += quota[l];
            per *= factor;
        }
        quota[nl - 1] = std::max(nf - sum, 0);
        int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (int v = 0; v <= vmax; ++v) umax[v] = round_half_even_d(std::sqrt(hp2 - v * v));
        for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // ORBextractor::ComputePyramid, ORBextractor.cc:1170-1195
    void pyramid(const uint8_t* img, int w, int h, int stride) {
        lvl.resize(nlevels); lw.resize(nlevels); lh.resize(nlevels);
        for (int l = 0; l < nlevels; ++l) {
            lw[l] = round_half_even((float)w * invScale[l]);
            lh[l] = round_half_even((float)h * invScale[l]);
            lvl[l].assign((size_t)lw[l] * lh[l], 0);
            if (l == 0) {
                for (int y = 0; y < h; ++y) memcpy(&lvl[0][(size_t)y * w], img + (size_t)y * stride, w);
            } else {
                resize_linear_u8(lvl[l - 1].data(), lw[l - 1], lh[l - 1], lw[l - 1],
                                 lvl[l].data(), lw[l], lh[l], lw[l]);
            }
        }
    }

    // IC_Angle, ORBextractor.cc:76-103
    float ic_angle(const uint8_t* img, int stride, float px, float py) const {
        int m01 = 0, m10 = 0;
        const uint8_t* c = img + (size_t)round_half_even(py) * stride + round_half_even(px);
        for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
        for (int v = 1; v <= kHalfPatch; ++v) {
            int vs = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int a = c[u + v * stride], b = c[u - v * stride];
                vs += a - b;
                m10 += u * (a + b);
            }
            m01 += v * vs;
        }
        return fast_atan2_deg((float)m01, (float)m10);
    }

    // ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:781-896
    void keypoints() {
        cands.assign(nlevels, {});
        kps.assign(nlevels, {});
        const float W = 35;
        std::vector<Cand> cell;
        for (int l = 0; l < nlevels; ++l) {
            const int minBX = kEdge - 3, minBY = minBX;
            const int maxBX = lw[l] - kEdge + 3, maxBY = lh[l] - kEdge + 3;
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            std::vector<KeyPoint> toDist;
            for (int i = 0; i < nRows; ++i) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; ++j) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                    const uint8_t* win = lvl[l].data() + (size_t)y0 * lw[l] + x0;
                    fast_window(win, cw, ch, lw[l], iniTh, cell);
                    if (cell.empty()) fast_window(win, cw, ch, lw[l], minTh, cell);
                    for (const Cand& c : cell) {
                        KeyPoint kp;
                        kp.x = (float)c.x + j * wCell;
                        kp.y = (float)c.y + i * hCell;
                        kp.size = 7.f; kp.angle = -1.f; kp.response = (float)c.score;
                        kp.octave = 0; kp.class_id = -1;
                        toDist.push_back(kp);
                        cands[l].push_back({(int)kp.x, (int)kp.y, c.score});
                    }
                }
            }
            std::vector<KeyPoint> sel = distribute_quadtree(toDist, minBX, maxBX, minBY, maxBY, quota[l]);
            const int scaledPatch = (int)(kPatch * scale[l]);
            for (KeyPoint& kp : sel) {
                kp.x += minBX; kp.y += minBY;
                kp.octave = l;
                kp.size = (float)scaledPatch;
            }
            kps[l] = sel;
        }
        for (int l = 0; l < nlevels; ++l)
            for (KeyPoint& kp : kps[l]) kp.angle = ic_angle(lvl[l].data(), lw[l], kp.x, kp.y);
    }
};

// computeOrbDescriptor, ORBextractor.cc:107-146
void orb_descriptor(const KeyPoint& kp, const uint8_t* img, int stride, uint8_t* desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = kp.angle * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = img + (size_t)round_half_even(kp.y) * stride + round_half_even(kp.x);
    const int8_t* pat = kPattern;
    auto sample = [&](int idx) -> int {
        float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        float fr = px * b + py * a;
        float fc = px * a - py * b;
        return c[round_half_even(fr) * stride + round_half_even(fc)];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; ++j) {
            int t0 = sample(2 * j), t1 = sample(2 * j + 1);
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

}  // namespace

extern "C" {

struct orc_orb_params { int nfeatures; float scale_factor; int nlevels; int ini_th; int min_th; int lap0; int lap1; };

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}
void orc_gaussian_blur7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    gaussian_blur7_u8(src, w, h, sstride, dst, dstride);
}
// out: triples (x, y, score) row-major; returns count (<= cap written)
int orc_fast_window(const uint8_t* win, int w, int h, int stride, int th, int* out, int cap) {
    std::vector<Cand> c;
    fast_window(win, w, h, stride, th, c);
    int n = (int)c.size();
    for (int i = 0; i < n && i < cap; ++i) { out[3 * i] = c[i].x; out[3 * i + 1] = c[i].y; out[3 * i + 2] = c[i].score; }
    return n;
}
float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }

// Quad-tree alone: in/out are 28-byte keypoints; coordinates relative to (minX,minY).
int orc_distribute_quadtree(const KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, KeyPoint* out, int cap) {
    std::vector<KeyPoint> v(in, in + n);
    std::vector<KeyPoint> r = distribute_quadtree(v, minX, maxX, minY, maxY, N);
    int m = (int)r.size();
    for (int i = 0; i < m && i < cap; ++i) out[i] = r[i];
    return m;
}

struct orc_extractor;   // opaque = Extractor
void* orc_extractor_create(const orc_orb_params* p) {
    Extractor* e = new Extractor();
    e->init(p->nfeatures, p->scale_factor, p->nlevels, p->ini_th, p->min_th);
    return e;
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }
void orc_extractor_tables(void* h, float* scale, float* inv_scale, int* quota, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int l = 0; l < e->nlevels; ++l) { scale[l] = e->scale[l]; inv_scale[l] = e->invScale[l]; quota[l] = e->quota[l]; }
    for (int v = 0; v <= kHalfPatch; ++v) umax16[v] = e->umax[v];
}

// ORBextractor::operator(), ORBextractor.cc:1086-1168.  Returns monoIndex (or -1 on empty image).
int orc_extract(void* h, const uint8_t* img, int w, int hgt, int stride, int lap0, int lap1,
                KeyPoint* kps_out, uint8_t* desc_out, int cap, int* n_out) {
    Extractor* e = (Extractor*)h;
    *n_out = 0;
    if (!img || w <= 0 || hgt <= 0) return -1;
    e->pyramid(img, w, hgt, stride);
    // DistributeOctTree computes nIni = round(width / height) of the FAST window and divides by it (src/ORBextractor.cc:559-561):
    // for a level taller than about twice its width nIni is 0 and the reference indexes an empty vector (undefined behaviour).
    // The oracle refuses such images instead of restating a crash; the product reports RGBL_E_UNSUPPORTED for them.
    for (int l = 0; l < e->nlevels; ++l)
        if ((int)std::round(static_cast<float>(e->lw[l] - 32) / (e->lh[l] - 32)) < 1) return -5;
    e->keypoints();
    int total = 0;
    for (int l = 0; l < e->nlevels; ++l) total += (int)e->kps[l].size();
    *n_out = total;
    if (total > cap) return -2;
    int mono = 0, stereo = total - 1;
    std::vector<uint8_t> blur;
    for (int l = 0; l < e->nlevels; ++l) {
        std::vector<KeyPoint>& k = e->kps[l];
        if (k.empty()) continue;
        blur.resize((size_t)e->lw[l] * e->lh[l]);
        gaussian_blur7_u8(e->lvl[l].data(), e->lw[l], e->lh[l], e->lw[l], blur.data(), e->lw[l]);
        float sc = e->scale[l];
        for (KeyPoint kp : k) {
            uint8_t d[32];
            orb_descriptor(kp, blur.data(), e->lw[l], d);
            if (l != 0) { kp.x *= sc; kp.y *= sc; }
            int slot;
            if (kp.x >= lap0 && kp.x <= lap1) slot = stereo--;
            else slot = mono++;
            kps_out[slot] = kp;
            memcpy(desc_out + (size_t)slot * 32, d, 32);
        }
    }
    return mono;
}

// Introspection for stage-level parity tests (valid after orc_extract).
int orc_level_size(void* h, int l, int* w, int* hgt) { Extractor* e = (Extractor*)h; *w = e->lw[l]; *hgt = e->lh[l]; return 0; }
const uint8_t* orc_level_ptr(void* h, int l) { return ((Extractor*)h)->lvl[l].data(); }
int orc_level_candidates(void* h, int l, int* out, int cap) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->cands[l].size();
    for (int i = 0; i < n && i < cap; ++i) { out[3 * i] = e->cands[l][i].x; out[3 * i + 1] = e->cands[l][i].y; out[3 * i + 2] = e->cands[l][i].score; }
    return n;
}
int orc_level_keypoints(void* h, int l, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    int n = (int)e->kps[l].size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = e->kps[l][i];
    return n;
}
void orc_descriptor(const KeyPoint* kp, const uint8_t* blurred, int stride, uint8_t* desc) { orb_descriptor(*kp, blurred, stride, desc); }

}  // extern "C"
