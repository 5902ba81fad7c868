// ORB extractor kernels for sm_100a: pyramid, per-cell FAST-9/16 + cell-local NMS, candidate
// compaction, 7x7 Gaussian pre-filter, and the fused orientation + rBRIEF describe kernel.
// Reference semantics: src/ORBextractor.cc (cited per kernel) with the OpenCV primitives restated
// in integer/fixed-point form.  Every frame of a batch is processed by the
// same launch (blockIdx.z / blockIdx.y = frame slot).
#include "rgbl_device.cuh"
#include "rgbl_kernels.h"

namespace rgbl {

// ------------------------------------------------------------------------------------------------
// Pyramid: level l = cv::resize(level l-1, INTER_LINEAR)  (src/ORBextractor.cc:1183).
// One thread produces 4 horizontally adjacent output bytes (one 32-bit store).
// ------------------------------------------------------------------------------------------------
// CTA tile: 128 x 16 outputs.  The source rectangle of the tile (<= 22 rows x 176 bytes at scale 1.2) is staged in shared
// memory with 16-byte loads (rows are 64-byte aligned); every thread then produces 4 adjacent outputs in two rows from
// shared-memory bytes.  (The first version gathered 16 single bytes per thread straight from global memory and was bound by
// the number of load instructions, not by HBM.)
constexpr int kRzTW = 128, kRzTH = 16, kRzSrcRows = 24, kRzSrcVec = 12;     // staged source: 24 rows x 192 bytes max
__global__ void __launch_bounds__(256) resize_level_kernel(uint8_t* __restrict__ pyr, size_t frame_stride,
                                                           LevelGeom src, LevelGeom dst,
                                                           const LinCoef* __restrict__ tabx,
                                                           const LinCoef* __restrict__ taby) {
    __shared__ uint4 tile[kRzSrcRows][kRzSrcVec];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int x0 = blockIdx.x * kRzTW, y0 = blockIdx.y * kRzTH;
    const uint8_t* s = pyr + (size_t)blockIdx.z * frame_stride + src.off;
    uint8_t* d = pyr + (size_t)blockIdx.z * frame_stride + dst.off;
    // source rectangle of this tile (tables are monotone)
    const int xl = min(x0 + kRzTW - 1, dst.w - 1), yl = min(y0 + kRzTH - 1, dst.h - 1);
    const int sx_lo = tabx[x0].s & ~15, sx_hi = min(tabx[xl].s + 1, src.w - 1);
    const int sy_lo = taby[y0].s, sy_hi = min(taby[yl].s + 1, src.h - 1);
    const int n_vec = min(min((sx_hi - sx_lo) / 16 + 1, kRzSrcVec), (src.pitch - sx_lo) / 16);
    const int n_rows = min(sy_hi - sy_lo + 1, kRzSrcRows);
    for (int i = tid; i < n_rows * n_vec; i += 256) {
        const int r = i / n_vec, v = i - r * n_vec;
        tile[r][v] = __ldg(reinterpret_cast<const uint4*>(s + (size_t)(sy_lo + r) * src.pitch + sx_lo) + v);
    }
    __syncthreads();
    const int x4 = x0 + 4 * tx;
    if (x4 >= dst.w) return;
    const uint8_t* tb = reinterpret_cast<const uint8_t*>(&tile[0][0]);
    constexpr int kTP = kRzSrcVec * 16;
    int s0[4], s1[4], c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const LinCoef cx = tabx[min(x4 + i, dst.w - 1)];
        s0[i] = cx.s - sx_lo; s1[i] = min(cx.s + 1, src.w - 1) - sx_lo; c0[i] = cx.c0; c1[i] = cx.c1;
    }
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
        const int y = y0 + ty + 8 * ry;
        if (y >= dst.h) break;
        const LinCoef cy = taby[y];
        const uint8_t* r0 = tb + (cy.s - sy_lo) * kTP;
        const uint8_t* r1 = tb + (min(cy.s + 1, src.h - 1) - sy_lo) * kTP;
        const int b0 = cy.c0, b1 = cy.c1;
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (x4 + i < dst.w) {
                const int h0 = (int)r0[s0[i]] * c0[i] + (int)r0[s1[i]] * c1[i];
                const int h1 = (int)r1[s0[i]] * c0[i] + (int)r1[s1[i]] * c1[i];
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                out |= (uint32_t)(v & 0xff) << (8 * i);
            }
        }
        *reinterpret_cast<uint32_t*>(d + (size_t)y * dst.pitch + x4) = out;
    }
}

// Fallback for level ratios whose source rectangle does not fit the staged tile (scale factors above ~1.35): one thread
// gathers the 16 source bytes of 4 adjacent outputs directly from global memory.
__global__ void __launch_bounds__(256) resize_level_gather_kernel(uint8_t* __restrict__ pyr, size_t frame_stride,
                                                                  LevelGeom src, LevelGeom dst,
                                                                  const LinCoef* __restrict__ tabx,
                                                                  const LinCoef* __restrict__ taby) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= dst.w || y >= dst.h) return;
    const uint8_t* s = pyr + (size_t)blockIdx.z * frame_stride + src.off;
    uint8_t* d = pyr + (size_t)blockIdx.z * frame_stride + dst.off;
    const LinCoef cy = taby[y];
    const uint8_t* r0 = s + (size_t)cy.s * src.pitch;
    const uint8_t* r1 = s + (size_t)min(cy.s + 1, src.h - 1) * src.pitch;
    const int b0 = cy.c0, b1 = cy.c1;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x4 + i;
        if (x < dst.w) {
            const LinCoef cx = tabx[x];
            const int s0 = cx.s, s1 = min(s0 + 1, src.w - 1);
            const int h0 = (int)__ldg(r0 + s0) * cx.c0 + (int)__ldg(r0 + s1) * cx.c1;
            const int h1 = (int)__ldg(r1 + s0) * cx.c0 + (int)__ldg(r1 + s1) * cx.c1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xff) << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t*>(d + (size_t)y * dst.pitch + x4) = out;
}

// ------------------------------------------------------------------------------------------------
// FAST: one CTA per (cell, frame).  Equivalent of the per-cell cv::FAST(th=iniTh, nms) with the
// cv::FAST(th=minTh, nms) fallback when the first call returns nothing (src/ORBextractor.cc:805-868).
//   score(p) = K-1 with K the arc strength (corner at threshold t  <=>  K > t  <=>  score >= t);
//   NMS survivors at minTh = {s >= minTh, s > all 8 neighbours (non-corners / outside the window = 0)};
//   survivors at iniTh = survivors at minTh with s >= iniTh (a weaker neighbour never suppresses).
// Survivors are written in row-major order (the order cv::FAST returns) into the cell's slot array.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool fast_quick_test(const uint8_t* p, int th) {
    constexpr int P = kFastTilePitch;
    const int v = p[0];
    const int r0 = p[3 * P], r4 = p[3], r8 = p[-3 * P], r12 = p[-3];
    const int hi_t = v + th, lo_t = v - th;
    const int ndark = (r0 < lo_t) + (r4 < lo_t) + (r8 < lo_t) + (r12 < lo_t);
    const int nbright = (r0 > hi_t) + (r4 > hi_t) + (r8 > hi_t) + (r12 > hi_t);
    return ndark >= 2 || nbright >= 2;           // a 9-arc always covers >= 2 of the 4 compass points
}

__device__ __forceinline__ int fast_score_at(const uint8_t* p, int th) {
    constexpr int P = kFastTilePitch;
    const int v = p[0];
    int r[16];
    r[0] = p[3 * P]; r[4] = p[3]; r[8] = p[-3 * P]; r[12] = p[-3];
    r[1] = p[3 * P + 1];  r[2] = p[2 * P + 2];   r[3] = p[P + 3];
    r[5] = p[-P + 3];     r[6] = p[-2 * P + 2];  r[7] = p[-3 * P + 1];
    r[9] = p[-3 * P - 1]; r[10] = p[-2 * P - 2]; r[11] = p[-P - 3];
    r[13] = p[P - 3];     r[14] = p[2 * P - 2];  r[15] = p[3 * P - 1];
    const int K = fast_arc_strength16(v, r);
    return (K > th) ? (K - 1) : 0;
}

// Phases: (1) window -> shared memory with aligned 32-bit loads; (2) high-speed test on every tested pixel,
// survivors appended (warp ballot) to a shared list; (3) full arc score only for the listed pixels, densely
// packed across the CTA (the heavy path no longer runs divergently on whole warps); (4) cell-local NMS +
// threshold fallback + ordered emission.
__global__ void __launch_bounds__(256) fast_cells_kernel(const uint8_t* __restrict__ pyr, size_t frame_stride,
                                                         const LevelGeom* __restrict__ levels,
                                                         const CellInfo* __restrict__ cells, int n_cells,
                                                         int ini_th, int min_th,
                                                         uint32_t* __restrict__ slots, int* __restrict__ counts,
                                                         int* __restrict__ overflow) {
    constexpr int P = kFastTilePitch;            // 88: 3 alignment bytes + 78-byte window, multiple of 4
    constexpr int ROWS = kFastTilePitch - 8;     // 80 >= max window height 78
    __shared__ __align__(16) uint8_t tile[P * ROWS];
    __shared__ __align__(16) uint8_t sc[P * ROWS];
    __shared__ uint16_t list[72 * 72 + 32];
    __shared__ int warp_sums[8];
    __shared__ int n_list;

    const int cell = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const CellInfo ci = cells[cell];
    const LevelGeom lg = levels[ci.level];
    const int cw = ci.cw, ch = ci.ch;
    const int a = ci.x0 & 3;                      // byte offset of the window inside its first aligned word
    const uint8_t* src = pyr + (size_t)frame * frame_stride + lg.off + (size_t)ci.y0 * lg.pitch + (ci.x0 - a);
    const int nw = (cw + a + 3) >> 2;             // words per row (<= 21)

    if (tid == 0) n_list = 0;
    for (int y = warp; y < ch; y += 8) {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(src + (size_t)y * lg.pitch);
        uint32_t* t = reinterpret_cast<uint32_t*>(tile + y * P);
        uint32_t* z = reinterpret_cast<uint32_t*>(sc + y * P);
        if (lane < nw) { t[lane] = __ldg(g + lane); z[lane] = 0; }
    }
    __syncthreads();

    const int tw = cw - 6, th = ch - 6;          // tested region [3, cw-3) x [3, ch-3)
    const int n_t = (tw > 0 && th > 0) ? tw * th : 0;
    {
        // incremental (x, y) of flattened index i = tid + 256*k without per-iteration divisions
        const int q = (tw > 0) ? 256 / tw : 0, r = (tw > 0) ? 256 - q * tw : 0;
        int y = (tw > 0) ? tid / tw : 0, x = tid - y * tw;
        for (int i0 = 0; i0 < n_t; i0 += 256) {
            const bool valid = (i0 + tid) < n_t;
            bool pass = false;
            int pos = 0;
            if (valid) {
                pos = (y + 3) * P + x + 3 + a;
                pass = fast_quick_test(&tile[pos], min_th);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            int base = 0;
            if (lane == 0 && m) base = atomicAdd(&n_list, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)pos;
            x += r; y += q;
            if (x >= tw) { x -= tw; ++y; }
        }
    }
    __syncthreads();
    const int nl = n_list;
    for (int j = tid; j < nl; j += 256) {
        const int pos = list[j];
        sc[pos] = (uint8_t)fast_score_at(&tile[pos], min_th);
    }
    __syncthreads();

    // contiguous row-major chunk per thread -> ordered emission
    const int per = (n_t + 255) / 256;           // <= 21 for windows <= 78 px
    const int beg = min(tid * per, n_t), end = min(beg + per, n_t);
    uint32_t m_min = 0, m_ini = 0;
    int y0c = (tw > 0) ? beg / tw : 0, x0c = beg - y0c * tw;
    {
        int y = y0c, x = x0c;
        for (int i = beg; i < end; ++i) {
            const uint8_t* c = &sc[(y + 3) * P + x + 3 + a];
            const int s = c[0];
            if (s != 0) {
                const bool keep = s > c[-1] && s > c[1] && s > c[-P - 1] && s > c[-P] && s > c[-P + 1] &&
                                  s > c[P - 1] && s > c[P] && s > c[P + 1];
                if (keep) {
                    m_min |= 1u << (i - beg);
                    if (s >= ini_th) m_ini |= 1u << (i - beg);
                }
            }
            if (++x == tw) { x = 0; ++y; }
        }
    }
    const int any_ini = __syncthreads_or(m_ini != 0);
    const uint32_t m = any_ini ? m_ini : m_min;
    const int cnt = __popc(m);

    // block-wide exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int ws = warp_sums[w];
        if (w < warp) base += ws;
        total += ws;
    }
    int pos = base + incl - cnt;
    uint32_t* out = slots + ((size_t)frame * n_cells + cell) * kCellCap;
    uint32_t mm = m;
    while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        int x = x0c + b, y = y0c;
        while (x >= tw) { x -= tw; ++y; }
        if (pos < kCellCap) out[pos] = pack_cand_dev(x + 3 + ci.off_x, y + 3 + ci.off_y, sc[(y + 3) * P + x + 3 + a]);
        ++pos;
    }
    if (tid == 0) {
        counts[(size_t)frame * n_cells + cell] = min(total, kCellCap);
        if (total > kCellCap) atomicExch(overflow, 1);
    }
}

// ------------------------------------------------------------------------------------------------
// Candidate compaction.  Pass A (one CTA per frame): exclusive scan of the cell counts in reference
// cell order -> per-cell offsets, per-level counts/offsets, frame total.  Pass B (grid of CTAs):
// every CTA recomputes the tiny prefix over frame totals and copies its cells' slot arrays into one
// dense buffer, so the host fetches exactly sum(total) candidates with a single copy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cand_scan_kernel(const int* __restrict__ counts, int n_cells,
                                                        const LevelGeom* __restrict__ levels, int n_levels,
                                                        int* __restrict__ cell_off, int* __restrict__ level_cnt,
                                                        int* __restrict__ frame_total) {
    __shared__ int warp_sums[8];
    __shared__ int carry;
    const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* c = counts + (size_t)frame * n_cells;
    int* o = cell_off + (size_t)frame * n_cells;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < n_cells; b += 256) {
        const int i = b + tid;
        const int v = (i < n_cells) ? c[i] : 0;
        int incl = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        int base = carry;
        for (int w = 0; w < warp; ++w) base += warp_sums[w];
        if (i < n_cells) o[i] = base + incl - v;
        __syncthreads();
        if (tid == 255) carry = base + incl;
        __syncthreads();
    }
    if (tid < n_levels) {
        const LevelGeom lg = levels[tid];
        int s = 0;
        for (int k = 0; k < lg.n_cells; ++k) s += c[lg.cell_base + k];
        level_cnt[frame * RGBL_MAX_LEVELS + tid] = s;
    }
    if (tid == 0) frame_total[frame] = carry;
}

__global__ void __launch_bounds__(256) cand_gather_kernel(const uint32_t* __restrict__ slots,
                                                          const int* __restrict__ counts,
                                                          const int* __restrict__ cell_off,
                                                          const int* __restrict__ frame_total, int n_cells,
                                                          uint32_t* __restrict__ dense, int dense_cap,
                                                          int* __restrict__ overflow) {
    const int frame = blockIdx.y;
    int fbase = 0;
    for (int f = 0; f < frame; ++f) fbase += frame_total[f];
    const int warp_global = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * 8;
    for (int cell = warp_global; cell < n_cells; cell += n_warps) {
        const int n = counts[(size_t)frame * n_cells + cell];
        const int off = fbase + cell_off[(size_t)frame * n_cells + cell];
        const uint32_t* s = slots + ((size_t)frame * n_cells + cell) * kCellCap;
        for (int k = lane; k < n; k += 32) {
            if (off + k < dense_cap) dense[off + k] = s[k];
            else atomicExch(overflow, 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on each un-padded level (src/ORBextractor.cc:
// 1132-1133): 8.8 fixed-point kernel {18,34,48,56,48,34,18}, exact separable sums, one
// final rounding.  CTA tile: 64 x 32 outputs; each thread writes 4 adjacent bytes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) blur_level_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
                                                         size_t frame_stride, LevelGeom lg) {
    // CTA tile: 128 x 32 outputs.  Input rows are staged as 32-bit words (aligned loads in the interior, per-byte
    // REFLECT_101 assembly on the level's edges); the horizontal pass keeps 8.8 sums packed as u16x2; the vertical
    // pass gives every thread a 4 x 4 output block (one 32-bit store per row).
    constexpr int TW = 128, TH = 32, IH = TH + 6, IWW = (TW + 8) / 4;     // 34 input words per row: x0-4 .. x0+131
    __shared__ uint32_t in[IH][IWW + 1];
    __shared__ uint32_t hb[IH][TW / 2 + 1];
    const uint8_t* s = pyr + (size_t)blockIdx.z * frame_stride + lg.off;
    uint8_t* d = blur + (size_t)blockIdx.z * frame_stride + lg.off;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int r = warp; r < IH; r += 8) {
        const int gy = reflect101(y0 + r - 3, lg.h);
        const uint8_t* row = s + (size_t)gy * lg.pitch;
        for (int w = lane; w < IWW; w += 32) {
            const int gx = x0 - 4 + 4 * w;
            uint32_t v;
            if (gx >= 0 && gx + 3 < lg.w) {
                v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
            } else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int px = gx + k;
                    px = reflect101(px, lg.w);
                    px = min(max(px, 0), lg.w - 1);             // tiles hanging far over the right edge: value unused
                    v |= (uint32_t)__ldg(row + px) << (8 * k);
                }
            }
            in[r][w] = v;
        }
    }
    __syncthreads();
    for (int item = tid; item < IH * 32; item += 256) {
        const int r = item >> 5, g = item & 31;                 // g: group of 4 output columns
        const uint32_t w0 = in[r][g], w1 = in[r][g + 1], w2 = in[r][g + 2];
        int b[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) { b[k] = (w0 >> (8 * k)) & 0xff; b[4 + k] = (w1 >> (8 * k)) & 0xff; b[8 + k] = (w2 >> (8 * k)) & 0xff; }
        uint32_t acc[4];
#pragma unroll
        for (int jx = 0; jx < 4; ++jx)
            acc[jx] = 18 * (b[1 + jx] + b[7 + jx]) + 34 * (b[2 + jx] + b[6 + jx]) + 48 * (b[3 + jx] + b[5 + jx]) + 56 * b[4 + jx];
        hb[r][2 * g] = acc[0] | (acc[1] << 16);
        hb[r][2 * g + 1] = acc[2] | (acc[3] << 16);
    }
    __syncthreads();
    {
        const int g = lane, rg = warp;                          // 32 column groups x 8 row groups of 4 rows
        const int gx = x0 + 4 * g;
        if (gx >= lg.w) return;
        uint32_t c0[10], c1[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { c0[k] = hb[4 * rg + k][2 * g]; c1[k] = hb[4 * rg + k][2 * g + 1]; }
#pragma unroll
        for (int ry = 0; ry < 4; ++ry) {
            const int gy = y0 + 4 * rg + ry;
            if (gy >= lg.h) break;
            uint32_t out = 0;
#pragma unroll
            for (int jx = 0; jx < 4; ++jx) {
                uint32_t h[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const uint32_t w = (jx < 2) ? c0[ry + k] : c1[ry + k];
                    h[k] = (jx & 1) ? (w >> 16) : (w & 0xffffu);
                }
                const uint32_t acc = 18u * (h[0] + h[6]) + 34u * (h[1] + h[5]) + 48u * (h[2] + h[4]) + 56u * h[3];
                out |= ((acc + 32768u) >> 16) << (8 * jx);
            }
            *reinterpret_cast<uint32_t*>(d + (size_t)gy * lg.pitch + gx) = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Describe: one warp per selected keypoint.  IC_Angle (src/ORBextractor.cc:76-103) on the un-blurred
// level, then computeOrbDescriptor (:107-146) on the blurred level, then the cv::KeyPoint record with
// the level->image coordinate scaling of operator() (:1143-1151).
// ------------------------------------------------------------------------------------------------
__device__ const int8_t g_pattern[1024] = {
#include "orb_pattern_31.inc"
};

struct UmaxTable { int v[16]; };

__global__ void __launch_bounds__(256) describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                       size_t frame_stride, const LevelGeom* __restrict__ levels,
                                                       const SelKp* __restrict__ sel, const int* __restrict__ n_sel,
                                                       int cap, UmaxTable umax, rgbl_keypoint* __restrict__ kps,
                                                       uint8_t* __restrict__ desc) {
    const int frame = blockIdx.y;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n_sel[frame]) return;
    const SelKp kp = sel[(size_t)frame * cap + k];
    const LevelGeom lg = levels[kp.level];
    const uint8_t* img = pyr + (size_t)frame * frame_stride + lg.off;
    const uint8_t* c = img + (size_t)kp.y * lg.pitch + kp.x;

    // intensity centroid: lane <-> row v = lane-15 of the radius-15 disc (lane 31 idle)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - kHalfPatch;
        const int dmax = umax.v[v < 0 ? -v : v];
        const uint8_t* row = c + v * lg.pitch;
        int rs = 0;
        for (int u = -dmax; u <= dmax; ++u) {
            const int val = __ldg(row + u);
            m10 += u * val;
            rs += val;
        }
        m01 = v * rs;
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // steered BRIEF: lane i computes descriptor byte i (8 comparisons, 16 samples)
    const float factor_pi = (float)(3.14159265358979323846 / 180.0);
    float a, b;
    glibc_sincosf(__fmul_rn(angle, factor_pi), &b, &a);       // a = cos, b = sin
    const uint8_t* bc = blur + (size_t)frame * frame_stride + lg.off + (size_t)kp.y * lg.pitch + kp.x;
    const int8_t* pat = g_pattern + lane * 32;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int t[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float px = (float)pat[4 * j + 2 * e], py = (float)pat[4 * j + 2 * e + 1];
            const int rr = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int cc = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            t[e] = __ldg(bc + rr * lg.pitch + cc);
        }
        val |= (t[0] < t[1]) << j;
    }
    desc[((size_t)frame * cap + k) * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        rgbl_keypoint o;
        const float fx = (float)kp.x, fy = (float)kp.y;
        o.x = (kp.level != 0) ? __fmul_rn(fx, lg.scale) : fx;
        o.y = (kp.level != 0) ? __fmul_rn(fy, lg.scale) : fy;
        o.size = (float)lg.scaled_patch;
        o.angle = angle;
        o.response = (float)kp.score;
        o.octave = kp.level;
        o.class_id = -1;
        kps[(size_t)frame * cap + k] = o;
    }
}

// mvImagePyramid[level] with its 19 px BORDER_REFLECT_101 frame (src/ORBextractor.cc:1185-1191),
// synthesised on demand: the hot path itself never reads the border.
__global__ void padded_level_kernel(const uint8_t* __restrict__ pyr, size_t frame_stride, int frame, LevelGeom lg,
                                    uint8_t* __restrict__ dst, int dst_pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int W = lg.w + 2 * kEdgeThreshold, H = lg.h + 2 * kEdgeThreshold;
    if (x >= W || y >= H) return;
    const uint8_t* s = pyr + (size_t)frame * frame_stride + lg.off;
    dst[(size_t)y * dst_pitch + x] = s[(size_t)reflect101(y - kEdgeThreshold, lg.h) * lg.pitch + reflect101(x - kEdgeThreshold, lg.w)];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
void launch_pyramid(cudaStream_t st, uint8_t* pyr, size_t frame_stride, const LevelGeom* h_levels, int n_levels,
                    const LinCoef* d_coefs, int n_frames) {
    for (int l = 1; l < n_levels; ++l) {
        const LevelGeom& src = h_levels[l - 1];
        const LevelGeom& dst = h_levels[l];
        const double rx = (double)src.w / dst.w, ry = (double)src.h / dst.h;
        if (rx * (kRzTW - 1) + 18.0 <= kRzSrcVec * 16 - 1 && ry * (kRzTH - 1) + 3.0 <= kRzSrcRows) {
            dim3 grd((dst.w + kRzTW - 1) / kRzTW, (dst.h + kRzTH - 1) / kRzTH, n_frames);
            resize_level_kernel<<<grd, 256, 0, st>>>(pyr, frame_stride, src, dst, d_coefs + dst.tabx_off, d_coefs + dst.taby_off);
        } else {
            dim3 blk(32, 8), grd(((dst.w + 3) / 4 + 31) / 32, (dst.h + 7) / 8, n_frames);
            resize_level_gather_kernel<<<grd, blk, 0, st>>>(pyr, frame_stride, src, dst, d_coefs + dst.tabx_off, d_coefs + dst.taby_off);
        }
    }
}

void launch_fast(cudaStream_t st, const uint8_t* pyr, size_t frame_stride, const LevelGeom* d_levels,
                 const CellInfo* d_cells, int n_cells, int ini_th, int min_th, uint32_t* slots, int* counts,
                 int* overflow, int n_frames) {
    fast_cells_kernel<<<dim3(n_cells, n_frames), 256, 0, st>>>(pyr, frame_stride, d_levels, d_cells, n_cells, ini_th,
                                                             min_th, slots, counts, overflow);
}

void launch_compact(cudaStream_t st, const LevelGeom* d_levels, int n_levels, int n_cells, const uint32_t* slots,
                    const int* counts, int* cell_off, int* level_cnt, int* frame_total, uint32_t* dense, int dense_cap,
                    int* overflow, int n_frames) {
    cand_scan_kernel<<<n_frames, 256, 0, st>>>(counts, n_cells, d_levels, n_levels, cell_off, level_cnt, frame_total);
    const int gx = (n_cells + 63) / 64;
    cand_gather_kernel<<<dim3(gx, n_frames), 256, 0, st>>>(slots, counts, cell_off, frame_total, n_cells, dense,
                                                         dense_cap, overflow);
}

void launch_blur(cudaStream_t st, const uint8_t* pyr, uint8_t* blur, size_t frame_stride, const LevelGeom* h_levels,
                 int n_levels, int n_frames) {
    for (int l = 0; l < n_levels; ++l) {
        const LevelGeom& lg = h_levels[l];
        dim3 grd((lg.w + 127) / 128, (lg.h + 31) / 32, n_frames);
        blur_level_kernel<<<grd, 256, 0, st>>>(pyr, blur, frame_stride, lg);
    }
}

void launch_describe(cudaStream_t st, const uint8_t* pyr, const uint8_t* blur, size_t frame_stride,
                     const LevelGeom* d_levels, const SelKp* sel, const int* n_sel, int cap, int max_n,
                     const int umax[16], rgbl_keypoint* kps, uint8_t* desc, int n_frames) {
    if (max_n <= 0) return;
    UmaxTable t;
    for (int i = 0; i < 16; ++i) t.v[i] = umax[i];
    describe_kernel<<<dim3((max_n + 7) / 8, n_frames), 256, 0, st>>>(pyr, blur, frame_stride, d_levels, sel, n_sel, cap,
                                                                   t, kps, desc);
}

void launch_padded_level(cudaStream_t st, const uint8_t* pyr, size_t frame_stride, int frame, const LevelGeom& lg,
                         uint8_t* dst, int dst_pitch) {
    const int W = lg.w + 2 * kEdgeThreshold, H = lg.h + 2 * kEdgeThreshold;
    padded_level_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), dim3(32, 8), 0, st>>>(pyr, frame_stride, frame, lg, dst, dst_pitch);
}

}  // namespace rgbl
