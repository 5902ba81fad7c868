// Fused per-level tile kernel for sm_100a: ONE read of a pyramid level produces
//   (a) the level's 7x7 Gaussian pre-filter  (cv::GaussianBlur(7x7, 2, 2, BORDER_REFLECT_101), src/ORBextractor.cc:1132-1133), and
//   (b) the next pyramid level               (cv::resize(level l-1, INTER_LINEAR),              src/ORBextractor.cc:1170-1195),
// instead of the two passes of orb_kernels.cu (resize_level_kernel reads level l once, blur_level_kernel reads it again).
//
// Data movement: the tile and its halo come in by TMA.  Every level of the pyramid buffer is described by one 3-D tensor map
// (x, y, frame slot; uint8; row pitch % 64 == 0, level offset % 256 == 0, see build_geometry): a CTA asks for the box
// [x0-16, x0+144) x [y0-3, y0+45) x {frame} with `cp.async.bulk.tensor.3d...mbarrier::complete_tx::bytes`; out-of-range bytes arrive as 0
// and the few REFLECT_101 halo bytes of edge tiles are patched in shared memory.  CTAs are persistent (grid = SM count x resident
// CTAs) and double-buffered: the box of tile i+1 is in flight while tile i is computed, one mbarrier per stage.
//
// Arithmetic: the blur is the exact integer form of SURVEY A.2 (8.8 fixed-point kernel {18,34,48,56,48,34,18}, both passes summed
// exactly, one rounding at the end), so the order of the passes and the instruction mix are free:
//   horizontal pass on bytes      : 2 x IDP4A per output   (byte windows aligned with funnel shifts),
//   vertical pass on 16-bit sums  : 4 x IDP2A per output   (the horizontal pass stores rows (2p, 2p+1) of a column as one 16x2 word),
//   rounding                      : accumulator starts at 32768, the result is byte 2 of the sum (PRMT).
// About 10 integer instructions per pixel instead of ~50 for the scalar form of blur_level_kernel.
// The resize is the scalar form of resize_level_kernel reading the same shared-memory tile: a CTA owns the destination
// 4-pixel groups whose first source column lies in its tile and the destination rows whose first source row lies in its tile.
#include <cuda.h>

#include "rgbl_device.cuh"
#include "rgbl_kernels.h"

namespace rgbl {
namespace {

constexpr int kTW = 128, kTH = 42;                 // output tile (level 0 of a 376-row KITTI frame: 9 tile rows exactly)
constexpr int kHaloX = 16;                         // the box origin must be 16-byte aligned in x (measured: x0-4 is an illegal instruction, tests/tools/tma_probe.cu)
constexpr int kBoxW = kTW + 2 * kHaloX, kBoxH = 48; // TMA box: x0-16 .. x0+143, y0-3 .. y0+44
constexpr int kBoxWords = kBoxW / 4;
constexpr int kStageBytes = kBoxW * kBoxH;         // 7680 = 60 x 128
constexpr int kPairs = kBoxH / 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}

// smallest i in [0, n] with i == n or tab[step * i].s >= x   (the tables are monotone)
__device__ __forceinline__ int first_at_least(const LinCoef* __restrict__ tab, int n, int step, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tab[step * mid].s >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <bool kResize>
__global__ void __launch_bounds__(256) level_tile_kernel(const __grid_constant__ CUtensorMap tm, uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
                                                         size_t frame_stride, LevelGeom lg, LevelGeom dst, const LinCoef* __restrict__ tabx,
                                                         const LinCoef* __restrict__ taby, int tiles_x, int tiles_y, int n_tiles) {
    __shared__ alignas(128) uint8_t s_in[2][kStageBytes];
    __shared__ alignas(16) uint32_t s_hb[kPairs][kTW];          // horizontal sums, rows (2p, 2p+1) packed 16x2
    __shared__ alignas(8) uint64_t s_bar[2];
    __shared__ int s_rz[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_frame = tiles_x * tiles_y;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the barriers are visible to the async proxy
    }
    __syncthreads();

    auto issue = [&](int t, int stage) {                       // one thread: box of tile t -> stage
        const int f = t / per_frame, r = t - f * per_frame, ty = r / tiles_x, tx = r - ty * tiles_x;
        const uint32_t bar = smem_u32(&s_bar[stage]), dstp = smem_u32(&s_in[stage][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kStageBytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(dstp), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(tx * kTW - kHaloX), "r"(ty * kTH - 3), "r"(f), "r"(bar) : "memory");
    };

    int t = blockIdx.x;
    if (tid == 0 && t < n_tiles) issue(t, 0);
    for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
        const int stage = it & 1;
        if (tid == 0 && t + (int)gridDim.x < n_tiles) issue(t + gridDim.x, stage ^ 1);     // stage^1 was released by the barrier ending iteration it-1
        const int f = t / per_frame, r = t - f * per_frame, ty = r / tiles_x, tx = r - ty * tiles_x;
        const int x0 = tx * kTW, y0 = ty * kTH;
        if (kResize && tid < 4) {
            // destination 4-pixel groups [s_rz[0], s_rz[1]) and rows [s_rz[2], s_rz[3]) owned by this tile
            const int ng = (dst.w + 3) >> 2;
            s_rz[tid] = (tid < 2) ? first_at_least(tabx, ng, 4, x0 + (tid ? kTW : 0)) : first_at_least(taby, dst.h, 1, y0 + (tid == 3 ? kTH : 0));
        }
        mbar_wait(smem_u32(&s_bar[stage]), (it >> 1) & 1);
        uint8_t* in = s_in[stage];
        uint32_t* in32 = reinterpret_cast<uint32_t*>(in);

        // ---- BORDER_REFLECT_101 halo of edge tiles (the box arrives zero-filled outside the level) ----
        const bool left = x0 == 0, right = x0 + kTW + 3 > lg.w, top = y0 == 0, bottom = y0 + kTH + 3 > lg.h;
        if (left || right || top || bottom) {
            if ((left || right) && tid < kBoxH) {
                uint8_t* row = in + tid * kBoxW;
                if (left) { row[kHaloX - 1] = row[kHaloX + 1]; row[kHaloX - 2] = row[kHaloX + 2]; row[kHaloX - 3] = row[kHaloX + 3]; }
                if (right) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int c = lg.w + k - x0 + kHaloX;
                        if (c < kBoxW) row[c] = row[c - 2 - 2 * k];                    // x = w + k  <-  w - 2 - k
                    }
                }
            }
            __syncthreads();
            if (top)
                for (int i = tid; i < 3 * kBoxWords; i += 256) { const int k = i / kBoxWords, w = i - k * kBoxWords; in32[k * kBoxWords + w] = in32[(6 - k) * kBoxWords + w]; }
            if (bottom)
                for (int i = tid; i < 3 * kBoxWords; i += 256) {
                    const int k = i / kBoxWords, w = i - k * kBoxWords;
                    const int rb = lg.h + k - y0 + 3;                                  // y = h + k  <-  h - 2 - k
                    if (rb < kBoxH) in32[rb * kBoxWords + w] = in32[(rb - 2 - 2 * k) * kBoxWords + w];
                }
            __syncthreads();
        }

        // ---- horizontal pass: rows (2p, 2p+1) x 4 columns per item, two IDP4A per output ----
#pragma unroll
        for (int k = 0; k < kPairs / 8; ++k) {
            const int p = warp + 8 * k, g = lane;
            uint32_t h[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint32_t* row = in32 + (2 * p + rr) * kBoxWords + g + (kHaloX / 4 - 1);
                const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // output column 4g + j = box byte 16 + 4g + j = byte 4 + j of (w0, w1, w2); taps = bytes 1+j .. 7+j
                    const uint32_t a = (j == 3) ? w1 : __funnelshift_r(w0, w1, 8 * (j + 1));
                    const uint32_t b = (j == 3) ? w2 : __funnelshift_r(w1, w2, 8 * (j + 1));
                    h[rr][j] = __dp4a(b, 0x00122230u, __dp4a(a, 0x38302212u, 0u));      // (18,34,48,56 | 48,34,18,0)
                }
            }
            uint4 o;
            o.x = h[0][0] | (h[1][0] << 16); o.y = h[0][1] | (h[1][1] << 16); o.z = h[0][2] | (h[1][2] << 16); o.w = h[0][3] | (h[1][3] << 16);
            *reinterpret_cast<uint4*>(&s_hb[p][4 * g]) = o;
        }
        __syncthreads();

        // ---- vertical pass: 6 rows x 4 columns per thread, four IDP2A per output ----
        if (warp < kTH / 6) {
            const int g = lane, gx = x0 + 4 * g;
            if (gx < lg.w) {
                uint4 P[6];
#pragma unroll
                for (int q = 0; q < 6; ++q) P[q] = *reinterpret_cast<const uint4*>(&s_hb[3 * warp + q][4 * g]);
                uint8_t* d = blur + (size_t)f * frame_stride + lg.off + gx;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int gy = y0 + 6 * warp + j;
                    if (gy >= lg.h) break;
                    const int q = j >> 1;
                    // even output row: box rows 2q .. 2q+6   -> pairs q..q+3 weighted (18,34)(48,56)(48,34)(18,0)
                    // odd output row : box rows 2q+1 .. 2q+7 -> pairs q..q+3 weighted (0,18)(34,48)(56,48)(34,18)
                    const uint32_t B1 = (j & 1) ? 0x30221200u : 0x38302212u, B2 = (j & 1) ? 0x12223038u : 0x00122230u;
                    uint32_t acc[4];
#define RGBL_VCOL(c, m) acc[c] = __dp2a_hi(P[q + 3].m, B2, __dp2a_lo(P[q + 2].m, B2, __dp2a_hi(P[q + 1].m, B1, __dp2a_lo(P[q].m, B1, 32768u))))
                    RGBL_VCOL(0, x); RGBL_VCOL(1, y); RGBL_VCOL(2, z); RGBL_VCOL(3, w);
#undef RGBL_VCOL
                    const uint32_t lo = __byte_perm(acc[0], acc[1], 0x0062), hi = __byte_perm(acc[2], acc[3], 0x0062);
                    *reinterpret_cast<uint32_t*>(d + (size_t)gy * lg.pitch) = __byte_perm(lo, hi, 0x5410);
                }
            }
        }

        // ---- the next level's pixels whose bilinear footprint starts in this tile ----
        // lane = destination 4-pixel group (its four column coefficients stay in registers), warp = destination row phase.
        // Per pixel: the source pair (p[s], p[s+1]) of each of the two rows is one funnel shift of two box words, the horizontal
        // interpolation one IDP2A with the 16x2 weight pair.  s+1 (and row s+1) may lie outside the level only where its weight is 0
        // (linear_coefs clamps s to n-1 with f = 0), and the box holds finite bytes there.
        if (kResize) {
            const int g0 = s_rz[0], n_g = s_rz[1] - g0, d0 = s_rz[2], n_dy = s_rz[3] - d0;
            uint8_t* d = pyr + (size_t)f * frame_stride + dst.off;
            for (int gb = lane; gb < n_g; gb += 32) {                  // one pass at ratio 1.2 (a tile owns <= 28 groups)
                const int x4 = 4 * (g0 + gb);
                int widx[4]; uint32_t sh[4], cpair[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const LinCoef cx = tabx[min(x4 + k, dst.w - 1)];
                    const int o = cx.s - x0 + kHaloX;
                    widx[k] = o >> 2; sh[k] = 8 * (o & 3);
                    cpair[k] = (x4 + k < dst.w) ? ((uint32_t)(uint16_t)cx.c0 | ((uint32_t)(uint16_t)cx.c1 << 16)) : 0u;
                }
                for (int ry = warp; ry < n_dy; ry += 8) {
                    const int y = d0 + ry;
                    const LinCoef cy = taby[y];
                    const uint32_t* r0 = in32 + (cy.s - y0 + 3) * kBoxWords;
                    const uint32_t* r1 = r0 + kBoxWords;
                    const int b0 = cy.c0, b1 = cy.c1;
                    uint32_t out = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t a0 = __funnelshift_r(r0[widx[k]], r0[widx[k] + 1], sh[k]);
                        const uint32_t a1 = __funnelshift_r(r1[widx[k]], r1[widx[k] + 1], sh[k]);
                        const int h0 = (int)__dp2a_lo(cpair[k], a0, 0u), h1 = (int)__dp2a_lo(cpair[k], a1, 0u);
                        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                        out |= (uint32_t)(v & 0xff) << (8 * k);
                    }
                    *reinterpret_cast<uint32_t*>(d + (size_t)y * dst.pitch + x4) = out;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy accesses to this stage are ordered before the next box written into it
        __syncthreads();
    }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn tensor_map_encoder() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}

}  // namespace

static_assert(sizeof(CUtensorMap) == sizeof(((LevelTensorMaps*)nullptr)->map[0]), "LevelTensorMaps holds raw CUtensorMap objects");

// One 3-D uint8 tensor map per pyramid level over the context's pyramid buffer: (w, h, n_slots), strides (pitch, frame_stride).
int make_level_tensor_maps(uint8_t* pyr, size_t frame_stride, int n_slots, const LevelGeom* levels, int n_levels, LevelTensorMaps* out) {
    out->n_levels = 0;
    EncodeFn enc = tensor_map_encoder();
    if (!enc) return -1;
    if (frame_stride % 16 || n_levels > RGBL_MAX_LEVELS) return -2;
    for (int l = 0; l < n_levels; ++l) {
        const LevelGeom& g = levels[l];
        if (g.pitch % 16 || g.off % 16 || g.w < 8 || g.h < 8) return -2;
        const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)n_slots};
        const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)frame_stride};
        const cuuint32_t box[3] = {kBoxW, kBoxH, 1}, estr[3] = {1, 1, 1};
        const CUresult r = enc(reinterpret_cast<CUtensorMap*>(out->map[l]), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr + g.off, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return -3;
    }
    out->n_levels = n_levels;
    return 0;
}

// Pyramid + pre-filter of every level for `n_frames` frame slots: level l's launch writes blur(l) and level l+1.
// Returns 0, or -1 when a level ratio does not fit the tile's halo (the caller then uses launch_pyramid + launch_blur).
int launch_level_tiles(cudaStream_t st, const LevelTensorMaps& tms, uint8_t* pyr, uint8_t* blur, size_t frame_stride, const LevelGeom* h_levels,
                       int n_levels, const LinCoef* d_coefs, int n_frames) {
    if (tms.n_levels != n_levels) return -1;
    for (int l = 1; l < n_levels; ++l) {
        // group span: first source column of a 4-group + 3 * ratio + 1 must stay inside the 15-byte right halo; rows need 1 of the 3 halo rows
        if ((double)h_levels[l - 1].w / h_levels[l].w > 3.0) return -1;
    }
    static int resident = 0;
    static int sms = 0;
    if (!resident) {
        int dev = 0, a = 0, b = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, level_tile_kernel<true>, 256, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, level_tile_kernel<false>, 256, 0);
        resident = a < b ? a : b;
        if (resident < 1) resident = 1;
    }
    for (int l = 0; l < n_levels; ++l) {
        const LevelGeom& lg = h_levels[l];
        const int tiles_x = (lg.w + kTW - 1) / kTW, tiles_y = (lg.h + kTH - 1) / kTH, n_tiles = tiles_x * tiles_y * n_frames;
        const int grid = n_tiles < sms * resident ? n_tiles : sms * resident;
        const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tms.map[l]);
        if (l + 1 < n_levels) {
            const LevelGeom& dst = h_levels[l + 1];
            level_tile_kernel<true><<<grid, 256, 0, st>>>(tm, pyr, blur, frame_stride, lg, dst, d_coefs + dst.tabx_off, d_coefs + dst.taby_off, tiles_x,
                                                         tiles_y, n_tiles);
        } else {
            level_tile_kernel<false><<<grid, 256, 0, st>>>(tm, pyr, blur, frame_stride, lg, lg, nullptr, nullptr, tiles_x, tiles_y, n_tiles);
        }
    }
    return 0;
}

}  // namespace rgbl
