// Kernel launchers (defined in the .cu files) and the device-side record types.
#ifndef RGBL_KERNELS_H
#define RGBL_KERNELS_H

#include <cuda_runtime.h>

#include "rgbl_internal.h"

namespace rgbl {

// Selected keypoint handed from the quad-tree to the describe kernel (level coordinates, +16 applied).
struct __align__(8) SelKp {
    uint16_t x, y;
    uint8_t level, score;
    uint16_t pad;
};

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t pack_cand_dev(int x, int y, int s) {
    return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
}
#endif

// orb_kernels.cu
void launch_pyramid(cudaStream_t st, uint8_t* pyr, size_t frame_stride, const LevelGeom* h_levels, int n_levels,
                    const LinCoef* d_coefs, int n_frames);
void launch_fast(cudaStream_t st, const uint8_t* pyr, size_t frame_stride, const LevelGeom* d_levels,
                 const CellInfo* d_cells, int n_cells, int ini_th, int min_th, uint32_t* slots, int* counts,
                 int* overflow, int n_frames);
void launch_compact(cudaStream_t st, const LevelGeom* d_levels, int n_levels, int n_cells, const uint32_t* slots,
                    const int* counts, int* cell_off, int* level_cnt, int* frame_total, uint32_t* dense, int dense_cap,
                    int* overflow, int n_frames);
void launch_blur(cudaStream_t st, const uint8_t* pyr, uint8_t* blur, size_t frame_stride, const LevelGeom* h_levels,
                 int n_levels, int n_frames);
void launch_describe(cudaStream_t st, const uint8_t* pyr, const uint8_t* blur, size_t frame_stride,
                     const LevelGeom* d_levels, const SelKp* sel, const int* n_sel, int cap, int max_n,
                     const int umax[16], rgbl_keypoint* kps, uint8_t* desc, int n_frames);
void launch_padded_level(cudaStream_t st, const uint8_t* pyr, size_t frame_stride, int frame, const LevelGeom& lg,
                         uint8_t* dst, int dst_pitch);

// depth_kernels.cu
struct DepthDev {
    float P[12];
    float min_dist, max_dist, bf, inv_scale_m;   // inv_scale_m = max_dist * ScaleFactor (the inversion constant M)
    int ku, kv;
    uint8_t mask[81];
};
void launch_depth_project(cudaStream_t st, const float* pts, int pts_stride, const int* n_pts, int max_n_pts,
                          const DepthDev& prm, int W, int H, uint32_t* idx_map, uint32_t stamp, int n_frames);
void launch_depth_resolve_dilate(cudaStream_t st, const float* pts, int pts_stride, const int* n_pts,
                                 const DepthDev& prm, int W, int H, const uint32_t* idx_map, uint32_t stamp,
                                 float* raw, float* processed, int n_frames);
void launch_depth_gather(cudaStream_t st, const float* processed, int W, int H, const rgbl_keypoint* kps,
                         const rgbl_keypoint* kps_un, const int* n_kp, int cap, int max_n, float bf, float* depth,
                         float* uright, int n_frames);

}  // namespace rgbl
#endif
