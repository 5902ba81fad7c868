/* Plain C99 client of the C ABI (include/rgbl_b200.h): the calls a non-C++ host (cgo / JNI / ctypes) would bind.
 * Build:  gcc -std=c99 -Iinclude examples/abi_demo.c -Lorb_slam3_rgbl_b200 -lrgbl_b200 -Wl,-rpath,$PWD/orb_slam3_rgbl_b200 -o abi_demo
 * Runs one RGB-L frame (synthetic gradient image + a small planar cloud) through frame construction when a CUDA device is
 * present; without one rgbl_create fails with RGBL_E_CUDA - there is no CPU fallback - and the program reports that.          */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rgbl_b200.h"

int main(void) {
    const int W = 640, H = 376, N = 20000;
    rgbl_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.width = W; cfg.height = H; cfg.max_batch = 1; cfg.max_points = N; cfg.device = 0;
    cfg.orb.nfeatures = 1000; cfg.orb.scale_factor = 1.2f; cfg.orb.nlevels = 8; cfg.orb.ini_th_fast = 12; cfg.orb.min_th_fast = 7;
    rgbl_ctx* ctx = NULL;
    int rc = rgbl_create(&cfg, &ctx);
    printf("abi version %d\n", rgbl_abi_version());
    if (rc != RGBL_OK) {
        printf("rgbl_create failed: %d (%s)\n", rc, rgbl_last_error(NULL));
        return rc == RGBL_E_CUDA ? 3 : 1;          /* 3 = no device: expected on a CPU-only box */
    }
    uint8_t* img = (uint8_t*)malloc((size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) img[(size_t)y * W + x] = (uint8_t)(((x / 16) ^ (y / 16)) & 1 ? 200 - (x & 15) * 3 : 40 + (y & 15) * 5);
    float* pts = (float*)malloc(sizeof(float) * 4 * N);          /* planar rows x | y | z | 1 */
    for (int i = 0; i < N; ++i) {
        pts[i] = -10.f + 20.f * (float)(i % 200) / 200.f; pts[N + i] = -1.f + 2.f * (float)(i / 200) / 100.f; pts[2 * N + i] = 12.f; pts[3 * N + i] = 1.f;
    }
    const float P[12] = {718.856f, 0, 320.f, 0, 0, 718.856f, 188.f, 0, 0, 0, 1, 0};
    rgbl_depth_params dp;
    memset(&dp, 0, sizeof(dp));
    dp.method = RGBL_DEPTH_INVERSE_DILATION; dp.min_dist = 5.f; dp.max_dist = 200.f; dp.bf = 387.f; dp.inv_dilation_scale = 1.f; dp.ku = 5; dp.kv = 5;
    rgbl_depth_structuring_element("Diamond", 5, 5, dp.mask);
    const int cap = rgbl_keypoint_capacity(ctx);
    rgbl_keypoint* kps = (rgbl_keypoint*)malloc(sizeof(rgbl_keypoint) * cap);
    uint8_t* desc = (uint8_t*)malloc((size_t)cap * 32);
    float* depth = (float*)malloc(sizeof(float) * cap); float* uright = (float*)malloc(sizeof(float) * cap);
    const uint8_t* imgs[1] = {img}; const float* clouds[1] = {pts}; const int npts[1] = {N}; int n_out[1] = {0};
    rc = rgbl_frame_rgbl_batch(ctx, 1, imgs, W, H, W, clouds, npts, P, &dp, kps, desc, depth, uright, cap, n_out);
    if (rc != RGBL_OK) { printf("rgbl_frame_rgbl_batch failed: %d (%s)\n", rc, rgbl_last_error(ctx)); rgbl_destroy(ctx); return 1; }
    int with_depth = 0;
    for (int i = 0; i < n_out[0]; ++i) with_depth += depth[i] > 0;
    printf("%d keypoints, %d with LiDAR depth\n", n_out[0], with_depth);
    free(img); free(pts); free(kps); free(desc); free(depth); free(uright);
    rgbl_destroy(ctx);
    return 0;
}
