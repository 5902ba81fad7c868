// Debug probe (GPU box): which cp.async.bulk.tensor.3d box origins does the TMA unit accept for a uint8 tensor?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
constexpr int BW = 144, BH = 48;
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint8_t* out) {
    __shared__ alignas(128) uint8_t buf[BW * BH];
    __shared__ alignas(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
    }
    uint32_t done;
    do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(0) : "memory"); } while (!done);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    auto enc = (CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    const int W = 1241, H = 376, P = 1280, F = 2; const size_t FS = (size_t)P * H;
    std::vector<uint8_t> h(FS * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / P);
    uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, BW * BH); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, F}, str[2] = {P, FS}; cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int xs[] = {0, 16, 128, -16, 124, -4, 764, 1, 1200}, ys[] = {0, -3, 39, 350};
    std::vector<uint8_t> got(BW * BH);
    for (int x : xs) for (int y : ys) for (int z = 0; z < 2; ++z) {
        probe<<<1, 128>>>(tm, x, y, z, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("x=%d y=%d z=%d: %s\n", x, y, z, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
            const int gx = x + c, gy = y + r2;
            const uint8_t want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)z * FS + (size_t)gy * P + gx] : 0;
            bad += got[r2 * BW + c] != want;
        }
        printf("x=%d y=%d z=%d: ok, mismatches %ld\n", x, y, z, bad);
    }
    return 0;
}
