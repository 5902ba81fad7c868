// TEST INFRASTRUCTURE (emulated library only): pose_kernels.cu is a thread-block-cluster kernel (DSMEM st.async + mbarrier) that the
// CUDA-on-CPU shim does not model, so the emulated library refuses PoseOptimization loudly instead of computing it some other way.
#include <cstdio>
#include <cstdlib>
#include "rgbl_kernels.h"

namespace rgbl {
void launch_pose_optimize(cudaStream_t, const PoseProblemDev&, double*, uint8_t*, uint8_t*, float*, int*, const ChainPrepDev*) {
    std::fprintf(stderr, "librgbl_b200_emu: PoseOptimization (pose_kernels.cu, 4-CTA cluster) is not emulated; run it on the GPU\n");
    std::abort();
}
}  // namespace rgbl
