// Resident tracking chain, TrackLocalMap half (src/Tracking.cc:2983-3050, 3377-3460): after TrackWithMotionModel's
// SearchByProjection(last frame) + PoseOptimization, the reference discards the outliers, projects the local map points that are not
// matched yet (Frame::isInFrustum, src/Frame.cc:602-664), searches them (ORBmatcher::SearchByProjection(F, vpMapPoints, th),
// src/ORBmatcher.cc:43-213) and runs PoseOptimization again on all the frame's map points.  The two single-CTA kernels below are the
// device-side glue between those reference functions (the functions themselves are match_kernels.cu / pose_kernels.cu):
//   tlm_prepare_kernel : outlier discard (src/Tracking.cc:2944-2966) -> isInFrustum over the local map -> ORDERED compaction of the
//                        visible points into the query arrays of the local search (the order of vpMapPoints decides ties);
//   tlm_edges_kernel   : PoseOptimization's edge list in keypoint order from (inliers of the first search) + (local matches)
//                        (src/Optimizer.cc:857-990), and the hand-over of the last frame's points into the local map
//                        (MapPoint::UpdateNormalAndDepth, src/MapPoint.cc:437-490, one observation).
// The local map of the harness is a ring of the K frames before the last one, each contributing its LiDAR-depth keypoints unprojected
// with the frame's final pose; slot = frame counter mod K, points inside a slot in keypoint order.
#include "rgbl_kernels.h"
#include "rgbl_device.cuh"

namespace rgbl {
namespace {

// block-wide ordered compaction helper: exclusive position of `flag` among the 1024 threads + running base
struct BlockScan {
    int* wsum; int* total;
    __device__ __forceinline__ int run(int flag, int tid, int& new_total) {
        const int lane = tid & 31, warp = tid >> 5;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = *total;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        int all = *total;
        for (int w = 0; w < 32; ++w) all += wsum[w];
        new_total = all;
        return base + incl - 1;           // position of this thread's element if flag == 1
    }
};

__global__ void __launch_bounds__(1024) tlm_prepare_kernel(FrameDev f, const float* __restrict__ pose, LocalRingDev ring, float cos_limit,
                                                           const int* __restrict__ n_edges, const int* __restrict__ e_idx,
                                                           const uint8_t* __restrict__ e_outlier, uint8_t* __restrict__ state,
                                                           int* __restrict__ match_last, LocalQueriesDev lq) {
    __shared__ float s_R[9], s_t[3], s_Ow[3];
    __shared__ int s_wsum[32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    // slot states for the local search: a slot is occupied exactly when it holds a map point of the first search that survived the
    // rotation check (match >= 0; the resolution kernel keeps its working states on chip) ...
    const int n_f = *f.n;
    for (int i = tid; i < n_f; i += 1024) state[i] = match_last[i] >= 0 ? 1 : 0;
    __syncthreads();
    // ... and was not an outlier of the first PoseOptimization: those slots are free again (mvpMapPoints[i] = NULL, src/Tracking.cc:2951-2953)
    const int ne = *n_edges;
    for (int e = tid; e < ne; e += 1024)
        if (e_outlier[e]) { const int i = e_idx[e]; state[i] = 0; match_last[i] = -1; }
    if (tid == 0) {
        // Frame::SetPose -> UpdatePoseMatrices (src/Frame.cc:562-569): mRcw = mTcw.rotationMatrix(), mtcw, mOw = Twc.translation()
        const float q[4] = {pose[0], pose[1], pose[2], pose[3]};
        float R[9]; quatf_to_matrix(q, R);
        float qinv[4], ow[3];
        se3f_inverse(pose, qinv, ow);
        for (int i = 0; i < 9; ++i) s_R[i] = R[i];
        for (int i = 0; i < 3; ++i) { s_t[i] = pose[4 + i]; s_Ow[i] = ow[i]; }
        s_total = 0;
    }
    __syncthreads();
    FrustumParams prm;
#pragma unroll
    for (int i = 0; i < 9; ++i) prm.Rcw[i] = s_R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { prm.tcw[i] = s_t[i]; prm.Ow[i] = s_Ow[i]; }
    prm.cos_limit = cos_limit;
    const int n_ring = ring.K * ring.cap;
    BlockScan scan{s_wsum, &s_total};
    for (int b = 0; b < n_ring; b += 1024) {
        const int p = b + tid;
        FrustumOut o{};
        int flag = 0;
        if (p < n_ring && ring.valid[p]) {
            const float P[3] = {ring.xw[3 * p], ring.xw[3 * p + 1], ring.xw[3 * p + 2]};
            o = frustum_point(f, prm, P, ring.normal + 3 * p, ring.mf_min[p], ring.mf_max[p]);
            flag = o.in_view;
        }
        int total = 0;
        const int pos = scan.run(flag, tid, total);
        if (flag && pos < lq.cap) {
            lq.in_view[pos] = 1; lq.obs_pos[pos] = 1;
            lq.proj_x[pos] = o.px; lq.proj_y[pos] = o.py; lq.proj_xr[pos] = o.pxr; lq.depth[pos] = o.depth; lq.level[pos] = o.level;
            lq.view_cos[pos] = o.view_cos; lq.src[pos] = p;
            const uint4* d = reinterpret_cast<const uint4*>(ring.desc + (size_t)p * 32);
            uint4* dd = reinterpret_cast<uint4*>(lq.desc + (size_t)pos * 32);
            dd[0] = d[0]; dd[1] = d[1];
        }
        __syncthreads();
        if (tid == 0) s_total = total;
        __syncthreads();
    }
    if (tid == 0) *lq.n = min(s_total, lq.cap);
}

__global__ void __launch_bounds__(1024) tlm_edges_kernel(FrameDev f, const int* __restrict__ match_last, const float* __restrict__ last_xw,
                                                         const int* __restrict__ match_local, const int* __restrict__ lq_src,
                                                         LocalRingDev ring, ChainEdgesOut eo, int* __restrict__ n_local_matches,
                                                         // hand-over of the last frame's points into the local map
                                                         int n_last_cap, const uint8_t* __restrict__ last_valid, const int* __restrict__ last_octave,
                                                         const uint8_t* __restrict__ last_desc, const float* __restrict__ last_pose) {
    __shared__ int s_wsum[32];
    __shared__ int s_total, s_nloc;
    __shared__ float s_Ow[3];
    const int tid = threadIdx.x;
    const int n_f = *f.n;
    if (tid == 0) {
        s_total = 0; s_nloc = 0;
        float qinv[4], ow[3];
        se3f_inverse(last_pose, qinv, ow);                 // KeyFrame::GetCameraCenter of the frame the points were created from
        s_Ow[0] = ow[0]; s_Ow[1] = ow[1]; s_Ow[2] = ow[2];
    }
    __syncthreads();
    BlockScan scan{s_wsum, &s_total};
    int nloc = 0;
    for (int b = 0; b < n_f; b += 1024) {
        const int i = b + tid;
        const int ma = (i < n_f) ? match_last[i] : -1;
        const int mb = (i < n_f && ma < 0) ? match_local[i] : -1;
        const int flag = (ma >= 0 || mb >= 0) ? 1 : 0;
        if (mb >= 0) ++nloc;
        int total = 0;
        const int e = scan.run(flag, tid, total);
        if (flag) {
            const float* x = (ma >= 0) ? (last_xw + 3 * (size_t)ma) : (ring.xw + 3 * (size_t)lq_src[mb]);
            const rgbl_keypoint kp = f.keys[i];
            eo.exw[3 * e] = x[0]; eo.exw[3 * e + 1] = x[1]; eo.exw[3 * e + 2] = x[2];
            const float ur = f.uright[i];
            eo.eobs[3 * e] = kp.x; eo.eobs[3 * e + 1] = kp.y; eo.eobs[3 * e + 2] = ur;
            const float sc = f.scale[kp.octave];
            eo.einfo[e] = __fdiv_rn(1.0f, __fmul_rn(sc, sc));                  // mvInvLevelSigma2 (src/ORBextractor.cc:421-429)
            eo.est[e] = ur >= 0.f;
            eo.eidx[e] = i;
        }
        __syncthreads();
        if (tid == 0) s_total = total;
        __syncthreads();
    }
    if (nloc) atomicAdd(&s_nloc, nloc);
    // the last frame's points become local map points (ring slot = frames inserted so far mod K); the search above has already read the ring
    const int slot = (*ring.count) % ring.K;
    const size_t base = (size_t)slot * ring.cap;
    for (int j = tid; j < ring.cap; j += 1024) {
        uint8_t v = 0;
        if (j < n_last_cap && last_valid[j]) {
            const float P[3] = {last_xw[3 * j], last_xw[3 * j + 1], last_xw[3 * j + 2]};
            const float PC[3] = {__fsub_rn(P[0], s_Ow[0]), __fsub_rn(P[1], s_Ow[1]), __fsub_rn(P[2], s_Ow[2])};
            const float dist = sqrtf(eig_sum3(__fmul_rn(PC[0], PC[0]), __fmul_rn(PC[1], PC[1]), __fmul_rn(PC[2], PC[2])));
            const size_t p = base + j;
            ring.xw[3 * p] = P[0]; ring.xw[3 * p + 1] = P[1]; ring.xw[3 * p + 2] = P[2];
            ring.normal[3 * p] = __fdiv_rn(PC[0], dist); ring.normal[3 * p + 1] = __fdiv_rn(PC[1], dist); ring.normal[3 * p + 2] = __fdiv_rn(PC[2], dist);
            const float mx = __fmul_rn(dist, f.scale[last_octave[j]]);
            ring.mf_max[p] = mx;
            ring.mf_min[p] = __fdiv_rn(mx, f.scale[f.n_levels - 1]);
            const uint4* d = reinterpret_cast<const uint4*>(last_desc + (size_t)j * 32);
            uint4* dd = reinterpret_cast<uint4*>(ring.desc + p * 32);
            dd[0] = d[0]; dd[1] = d[1];
            v = 1;
        }
        ring.valid[base + j] = v;
    }
    __syncthreads();
    if (tid == 0) { *eo.n_edges = s_total; *n_local_matches = s_nloc; *ring.count = *ring.count + 1; }
}

}  // namespace

void launch_tlm_prepare(cudaStream_t st, const FrameDev& f, const float* pose, const LocalRingDev& ring, float cos_limit, const int* n_edges,
                        const int* e_idx, const uint8_t* e_outlier, uint8_t* state, int* match_last, const LocalQueriesDev& lq) {
    tlm_prepare_kernel<<<1, 1024, 0, st>>>(f, pose, ring, cos_limit, n_edges, e_idx, e_outlier, state, match_last, lq);
}

void launch_tlm_edges(cudaStream_t st, const FrameDev& f, const int* match_last, const float* last_xw, const int* match_local, const int* lq_src,
                      const LocalRingDev& ring, const ChainEdgesOut& eo, int* n_local_matches, int n_last_cap, const uint8_t* last_valid,
                      const int* last_octave, const uint8_t* last_desc, const float* last_pose) {
    tlm_edges_kernel<<<1, 1024, 0, st>>>(f, match_last, last_xw, match_local, lq_src, ring, eo, n_local_matches, n_last_cap, last_valid,
                                         last_octave, last_desc, last_pose);
}

}  // namespace rgbl
