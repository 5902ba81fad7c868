// Frame::ComputeBoW (src/Frame.cc:828-835) = DBoW2 TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1206 and :1218-1259) for the ORB vocabulary (TF_IDF or TF weighting,
// L1 scoring), on a vocabulary flattened into device arrays.
//   bow_descend_kernel : one warp per descriptor walks the tree; the children of a node are scored by the lanes
//                        (FORB::distance = 256-bit Hamming, FORB.cpp:81-101) and the FIRST minimum wins, as the reference's
//                        strict "d < best_d" scan does.
//   bow_assemble_kernel: one CTA builds the two std::map results: a bitonic sort of (word | feature) and (node | feature)
//                        keys gives the maps' ascending-key order with features in insertion order inside a key; weights of
//                        equal words are added in feature order and the L1 norm is accumulated in ascending word order by
//                        one thread — the same sequence of double additions as BowVector::addWeight / normalize, so the
//                        values are bit-identical, not just close.
#include "rgbl_device.cuh"
#include "rgbl_kernels.h"

namespace rgbl {

__global__ void __launch_bounds__(256) bow_descend_kernel(VocabDev voc, int n, const uint8_t* __restrict__ desc, int nid_level,
                                                          int* __restrict__ f_word, double* __restrict__ f_weight,
                                                          int* __restrict__ f_node) {
    const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32)), q1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)f * 32) + 1);
    int node = 0, level = 0, nid = 0;
    int cb = voc.child_begin[0], ce = voc.child_begin[1];
    while (ce > cb) {                                   // !isLeaf()
        ++level;
        unsigned best = 0xffffffffu;                    // dist << 20 | position in the children list: min = first minimum
        for (int c0 = cb; c0 < ce; c0 += 32) {
            const int c = c0 + lane;
            if (c < ce) {
                const int id = voc.child_index[c];
                const uint4* p = reinterpret_cast<const uint4*>(voc.node_desc + (size_t)id * 32);
                const uint4 b0 = __ldg(p), b1 = __ldg(p + 1);
                const int d = __popc(q0.x ^ b0.x) + __popc(q0.y ^ b0.y) + __popc(q0.z ^ b0.z) + __popc(q0.w ^ b0.w) +
                              __popc(q1.x ^ b1.x) + __popc(q1.y ^ b1.y) + __popc(q1.z ^ b1.z) + __popc(q1.w ^ b1.w);
                best = min(best, ((unsigned)d << 20) | (unsigned)(c - cb));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        node = voc.child_index[cb + (int)(best & 0xfffffu)];
        if (level == nid_level) nid = node;
        cb = voc.child_begin[node]; ce = voc.child_begin[node + 1];
    }
    if (lane == 0) { f_word[f] = voc.word_id[node]; f_weight[f] = voc.node_weight[node]; f_node[f] = nid; }
}

namespace {

// in-place bitonic sort of n_pow2 64-bit keys in shared memory by the whole CTA
__device__ void bitonic_sort(unsigned long long* k, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n_pow2 / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
        }
    __syncthreads();
}

}  // namespace

// dynamic shared memory: 2 * n_pow2 keys
__global__ void __launch_bounds__(1024) bow_assemble_kernel(int n, int n_pow2, const int* __restrict__ f_word,
                                                            const double* __restrict__ f_weight, const int* __restrict__ f_node,
                                                            int* __restrict__ bow_word, double* __restrict__ bow_value,
                                                            int* __restrict__ fv_node, int* __restrict__ fv_start,
                                                            int* __restrict__ fv_feature, int* __restrict__ counts /* n_words, n_fv_nodes */,
                                                            int* __restrict__ scratch /* n_pow2 + 1 ints */) {
    extern __shared__ unsigned long long keys[];
    unsigned long long* kw = keys;
    unsigned long long* kn = keys + n_pow2;
    __shared__ int s_cnt[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < n_pow2; i += 1024) {
        unsigned long long a = ~0ull, b = ~0ull;
        if (i < n && f_weight[i] > 0) {                 // "not stopped" words only
            a = ((unsigned long long)(unsigned)f_word[i] << 32) | (unsigned)i;
            b = ((unsigned long long)(unsigned)f_node[i] << 32) | (unsigned)i;
        }
        kw[i] = a; kn[i] = b;
    }
    if (tid < 2) s_cnt[tid] = 0;
    bitonic_sort(kw, n_pow2);
    bitonic_sort(kn, n_pow2);
    // ---- BowVector: one thread per word adds its weights in feature order; dense word slots through a flag scan in `scratch`
    // (serial scan by one warp is avoided: heads are counted with an atomic per head and ranked by a second pass below)
    for (int i = tid; i < n_pow2; i += 1024) {
        const bool valid = kw[i] != ~0ull;
        const bool head = valid && (i == 0 || (kw[i - 1] >> 32) != (kw[i] >> 32));
        scratch[i] = head ? 1 : 0;
    }
    __syncthreads();
    // exclusive scan of the head flags (n_pow2 <= 8192): each thread owns a contiguous chunk
    __shared__ int s_part[1024];
    {
        const int per = (n_pow2 + 1023) / 1024, b = tid * per, e = min(n_pow2, b + per);
        int s = 0;
        for (int i = b; i < e; ++i) s += scratch[i];
        s_part[tid] = s;
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int t = 0; t < 1024; ++t) { const int v = s_part[t]; s_part[t] = acc; acc += v; } s_cnt[0] = acc; }
        __syncthreads();
        int acc = s_part[tid];
        for (int i = b; i < e; ++i) { const int v = scratch[i]; scratch[i] = acc; acc += v; }
    }
    __syncthreads();
    for (int i = tid; i < n_pow2; i += 1024) {
        if (kw[i] == ~0ull) continue;
        const unsigned w = (unsigned)(kw[i] >> 32);
        if (i != 0 && (unsigned)(kw[i - 1] >> 32) == w) continue;         // not a head
        double sum = f_weight[(unsigned)kw[i]];
        for (int j = i + 1; j < n_pow2 && (unsigned)(kw[j] >> 32) == w && kw[j] != ~0ull; ++j) sum += f_weight[(unsigned)kw[j]];
        const int slot = scratch[i];
        bow_word[slot] = (int)w; bow_value[slot] = sum;
    }
    __syncthreads();
    __shared__ double s_norm;
    if (tid == 0) {                                       // BowVector::normalize(L1): ascending word order, one accumulator
        double norm = 0.0;
        const int nw = s_cnt[0];
        for (int k = 0; k < nw; ++k) norm += fabs(bow_value[k]);
        s_norm = norm;
    }
    __syncthreads();
    if (s_norm > 0.0) for (int k = tid; k < s_cnt[0]; k += 1024) bow_value[k] = bow_value[k] / s_norm;
    __syncthreads();
    // ---- FeatureVector: heads of equal node ids; features already in insertion order inside a node
    for (int i = tid; i < n_pow2; i += 1024) {
        const bool valid = kn[i] != ~0ull;
        const bool head = valid && (i == 0 || (kn[i - 1] >> 32) != (kn[i] >> 32));
        scratch[i] = head ? 1 : 0;
        if (valid) fv_feature[i] = (int)(unsigned)kn[i];
    }
    __syncthreads();
    {
        const int per = (n_pow2 + 1023) / 1024, b = tid * per, e = min(n_pow2, b + per);
        int s = 0;
        for (int i = b; i < e; ++i) s += scratch[i];
        s_part[tid] = s;
        __syncthreads();
        if (tid == 0) { int acc = 0; for (int t = 0; t < 1024; ++t) { const int v = s_part[t]; s_part[t] = acc; acc += v; } s_cnt[1] = acc; }
        __syncthreads();
        int acc = s_part[tid];
        for (int i = b; i < e; ++i) {
            const int v = scratch[i];
            if (v) { fv_node[acc] = (int)(unsigned)(kn[i] >> 32); fv_start[acc] = i; }
            acc += v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n_valid = 0;                                  // valid keys sort first: their count = first ~0 position
        int lo = 0, hi = n_pow2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (kn[mid] != ~0ull) lo = mid + 1; else hi = mid; }
        n_valid = lo;
        fv_start[s_cnt[1]] = n_valid;
        counts[0] = s_cnt[0]; counts[1] = s_cnt[1]; counts[2] = n_valid;
    }
}

void launch_bow_descend(cudaStream_t st, const VocabDev& voc, int n, const uint8_t* desc, int nid_level, int* f_word, double* f_weight,
                        int* f_node) {
    if (n <= 0) return;
    bow_descend_kernel<<<(n + 7) / 8, 256, 0, st>>>(voc, n, desc, nid_level, f_word, f_weight, f_node);
}

int launch_bow_assemble(cudaStream_t st, int n, const int* f_word, const double* f_weight, const int* f_node, int* bow_word,
                        double* bow_value, int* fv_node, int* fv_start, int* fv_feature, int* counts, int* scratch) {
    int n_pow2 = 32;
    while (n_pow2 < n) n_pow2 <<= 1;
    const size_t smem = (size_t)2 * n_pow2 * sizeof(unsigned long long);
    if (smem > 200 * 1024) return -1;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaFuncSetAttribute(bow_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        done[dev] = true;
    }
    bow_assemble_kernel<<<1, 1024, smem, st>>>(n, n_pow2, f_word, f_weight, f_node, bow_word, bow_value, fv_node, fv_start, fv_feature,
                                              counts, scratch);
    return n_pow2;
}

}  // namespace rgbl
