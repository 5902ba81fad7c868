// ORACLE (test infrastructure, NOT product code): CPU restatement of Frame::ComputeBoW (src/Frame.cc:828-835) =
// DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1206, per-feature descent :1218-1259), FORB::distance
// (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101), BowVector::addWeight / normalize (BowVector.cpp:34-46, 62-84) and
// FeatureVector::addFeature (FeatureVector.cpp:27-41) for the TF_IDF / TF weighting with L1 scoring the ORB vocabulary uses.
// The vocabulary is passed flattened (what the C-ABI shim gathers from m_nodes).  Parity unpinned: DBoW2 needs OpenCV and is
// not buildable here; std::map gives the reference's iteration order by construction.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs load this library.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace {

inline int hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        std::memcpy(&x, a + 8 * i, 8); std::memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

}  // namespace

extern "C" {

// returns 0; n_words / n_fv_nodes out.  bow_word/bow_value: ascending word ids (map order); fv as CSR: ascending node ids,
// fv_start[n_fv_nodes + 1], feature indices in insertion (= feature) order.
int orc_compute_bow(int n_nodes, const int32_t* child_begin, const int32_t* child_index, const uint8_t* node_desc,
                    const double* node_weight, const int32_t* word_id, int L, int n, const uint8_t* desc, int levelsup,
                    int32_t* bow_word, double* bow_value, int* n_words, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feature,
                    int* n_fv_nodes) {
    std::map<unsigned, double> v;                          // BowVector
    std::map<unsigned, std::vector<unsigned>> fv;          // FeatureVector
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; ++f) {
        const uint8_t* q = desc + (size_t)f * 32;
        unsigned nid = 0;                                  // if nid_level <= 0: root
        int final_id = 0, current_level = 0;
        if (n_nodes > 1 && child_begin[1] > child_begin[0]) {
            do {
                ++current_level;
                const int cb = child_begin[final_id], ce = child_begin[final_id + 1];
                final_id = child_index[cb];
                double best_d = hamming256(q, node_desc + (size_t)final_id * 32);
                for (int c = cb + 1; c < ce; ++c) {
                    const int id = child_index[c];
                    const double d = hamming256(q, node_desc + (size_t)id * 32);
                    if (d < best_d) { best_d = d; final_id = id; }
                }
                if (current_level == nid_level) nid = (unsigned)final_id;
            } while (child_begin[final_id + 1] > child_begin[final_id]);   // !isLeaf()
        }
        const double w = node_weight[final_id];
        if (w > 0) {                                       // not stopped
            const unsigned id = (unsigned)word_id[final_id];
            auto it = v.lower_bound(id);                   // BowVector::addWeight
            if (it != v.end() && !(id < it->first)) it->second += w; else v.insert(it, {id, w});
            fv[nid].push_back((unsigned)f);                // FeatureVector::addFeature
        }
    }
    double norm = 0.0;                                     // BowVector::normalize(L1)
    for (auto& kv : v) norm += std::fabs(kv.second);
    if (norm > 0.0) for (auto& kv : v) kv.second /= norm;
    int k = 0;
    for (auto& kv : v) { bow_word[k] = (int32_t)kv.first; bow_value[k] = kv.second; ++k; }
    *n_words = k;
    int m = 0, o = 0;
    for (auto& kv : fv) {
        fv_node[m] = (int32_t)kv.first; fv_start[m] = o;
        for (unsigned idx : kv.second) fv_feature[o++] = (int32_t)idx;
        ++m;
    }
    fv_start[m] = o;
    *n_fv_nodes = m;
    return 0;
}

}  // extern "C"
