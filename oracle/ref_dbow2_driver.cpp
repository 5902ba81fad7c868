// ORACLE SUPPORT (test infrastructure, NOT product code): a C entry to the REFERENCE's own DBoW2, compiled unmodified from
// /root/reference/Thirdparty/DBoW2 (oracle/Makefile target `ref` -> oracle/_ref/libref_dbow2.so).  It does what
// Frame::ComputeBoW does (src/Frame.cc:828-835): ORBVocabulary::transform(descriptors, BowVector, FeatureVector, levelsup)
// with ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h:28) loaded through
// loadFromTextFile (the ORBvoc.txt format, src/System.cc:112), and is used to pin oracle/bow_oracle.cpp.
#include <cstdint>
#include <cstring>
#include <vector>

#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;

extern "C" {

void* ref_voc_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}

void ref_voc_free(void* v) { delete static_cast<ORBVocabulary*>(v); }

int ref_voc_size(void* v) { return (int)static_cast<ORBVocabulary*>(v)->size(); }

// desc: n x 32 bytes.  Outputs like oracle.compute_bow: the BowVector as (word id, value) in map order, the FeatureVector as CSR
// (node ids ascending, start offsets, feature indices).  Returns 0, or -1 when a capacity is too small.
int ref_voc_transform(void* vp, const uint8_t* desc, int n, int levelsup, int cap_words, int* n_words, uint32_t* word, double* value,
                      int cap_nodes, int* n_nodes, uint32_t* node, int* node_start, int* feature) {
    ORBVocabulary* voc = static_cast<ORBVocabulary*>(vp);
    std::vector<cv::Mat> features(n);                          // Converter::toDescriptorVector (src/Converter.cc:31-39): one row each
    for (int i = 0; i < n; ++i) { features[i].create(1, 32, CV_8U); std::memcpy(features[i].ptr<unsigned char>(), desc + 32 * (size_t)i, 32); }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc->transform(features, bv, fv, levelsup);
    if ((int)bv.size() > cap_words || (int)fv.size() > cap_nodes) return -1;
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { word[k] = it->first; value[k] = it->second; }
    *n_words = k;
    int m = 0, f = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++m) {
        node[m] = it->first; node_start[m] = f;
        for (size_t j = 0; j < it->second.size(); ++j) feature[f++] = (int)it->second[j];
    }
    node_start[m] = f;
    *n_nodes = m;
    return 0;
}

}  // extern "C"
