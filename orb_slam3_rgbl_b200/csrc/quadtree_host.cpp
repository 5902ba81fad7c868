// Host-side quad-tree keypoint distribution (product code; used between the FAST kernel and the
// describe kernel).  Semantics of the reference's ORBextractor::DistributeOctTree
// (src/ORBextractor.cc:555-779) and ExtractorNode::DivideNode (:480-536), re-designed around flat
// storage: nodes live in one arena, the "list" is an intrusive doubly-linked index chain, and every
// node's keys are a contiguous segment of one permutation array that is stably 4-way partitioned in
// place when the node is divided (children own sub-segments in n1..n4 order).  No per-node vectors.
//
// What must match the reference exactly:
//   * list order: children are inserted at the FRONT in n1,n2,n3,n4 order, the parent is erased;
//   * key order inside a child = stable partition of the parent's order;
//   * the budgeted expansion sorts (size, UL.x) ascending with std::sort (libstdc++ introsort, so
//     its tie order is inherited by calling std::sort on the same sequence with the same predicate)
//     and walks it from the back, stopping as soon as the list holds >= N nodes;
//   * the survivor of a node is the FIRST key with the maximal response.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "rgbl_internal.h"

namespace rgbl {

namespace {

struct Node {
    int ulx, uly, brx, bry;      // UL and BR corners (UR.x == BR.x, BL.y == BR.y)
    int begin, end;              // key segment in perm[]
    int prev, next;              // list links (-1 = none)
    bool leaf;
};

struct Tree {
    std::vector<Node> nodes;
    std::vector<int32_t> perm, scratch;
    const int32_t* xys;
    int head = -1, tail = -1, count = 0;

    int alloc(const Node& n) { nodes.push_back(n); return (int)nodes.size() - 1; }
    void push_front(int i) {
        nodes[i].prev = -1; nodes[i].next = head;
        if (head >= 0) nodes[head].prev = i; else tail = i;
        head = i; ++count;
    }
    void push_back(int i) {
        nodes[i].next = -1; nodes[i].prev = tail;
        if (tail >= 0) nodes[tail].next = i; else head = i;
        tail = i; ++count;
    }
    void erase(int i) {
        int p = nodes[i].prev, n = nodes[i].next;
        if (p >= 0) nodes[p].next = n; else head = n;
        if (n >= 0) nodes[n].prev = p; else tail = p;
        --count;
    }
    int kx(int k) const { return xys[3 * k]; }
    int ky(int k) const { return xys[3 * k + 1]; }

    // Divide node i; returns indices of the four children (-1 where empty) without linking them.
    void divide(int i, int child[4]) {
        const Node n = nodes[i];
        const int hx = (int)std::ceil(static_cast<float>(n.brx - n.ulx) / 2);
        const int hy = (int)std::ceil(static_cast<float>(n.bry - n.uly) / 2);
        const int mx = n.ulx + hx, my = n.uly + hy;
        int cnt[4] = {0, 0, 0, 0};
        for (int p = n.begin; p < n.end; ++p) {
            const int k = perm[p];
            // kp.pt.x < n1.UR.x compares float(int) with int: exact for our integer coordinates
            const int q = (kx(k) < mx) ? ((ky(k) < my) ? 0 : 2) : ((ky(k) < my) ? 1 : 3);
            ++cnt[q];
        }
        int start[4];
        start[0] = n.begin;
        for (int q = 1; q < 4; ++q) start[q] = start[q - 1] + cnt[q - 1];
        int w[4] = {start[0], start[1], start[2], start[3]};
        for (int p = n.begin; p < n.end; ++p) {
            const int k = perm[p];
            const int q = (kx(k) < mx) ? ((ky(k) < my) ? 0 : 2) : ((ky(k) < my) ? 1 : 3);
            scratch[w[q]++] = k;
        }
        std::copy(scratch.begin() + n.begin, scratch.begin() + n.end, perm.begin() + n.begin);
        const int cx0[4] = {n.ulx, mx, n.ulx, mx}, cy0[4] = {n.uly, n.uly, my, my};
        const int cx1[4] = {mx, n.brx, mx, n.brx}, cy1[4] = {my, my, n.bry, n.bry};
        for (int q = 0; q < 4; ++q) {
            if (cnt[q] == 0) { child[q] = -1; continue; }
            Node c;
            c.ulx = cx0[q]; c.uly = cy0[q]; c.brx = cx1[q]; c.bry = cy1[q];
            c.begin = start[q]; c.end = start[q] + cnt[q];
            c.prev = c.next = -1;
            c.leaf = (cnt[q] == 1);
            child[q] = alloc(c);
        }
    }
};

struct Open { int size; int node; int ulx; };

}  // namespace

int quadtree_select(const int32_t* xys, int n, int min_x, int max_x, int min_y, int max_y, int n_desired,
                    int32_t* out_idx, int cap) {
    if (n <= 0) return 0;
    Tree t;
    t.xys = xys;
    t.perm.resize(n);
    t.scratch.resize(n);
    t.nodes.reserve(4 * (size_t)std::max(n_desired, 16) + 64);

    const int n_ini = (int)std::round(static_cast<float>(max_x - min_x) / (max_y - min_y));
    if (n_ini <= 0) return RGBL_E_UNSUPPORTED;
    const float hX = static_cast<float>(max_x - min_x) / n_ini;

    // Root assignment keeps input order inside each root: counting sort by root index.
    std::vector<int> rcount(n_ini + 1, 0);
    std::vector<int> root_of(n);
    for (int k = 0; k < n; ++k) {
        int r = (int)(static_cast<float>(xys[3 * k]) / hX);
        if (r >= n_ini) r = n_ini - 1;           // cannot happen for x < max_x-min_x; guards the arena
        root_of[k] = r;
        ++rcount[r + 1];
    }
    for (int r = 0; r < n_ini; ++r) rcount[r + 1] += rcount[r];
    {
        std::vector<int> w(rcount.begin(), rcount.end() - 1);
        for (int k = 0; k < n; ++k) t.perm[w[root_of[k]]++] = k;
    }
    for (int r = 0; r < n_ini; ++r) {
        const int b = rcount[r], e = rcount[r + 1];
        if (b == e) continue;                    // empty roots are erased (:590-601)
        Node nd;
        nd.ulx = (int)(hX * static_cast<float>(r)); nd.uly = 0;
        nd.brx = (int)(hX * static_cast<float>(r + 1)); nd.bry = max_y - min_y;
        nd.begin = b; nd.end = e; nd.prev = nd.next = -1;
        nd.leaf = (e - b == 1);
        t.push_back(t.alloc(nd));
    }

    std::vector<Open> open, work;
    auto link_children = [&](const int child[4], int& n_expand) {
        for (int q = 0; q < 4; ++q) {
            if (child[q] < 0) continue;
            t.push_front(child[q]);
            const Node& c = t.nodes[child[q]];
            if (c.end - c.begin > 1) {
                ++n_expand;
                open.push_back({c.end - c.begin, child[q], c.ulx});
            }
        }
    };

    bool done = false;
    while (!done) {
        const int prev = t.count;
        int n_expand = 0;
        open.clear();
        for (int i = t.head; i >= 0;) {
            const int nxt = t.nodes[i].next;     // children go to the front, never revisited this pass
            if (!t.nodes[i].leaf) {
                int child[4];
                t.divide(i, child);
                link_children(child, n_expand);
                t.erase(i);
            }
            i = nxt;
        }
        if (t.count >= n_desired || t.count == prev) {
            done = true;
        } else if (t.count + n_expand * 3 > n_desired) {
            while (!done) {
                const int prev2 = t.count;
                work.swap(open);
                open.clear();
                std::sort(work.begin(), work.end(), [](const Open& a, const Open& b) {
                    if (a.size < b.size) return true;
                    if (a.size > b.size) return false;
                    return a.ulx < b.ulx;
                });
                for (int j = (int)work.size() - 1; j >= 0; --j) {
                    int child[4], dummy = 0;
                    t.divide(work[j].node, child);
                    link_children(child, dummy);
                    t.erase(work[j].node);
                    if (t.count >= n_desired) break;
                }
                if (t.count >= n_desired || t.count == prev2) done = true;
            }
        }
    }

    int m = 0;
    for (int i = t.head; i >= 0; i = t.nodes[i].next) {
        const Node& nd = t.nodes[i];
        int best = t.perm[nd.begin];
        int best_s = xys[3 * best + 2];
        for (int p = nd.begin + 1; p < nd.end; ++p) {
            const int k = t.perm[p];
            if (xys[3 * k + 2] > best_s) { best = k; best_s = xys[3 * k + 2]; }
        }
        if (m < cap) out_idx[m] = best;
        ++m;
    }
    return (m <= cap) ? m : RGBL_E_CAPACITY;
}

}  // namespace rgbl
