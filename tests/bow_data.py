"""Synthetic DBoW2-style vocabularies (no ORBvoc.txt here): k-ary trees of random 256-bit descriptors, flattened the way the
C-ABI shim flattens TemplatedVocabulary::m_nodes (node 0 = root, children in list order)."""
import numpy as np


def make_vocabulary(seed: int, k: int = 10, levels: int = 3, ragged: bool = False, stop_fraction: float = 0.05, shuffle_children: bool = True):
    """ragged: some inner nodes get fewer children and some become leaves above the last level (what k-means produces when a
    cluster cannot be split).  stop_fraction of the words get weight 0 (stopped words are skipped by transform)."""
    rng = np.random.default_rng(seed)
    children = [[]]                     # node -> list of child ids, built breadth first
    depth = [0]
    frontier = [0]
    for lv in range(levels):
        nxt = []
        for node in frontier:
            if ragged and lv > 0 and rng.random() < 0.15:
                continue                # stays a leaf above the last level
            nc = k if not ragged else int(rng.integers(2, k + 1))
            for _ in range(nc):
                children.append([]); depth.append(lv + 1)
                children[node].append(len(children) - 1)
                nxt.append(len(children) - 1)
        frontier = nxt
    n = len(children)
    if shuffle_children:                # the list order decides ties, not the node id
        for c in children:
            rng.shuffle(c)
    child_begin = np.zeros(n + 1, np.int32)
    for i, c in enumerate(children):
        child_begin[i + 1] = child_begin[i] + len(c)
    child_index = np.array([x for c in children for x in c], np.int32)
    node_desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    # make sibling descriptors related to their parent so that descents are not pure noise, and add exact duplicates (ties)
    for i, c in enumerate(children):
        for j, ch in enumerate(c):
            flip = rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8) & rng.integers(0, 256, 32, dtype=np.uint8)
            node_desc[ch] = node_desc[i] ^ flip
        if len(c) >= 3 and rng.random() < 0.3:
            node_desc[c[2]] = node_desc[c[0]]
    word_id = np.full(n, -1, np.int32)
    leaves = [i for i, c in enumerate(children) if not c and i != 0]
    word_id[leaves] = np.arange(len(leaves), dtype=np.int32)
    node_weight = np.zeros(n, np.float64)
    w = rng.uniform(0.1, 9.0, len(leaves))
    w[rng.random(len(leaves)) < stop_fraction] = 0.0
    node_weight[leaves] = w
    return dict(child_begin=child_begin, child_index=child_index, node_desc=node_desc, node_weight=node_weight, word_id=word_id,
                levels=levels, n_words=len(leaves))


def descriptors_near_words(vocab: dict, n: int, seed: int, noise_bits: int = 20):
    """n descriptors: leaf descriptors with a few flipped bits (so that several features share a word) + some pure noise."""
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(vocab["word_id"] >= 0)[0]
    pick = rng.permutation(leaves)[:max(8, n // 3)][rng.integers(0, min(len(leaves), max(8, n // 3)), n)]
    d = vocab["node_desc"][pick].copy()
    for i in range(n):
        bits = rng.integers(0, 256, noise_bits)
        for b in bits:
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    noise = rng.random(n) < 0.1
    d[noise] = rng.integers(0, 256, (int(noise.sum()), 32), dtype=np.uint8)
    return d
