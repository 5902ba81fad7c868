"""Mapping-thread oracle functions (oracle/mapping_oracle.cpp) against direct numpy restatements.  CPU only."""
import numpy as np

import oracle

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def test_distinctive_descriptors_vs_numpy():
    rng = np.random.default_rng(3)
    sizes = [0, 1, 2, 3, 7, 8, 33, 64, 100, 5, 0, 12]
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (start[-1], 32), dtype=np.uint8)
    desc[start[4]:start[4] + 3] = desc[start[4]]                  # duplicates -> tied medians
    best = oracle.distinctive_descriptors(start, desc)
    for p, n in enumerate(sizes):
        if n == 0:
            assert best[p] == -1
            continue
        d = desc[start[p]:start[p + 1]]
        D = POP[d[:, None, :] ^ d[None, :, :]].sum(axis=2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(med))
