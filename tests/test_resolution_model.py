"""Model check of the matcher resolution scheme (orb_slam3_rgbl_b200/csrc/match_kernels.cu: resolve_kernel).  The reference's matchers
are sequential greedy loops: map point q takes its best still-available frame feature and may block it for the points after it.  The
device resolves them in rounds: every waiting query proposes itself at each available candidate (minimum index wins per feature), and a
query becomes final when no lower-index waiting query can still interfere.  This test restates both in Python and lets hypothesis
compare them on small random instances full of ties, including both processing orders inside a round (the device has a documented
benign race on the state array inside one phase).  CPU only; it checks the ALGORITHM the kernel implements, the kernel itself is
checked against the oracle on the GPU."""
import numpy as np
from hypothesis import given, settings, strategies as st

TH_HIGH, TH_LOW = 100, 50


def sequential(mode, lists, obs_pos, state0, lvl, ratio, th_accept):
    """lists[q] = [(dist, pos, ft)] in scan order (ascending pos).  Returns (choice[q], state)."""
    state = list(state0); choice = [-1] * len(lists)
    for q, cand in enumerate(lists):
        best = best2 = None
        for (d, pos, ft) in cand:
            if state[ft] == 1:
                continue
            key = (d, pos)
            if mode == 0:
                if best is None or key < best[0]:
                    best = (key, ft)
            else:
                if best is None or key < best[0]:
                    best2 = best; best = (key, ft)
                elif best2 is None or key < best2[0]:
                    best2 = (key, ft)
        if best is None:
            continue
        bd = best[0][0]
        accept = bd <= th_accept
        if mode == 1 and accept and best2 is not None and lvl[best[1]] == lvl[best2[1]] and float(np.float32(bd)) > float(np.float32(ratio) * np.float32(best2[0][0])):
            accept = False
        if mode == 1 and accept and best2 is None:
            pass                                                # bestDist2 = 256, bestLevel2 = -1: levels differ
        if mode == 2 and accept:
            bd2 = 256 if best2 is None else best2[0][0]
            accept = float(np.float32(bd)) < float(np.float32(ratio) * np.float32(bd2))
        if accept:
            choice[q] = best[1]
            if obs_pos[q]:
                state[best[1]] = 1
            elif state[best[1]] == 0:
                state[best[1]] = 2
    return choice, state


def rounds(mode, lists, obs_pos, state0, lvl, ratio, th_accept, reverse, top2=True, inverse=False):
    """top2: a query depends only on its best and second-best AVAILABLE candidates (what the kernel does since round 2: lists sorted
    by key, the scan stops at the second available entry); False: on every available candidate (the first formulation)."""
    state = list(state0); n_q = len(lists); choice = [-1] * n_q
    resolved = [len(c) == 0 for c in lists]
    # inverse=True: what the kernel does since round 2 - no per-round proposal table; a query is blocked at feature ft when a lower-index
    # query that lists ft was still waiting when the round BEGAN (queries that become final during the round still count), and the
    # availability of ft is read live (it may already reflect decisions taken earlier in the same round)
    listing = {}
    for q, cand in enumerate(lists):
        for (_, _, ft) in cand:
            listing.setdefault(ft, []).append(q)
    for _ in range(4 * n_q + 4):
        waiting_at_start = [not r for r in resolved]
        minq = {}
        for q in range(n_q):
            if resolved[q]:
                continue
            for (d, pos, ft) in lists[q]:
                if state[ft] != 1:
                    minq[ft] = min(minq.get(ft, 1 << 30), q)
        waiting = False
        order = range(n_q - 1, -1, -1) if reverse else range(n_q)
        for q in order:
            if resolved[q]:
                continue
            best = best2 = None; depends_ok = True
            for (d, pos, ft) in lists[q]:
                if state[ft] == 1:
                    continue
                if mode >= 1 and minq.get(ft) != q:
                    depends_ok = False
                key = (d, pos)
                if mode == 0:
                    if best is None or key < best[0]:
                        best = (key, ft)
                else:
                    if best is None or key < best[0]:
                        best2 = best; best = (key, ft)
                    elif best2 is None or key < best2[0]:
                        best2 = (key, ft)
            if best is None:
                resolved[q] = True
                continue
            if top2:
                depends_ok = minq.get(best[1]) == q and (best2 is None or minq.get(best2[1]) == q)
            if inverse:
                blocked = lambda ft: any(o < q and waiting_at_start[o] for o in listing.get(ft, []))
                depends_ok = not blocked(best[1]) and (best2 is None or not blocked(best2[1]))
                if mode == 0:
                    depends_ok = not blocked(best[1])
            final_ok = depends_ok if (inverse or mode != 0) else (minq.get(best[1]) == q)
            if not final_ok:
                waiting = True
                continue
            resolved[q] = True
            bd = best[0][0]
            accept = bd <= th_accept
            if mode == 1 and accept and best2 is not None and lvl[best[1]] == lvl[best2[1]] and float(np.float32(bd)) > float(np.float32(ratio) * np.float32(best2[0][0])):
                accept = False
            if mode == 2 and accept:
                bd2 = 256 if best2 is None else best2[0][0]
                accept = float(np.float32(bd)) < float(np.float32(ratio) * np.float32(bd2))
            if accept:
                choice[q] = best[1]
                if obs_pos[q]:
                    state[best[1]] = 1
                elif state[best[1]] == 0:
                    state[best[1]] = 2
        if not waiting:
            return choice, state
    raise AssertionError("the rounds did not terminate")


@st.composite
def instance(draw):
    n_f = draw(st.integers(1, 7)); n_q = draw(st.integers(1, 10))
    lists = []
    for _ in range(n_q):
        fts = draw(st.lists(st.integers(0, n_f - 1), max_size=5, unique=True))
        # pos == ft here (any injective scan order works); distances from a tiny alphabet -> many ties
        lists.append(sorted([(draw(st.sampled_from([10, 10, 30, 50, 51, 100, 101])), ft, ft) for ft in fts], key=lambda c: c[1]))
    obs_pos = draw(st.lists(st.booleans(), min_size=n_q, max_size=n_q))
    state0 = draw(st.lists(st.sampled_from([0, 0, 0, 1]), min_size=n_f, max_size=n_f))
    lvl = draw(st.lists(st.integers(0, 2), min_size=n_f, max_size=n_f))
    ratio = draw(st.sampled_from([0.6, 0.8, 0.9, 1.0]))
    return lists, obs_pos, state0, lvl, ratio


@settings(max_examples=3000, deadline=None)
@given(instance(), st.sampled_from([0, 1, 2]), st.booleans())
def test_rounds_equal_the_sequential_greedy_matchers(inst, mode, reverse):
    lists, obs_pos, state0, lvl, ratio = inst
    if mode == 2:
        obs_pos = [True] * len(obs_pos)          # SearchByBoW: every assignment blocks the feature (src/ORBmatcher.cc:296-297)
    th = TH_LOW if mode == 2 else TH_HIGH
    ref = sequential(mode, lists, obs_pos, state0, lvl, ratio, th)
    got = rounds(mode, lists, obs_pos, state0, lvl, ratio, th, reverse)
    assert got == ref
    assert rounds(mode, lists, obs_pos, state0, lvl, ratio, th, reverse, top2=False) == ref
    assert rounds(mode, lists, obs_pos, state0, lvl, ratio, th, reverse, inverse=True) == ref


def three_maxima_sequential(h):
    """ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:2012-2053)."""
    m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
    for i, s in enumerate(h):
        if s > m1:
            m3, m2, m1 = m2, m1, s; i3, i2, i1 = i2, i1, i
        elif s > m2:
            m3, m2 = m2, s; i3, i2 = i2, i
        elif s > m3:
            m3 = s; i3 = i
    if np.float32(m2) < np.float32(0.1) * np.float32(m1):
        i2 = i3 = -1
    elif np.float32(m3) < np.float32(0.1) * np.float32(m1):
        i3 = -1
    return i1, i2, i3


def three_maxima_argmax(h):
    """What resolve_kernel / triangulation_finish_kernel do: three arg-max passes over count << 8 | (255 - index), empty bins excluded."""
    keys = [((c << 8) | (255 - i)) if c > 0 else 0 for i, c in enumerate(h)]
    top = []
    for _ in range(3):
        m = max(keys)
        top.append((m >> 8, 255 - (m & 0xff) if m > 0 else -1))
        if m > 0:
            keys[keys.index(m)] = 0
    (c1, i1), (c2, i2), (c3, i3) = top
    if np.float32(c2) < np.float32(0.1) * np.float32(c1):
        i2 = i3 = -1
    elif np.float32(c3) < np.float32(0.1) * np.float32(c1):
        i3 = -1
    return i1, i2, i3


@settings(max_examples=2000, deadline=None)
@given(st.lists(st.sampled_from([0, 0, 0, 1, 2, 3, 5, 5, 9, 10, 50, 100]), min_size=30, max_size=30))
def test_parallel_three_maxima_equals_the_sequential_scan(h):
    assert three_maxima_argmax(h) == three_maxima_sequential(h)


@settings(max_examples=2000, deadline=None)
@given(st.lists(st.tuples(st.sampled_from([0, 10, 10, 30, 49, 50, 50, 51, 80]), st.booleans()), max_size=12))
def test_triangulation_last_minimum_key(cands):
    """SearchForTriangulation's scan (src/ORBmatcher.cc:1002-1080) keeps a candidate when dist <= TH_LOW, dist <= bestDist and the
    geometric test passes; triangulation_search_kernel takes the minimum of dist << 20 | (0xfffff - position) over the passing ones."""
    best, idx = TH_LOW, -1
    for i, (d, ok) in enumerate(cands):
        if d > TH_LOW or d > best:
            continue
        if ok:
            idx, best = i, d
    keys = [(d << 20) | (0xfffff - i) for i, (d, ok) in enumerate(cands) if ok and d <= TH_LOW]
    got = (0xfffff - (min(keys) & 0xfffff)) if keys else -1
    assert got == idx


@settings(max_examples=1000, deadline=None)
@given(st.lists(st.integers(0, 256), min_size=1, max_size=70))
def test_median_by_histogram_rank(row):
    """distinctive_kernel: the element std::sort leaves at index (int)(0.5 * (N - 1)) is the first histogram bin whose cumulative count
    exceeds that rank."""
    k = int(0.5 * (len(row) - 1))
    hist = np.bincount(row, minlength=257)
    cum = np.cumsum(hist)
    assert int(np.argmax(cum > k)) == sorted(row)[k]
