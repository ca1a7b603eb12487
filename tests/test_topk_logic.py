"""CPU models of the two ideas the CUDA top-k rests on (csrc/topk.cu), checked against the oracle's
std::partial_sort on thousands of tie-heavy rows.  They restate the kernel's logic in plain Python; the
kernel itself is compared with the oracle in tests/test_gpu_stages.py / test_gpu_parity.py.

1. Floor mode: start the selection at a lower bound of the M-th largest value taken from per-word maxima,
   and accept the result only when it is free of ties.  Claim: whenever the detector accepts, the M largest
   (value, index) pairs in descending order ARE what the sequential heap algorithm returns.
2. WarpHeap::adjust: libstdc++'s __adjust_heap + __push_heap with every comparison made up front (node k's two
   children; every element against the value) and the walk done on the resulting bit masks.
"""
import numpy as np
import pytest

WORD = 32


# ------------------------------------------------------------------------------------------------------
# 1. floor mode
# ------------------------------------------------------------------------------------------------------
def word_maxima(row: np.ndarray) -> np.ndarray:
    """What the aggregation kernel writes per 32-pixel word: max over (survivor ? value : +0)."""
    words = row.reshape(-1, WORD)
    return np.where(words != 0, words, np.float32(0)).max(axis=1).astype(np.float32)


def floor_from_word_maxima(wm: np.ndarray, M: int) -> float:
    """WordMaxCandidates: lane l keeps the 4 largest of words l, l+32, ...; M pops of the warp maximum."""
    cands = []
    for lane in range(32):
        mine = np.sort(wm[lane::32])[::-1][:4]
        cands.extend(mine.tolist())
    cands.sort(reverse=True)
    return cands[M - 1] if len(cands) >= M else float("-inf")


def floor_mode(row: np.ndarray, M: int):
    """Returns (values, indices) or None when the detector calls the result ambiguous / the floor is unusable."""
    wm = word_maxima(row)
    floor_v = floor_from_word_maxima(wm, M)
    if not floor_v > 0:
        return None
    below = np.nextafter(np.float32(floor_v), np.float32(-np.inf))
    kept = [(below, -1)] * M          # sorted descending by value; placeholders
    top, evicted, any_evicted, rej_eq = below, -np.inf, False, -np.inf
    for w in range(len(wm)):
        if not wm[w] >= top:
            continue
        for j in range(WORD):
            idx = w * WORD + j
            cv = row[idx]
            if not cv >= top:
                continue
            if cv > top:                                   # SortedSink::insert
                pos = sum(1 for v, _ in kept if v >= cv)
                evicted, any_evicted = top, True
                kept.insert(pos, (cv, idx))
                kept.pop()
                top = kept[-1][0]
            else:                                          # cv == top
                rej_eq = cv
    vals = [v for v, _ in kept]
    tie = any(vals[i] == vals[i + 1] for i in range(M - 1))
    if tie or (any_evicted and evicted == top) or rej_eq == top:
        return None
    assert all(i >= 0 for _, i in kept), "a placeholder survived although the floor promised M elements above it"
    return np.array(vals, np.float32), np.array([i for _, i in kept], np.int32)


@pytest.mark.parametrize("levels,peaks,M", [(8, 200, 30), (64, 120, 30), (1024, 60, 30), (4096, 400, 30), (16, 40, 4)])
def test_floor_mode_result_is_the_heap_result_whenever_it_is_accepted(levels, peaks, M, oracle):
    rng = np.random.default_rng(levels * 1000 + peaks)
    K, H, W = 48, 64, 64                       # H*W >= 64*M (the partial_sort regime of torch's topk)
    rows = np.zeros((K, H * W), np.float32)    # zeros = suppressed pixels
    for k in range(K):
        n = peaks if k % 4 else max(3, M // 2)            # every fourth row has fewer than M positive peaks
        pos = rng.choice(H * W, n, replace=False)
        rows[k, pos] = rng.integers(1, levels + 1, n).astype(np.float32) / np.float32(levels)
        neg = rng.choice(H * W, 50, replace=False)
        neg = neg[rows[k, neg] == 0]
        rows[k, neg] = -rng.random(len(neg)).astype(np.float32)      # surviving negative local maxima
    tags = np.zeros((K, H, W, 1), np.float32)
    _, _, scores, idx = oracle.top_k(rows.reshape(K, H, W), tags, M)
    accepted = 0
    for k in range(K):
        got = floor_mode(rows[k], M)
        if got is None:
            continue
        accepted += 1
        assert np.array_equal(got[1], idx[k]), f"row {k}: indices differ"
        assert np.array_equal(got[0].view(np.uint32), scores[k].view(np.uint32))
    if levels >= 1024:
        assert accepted > 0                    # fine value grids: the floor path is actually exercised


@pytest.mark.parametrize("M,dups", [(30, 0), (30, 1), (30, 3), (8, 2)])
def test_floor_mode_with_ties_planted_around_the_boundary(M, dups, oracle):
    """Distinct random peaks, then `dups` copies of values ranked near M are planted at random places: ties
    inside the top M, exactly at the M-th value, or just below it.  Whatever the detector accepts must be right,
    and with no planted copy it must accept (distinct values never tie)."""
    rng = np.random.default_rng(100 * M + dups)
    K, H, W = 96, 64, 64
    rows = np.zeros((K, H * W), np.float32)
    for k in range(K):
        n = int(rng.integers(M + 5, 300))
        pos = rng.choice(H * W, n, replace=False)
        rows[k, pos] = (rng.random(n).astype(np.float32) + np.float32(0.01))
        order = pos[np.argsort(-rows[k, pos])]
        for _ in range(dups):
            src = order[int(rng.integers(max(0, M - 4), min(n, M + 3)))]     # a value ranked M-3 .. M+3
            free = np.flatnonzero(rows[k] == 0)
            rows[k, rng.choice(free)] = rows[k, src]
    tags = np.zeros((K, H, W, 1), np.float32)
    _, _, scores, idx = oracle.top_k(rows.reshape(K, H, W), tags, M)
    accepted = usable = 0
    for k in range(K):
        usable += floor_from_word_maxima(word_maxima(rows[k]), M) > 0    # M words with a positive peak
        got = floor_mode(rows[k], M)
        if got is None:
            continue
        accepted += 1
        assert np.array_equal(got[1], idx[k]), f"row {k}: indices differ"
        assert np.array_equal(got[0].view(np.uint32), scores[k].view(np.uint32))
    if dups == 0:
        assert accepted == usable and usable > K // 2
    else:
        assert 0 < accepted < K          # both outcomes occur: copies below the M-th value do not matter


# ------------------------------------------------------------------------------------------------------
# 2. ballot-based __adjust_heap
# ------------------------------------------------------------------------------------------------------
def adjust_heap_literal(v, i, hole, length, val, idx):
    """libstdc++ std::__adjust_heap + std::__push_heap with comp(a, b) = a.value > b.value."""
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if v[child] > v[child - 1]:
            child -= 1
        v[hole], i[hole] = v[child], i[child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        v[hole], i[hole] = v[child - 1], i[child - 1]
        hole = child - 1
    parent = (hole - 1) // 2
    while hole > top and v[parent] > val:
        v[hole], i[hole] = v[parent], i[parent]
        hole = parent
        parent = (hole - 1) // 2
    v[hole], i[hole] = val, idx


def adjust_heap_ballots(v, i, hole, length, val, idx):
    """WarpHeap::adjust: slot k in lane k; two ballots, a scalar walk, one shuffle."""
    lanes = 32
    vv = list(v) + [0.0] * (lanes - len(v))
    ii = list(i) + [0] * (lanes - len(i))
    take_left = 0
    above = 0
    for lane in range(lanes):
        lc, rc = min(2 * lane + 1, 31), min(2 * lane + 2, 31)
        if vv[rc] > vv[lc]:
            take_left |= 1 << lane
        if vv[lane] > val:
            above |= 1 << lane
    p = [hole] + [0] * 6
    n, child, finished = 0, hole, False
    for k in range(1, 7):
        if finished:
            continue
        if child < (length - 1) // 2:
            c = 2 * (child + 1)
            if (take_left >> child) & 1:
                c -= 1
            p[k], n, child = c, k, c
        else:
            if (length & 1) == 0 and child == (length - 2) // 2:
                p[k], n = 2 * (child + 1) - 1, k
            finished = True
    j = n
    for k in range(6, 0, -1):
        if k <= n and j == k and (above >> p[k]) & 1:
            j = k - 1
    src = list(range(lanes))
    for lane in range(lanes):
        for k in range(6):
            if k < j and lane == p[k]:
                src[lane] = p[k + 1]
    nv = [vv[s] for s in src]
    ni = [ii[s] for s in src]
    nv[p[j]], ni[p[j]] = val, idx
    v[:] = nv[:len(v)]
    i[:] = ni[:len(i)]


def test_ballot_adjust_heap_equals_libstdcxx_adjust_heap():
    rng = np.random.default_rng(7)
    for trial in range(4000):
        M = int(rng.integers(1, 33))
        levels = int(rng.choice([2, 3, 8, 1000]))
        v = (rng.integers(0, levels, M) / levels).astype(np.float32).tolist()
        i = list(range(M))
        a_v, a_i, b_v, b_i = list(v), list(i), list(v), list(i)
        # make_heap, a stream of insertions through the root, then sort_heap: every (hole, length) the kernel uses
        for parent in range((M - 2) // 2, -1, -1):
            adjust_heap_literal(a_v, a_i, parent, M, a_v[parent], a_i[parent])
            adjust_heap_ballots(b_v, b_i, parent, M, b_v[parent], b_i[parent])
            assert a_v == b_v and a_i == b_i
        for t in range(40):
            val = float(np.float32(rng.integers(0, levels) / levels))
            if val > a_v[0]:
                adjust_heap_literal(a_v, a_i, 0, M, val, 1000 + t)
                adjust_heap_ballots(b_v, b_i, 0, M, val, 1000 + t)
                assert a_v == b_v and a_i == b_i
        for last in range(M - 1, 0, -1):
            val_a, idx_a = a_v[last], a_i[last]
            a_v[last], a_i[last] = a_v[0], a_i[0]
            adjust_heap_literal(a_v, a_i, 0, last, val_a, idx_a)
            val_b, idx_b = b_v[last], b_i[last]
            b_v[last], b_i[last] = b_v[0], b_i[0]
            adjust_heap_ballots(b_v, b_i, 0, last, val_b, idx_b)
            assert a_v == b_v and a_i == b_i


# ------------------------------------------------------------------------------------------------------
# 3. the exact path as a log replay (what the kernels do for rows with ties; DESIGN 7b builds on it)
# ------------------------------------------------------------------------------------------------------
def entering_log(row, M, begin, end, seed_values):
    """Elements of [begin, end) that enter the sequential heap, given the M largest VALUES seen before `begin`
    (whether an element enters depends only on the M-th largest value before it, not on the heap's shape)."""
    kept = sorted(seed_values, reverse=True)
    log = []
    for t in range(begin, end):
        if row[t] > kept[-1]:
            log.append(t)
            kept.append(row[t])
            kept.sort(reverse=True)
            kept.pop()
    return log, kept


def replay(row, M, log):
    """std::partial_sort's heap: make_heap on the first M elements, __adjust_heap(0, M, x) per entering x,
    sort_heap -- using the literal functions above."""
    v = [float(x) for x in row[:M]]
    i = list(range(M))
    for parent in range((M - 2) // 2, -1, -1):
        adjust_heap_literal(v, i, parent, M, v[parent], i[parent])
    for t in log:
        assert row[t] > v[0]
        adjust_heap_literal(v, i, 0, M, float(row[t]), t)
    for last in range(M - 1, 0, -1):
        val, idx = v[last], i[last]
        v[last], i[last] = v[0], i[0]
        adjust_heap_literal(v, i, 0, last, val, idx)
    return np.array(v, np.float32), np.array(i, np.int32)


@pytest.mark.parametrize("levels,segments", [(4, 1), (16, 8), (1000, 8), (16, 5)])
def test_segmented_logs_replayed_through_the_heap_give_the_reference_order(levels, segments, oracle):
    """Tie-heavy rows: the log of entering elements, produced per segment from the top-M values of the
    segment's prefix, replayed through the libstdc++ heap, is bit-for-bit torch's / the oracle's top-k."""
    rng = np.random.default_rng(levels + segments)
    K, H, W, M = 24, 64, 64, 30
    N = H * W
    rows = np.zeros((K, N), np.float32)
    for k in range(K):
        n = int(rng.integers(5, 600))
        pos = rng.choice(N, n, replace=False)
        rows[k, pos] = rng.integers(-levels, levels + 1, n).astype(np.float32) / np.float32(levels)
    _, _, scores, idx = oracle.top_k(rows.reshape(K, H, W), np.zeros((K, H, W, 1), np.float32), M)
    for k in range(K):
        row = rows[k]
        # geometric segment bounds: the entries thin out like 1/t
        bounds = [M] + [max(M, N >> (segments - s)) for s in range(1, segments)] + [N]
        bounds = sorted(set(bounds))
        log = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            seed = np.sort(row[:a])[::-1][:M].tolist()          # top-M values of the prefix, any order of ties
            seg_log, _ = entering_log(row, M, a, b, seed)
            log.extend(seg_log)
        whole, _ = entering_log(row, M, M, N, row[:M].tolist())
        assert log == whole                                     # segments find exactly the sequential entries
        got_v, got_i = replay(row, M, log)
        assert np.array_equal(got_i, idx[k])
        assert np.array_equal(got_v.view(np.uint32), scores[k].view(np.uint32))
