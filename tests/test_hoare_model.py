"""CPU test: the parallel formulation of the reference's Hoare partition used by the GPU's reference-exact tie mode
(k_hoare_scan / k_hoare_swap / k_hoare_finish) against a verbatim transcription of partition.cpp:30-60."""
import numpy as np


def hoare_verbatim(v, c):
    v = v.copy()
    idx = np.arange(len(v))
    n = len(v)
    i, j = -1, n
    while True:
        i += 1
        while i < n and v[i] < c:           # do i++ while (P(i) < cut ...)      partition.cpp:35-38
            i += 1
        if i >= n:
            return None                     # the reference would read past the cell: degenerate, not emulated
        j -= 1
        while j >= 0 and v[j] > c:          # do j-- while (P(j) > cut ...)      partition.cpp:40-43
            j -= 1
        if i >= j:
            break
        v[i], v[j] = v[j], v[i]
        idx[i], idx[j] = idx[j], idx[i]
    v[i], v[n - 1] = v[n - 1], v[i]         # partition.cpp:52
    idx[i], idx[n - 1] = idx[n - 1], idx[i]
    return idx, i


def hoare_parallel(v, c):
    n = len(v)
    ge, le = v >= c, v <= c
    F = np.cumsum(ge) - ge                  # exclusive prefix counts = ranks from the left
    r = np.cumsum(le) - le
    nge, nle = int(ge.sum()), int(le.sum())
    if nge == 0:
        return None
    posI = np.zeros(nge, int); posI[F[ge]] = np.nonzero(ge)[0]
    posJ = np.zeros(nle, int); posJ[r[le]] = np.nonzero(le)[0]
    idx = np.arange(n)
    K = 0
    for k in range(min(nge, nle)):          # k_hoare_swap: one thread per k
        a, b = posI[k], posJ[nle - 1 - k]
        if a < b:
            idx[a], idx[b] = idx[b], idx[a]
            K += 1
    stop = min(posI[K] if K < nge else 1 << 60, posJ[nle - K] if K > 0 else 1 << 60)   # k_hoare_finish
    idx[stop], idx[n - 1] = idx[n - 1], idx[stop]
    return idx, stop


def test_parallel_hoare_equals_verbatim_loop():
    rng = np.random.default_rng(1)
    checked = 0
    for _ in range(4000):
        n = int(rng.integers(2, 80))
        v = rng.integers(-4, 5, n).astype(np.float32) / 4     # heavy ties
        c = np.float32(rng.integers(-4, 5) / 4)
        a, b = hoare_verbatim(v, c), hoare_parallel(v, c)
        assert (a is None) == (b is None)
        if a is None:
            continue
        assert a[1] == b[1] and np.array_equal(a[0], b[0])
        checked += 1
    assert checked > 3000
