"""CPU model of the selection-based cut search (CUDA: csrc/orb_select.cuh).

The reference bisects every cell with up to 32 count passes (orbit.cpp:146-232).  Each decision depends on the
count `#{x < cut}` only through `diff = (int)((float)cnt - (float)n * ratio)`, which is monotone in cnt.  So one
histogram over a monotone bin function brackets every count, the handful of particles whose bins can still flip a
decision (the candidates) are gathered once, and the whole 32-step loop is replayed exactly on them: two reads of the
cut-axis column per level instead of one per (multi-trial) bisection pass.

This file restates that algorithm in numpy float32 and checks it against the literal loop on adversarial inputs
(ties, clusters, signed zeros, denormals, particles outside the box, degenerate boxes).  The GPU implementation is
checked against the oracle in tests/test_gpu_parity.py; this model pins the logic without a GPU."""
import math

import numpy as np
import pytest

f32 = np.float32
MAX_ITER = 32


def mid_cut(L, R):                      # Cell::getCut, cell.h:74-76
    return f32(f32(R + L) * f32(0.5))


def make_prod(total, nleaf):            # orbit.cpp:204
    ratio = f32(math.ceil(nleaf / 2.0) / nleaf)
    return f32(f32(total) * ratio)


def diff_of(cnt, prod):                 # orbit.cpp:205: float subtraction, truncation to int
    return int(np.trunc(f32(f32(cnt) - prod)))


def literal_bisection(v, L, R, total, nleaf):
    """orbit.cpp:149-232 for one cell; returns (L, R, iterations, found, nLeft)."""
    prod = make_prod(total, nleaf)
    it, found, nleft = 0, False, None
    while it < MAX_ITER:
        cut = mid_cut(L, R)
        cnt = int(np.count_nonzero(v < cut))
        d = diff_of(cnt, prod)
        it += 1
        if abs(d) < 3:
            found, nleft = True, cnt
            break
        if d > 0:
            R = cut
        else:
            L = cut
    if not found:
        nleft = int(np.count_nonzero(v < mid_cut(L, R)))
    return L, R, it, found, nleft


def bin_params(lo, hi, nb):
    with np.errstate(all="ignore"):
        w = f32(hi - lo)
        scale = f32(nb) / w if (w > 0 and np.isfinite(w)) else f32(0)
    if not np.isfinite(scale):
        scale = f32(0)
    return f32(lo), f32(scale)


def sel_bin(x, lo, scale, nb):
    """Monotone non-decreasing in x for every float (the only property the algorithm needs)."""
    with np.errstate(all="ignore"):
        t = (np.asarray(x, f32) - lo).astype(f32) * scale
    t = np.where(np.isnan(t), f32(0), t)
    t = np.minimum(np.maximum(t, f32(0)), f32(nb - 1))
    return np.trunc(t).astype(np.int64)


def ambiguous_range(prefix, base, prod):
    """Bins whose count range can still hold |diff| < 3 or flip its sign: [first, last]."""
    nb = len(prefix) - 1
    first = next(b for b in range(nb) if diff_of(base + prefix[b + 1], prod) > -3)
    last = next(b for b in range(nb - 1, -1, -1) if diff_of(base + prefix[b], prod) < 3)
    assert first <= last
    return first, last


def select_bisection(v, L, R, total, nleaf, nb1=512, nb2=2048, amb_cap=2048, cand_cap=40960, bins=None):
    """Histogram -> candidates -> in-block histogram -> replay.  Returns the literal loop's tuple, or None when the
    cell must fall back to the iterative path (too many candidates, or a final cut outside the candidate bins).
    `bins` = (lo, hi): range of the bin function when it is not the cell's margins (zoom rounds of k_sel_percell)."""
    prod = make_prod(total, nleaf)
    lo1, s1 = bin_params(*(bins if bins is not None else (L, R)), nb1)
    b = sel_bin(v, lo1, s1, nb1)
    p1 = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=nb1))])
    f1, l1 = ambiguous_range(p1, 0, prod)
    base = int(p1[f1])
    cand = v[(b >= f1) & (b <= l1)]
    if cand.size > cand_cap:
        return None
    # second histogram over the staged candidates
    if cand.size:
        lo2, s2 = bin_params(cand.min(), cand.max(), nb2)
    else:
        lo2, s2 = f32(0), f32(0)
    b2 = sel_bin(cand, lo2, s2, nb2)
    p2 = np.concatenate([[0], np.cumsum(np.bincount(b2, minlength=nb2))])
    f2, l2 = ambiguous_range(p2, base, prod)
    if p2[l2 + 1] - p2[f2] <= amb_cap:
        amb, base2 = cand[(b2 >= f2) & (b2 <= l2)], base + int(p2[f2])
    else:
        amb, base2, f2, l2 = cand, base, 0, nb2 - 1
    it, found, nleft = 0, False, None
    while it < MAX_ITER:
        cut = mid_cut(L, R)
        c1 = int(sel_bin(cut, lo1, s1, nb1))
        dec = -1 if c1 < f1 else (1 if c1 > l1 else 0)
        if dec == 0:
            c2 = int(sel_bin(cut, lo2, s2, nb2))
            dec = -1 if c2 < f2 else (1 if c2 > l2 else 0)
        it += 1
        if dec == 0:
            cnt = base2 + int(np.count_nonzero(amb < cut))
            d = diff_of(cnt, prod)
            if abs(d) < 3:
                found, nleft = True, cnt
                break
            dec = 1 if d > 0 else -1
        if dec > 0:
            R = cut
        else:
            L = cut
    if not found:
        cut = mid_cut(L, R)
        c1 = int(sel_bin(cut, lo1, s1, nb1))
        if c1 < f1 or c1 > l1:
            return None
        nleft = base + int(np.count_nonzero(cand < cut))
    return L, R, it, found, nleft


def same(a, b):
    return (np.float32(a[0]).tobytes(), np.float32(a[1]).tobytes(), a[2:]) == (np.float32(b[0]).tobytes(), np.float32(b[1]).tobytes(), b[2:])


def cases():
    rng = np.random.default_rng(5)
    out = []
    for n in (1, 2, 5, 6, 7, 100, 4097, 60_000, 300_000):
        out.append((f"uniform{n}", (rng.random(n, dtype=f32) - f32(0.5)), f32(-0.5), f32(0.5), 8))
    v = (rng.normal(0.2, 0.01, 200_000)).clip(-0.5, 0.5).astype(f32)
    out.append(("cluster", v, f32(-0.5), f32(0.5), 16))
    out.append(("cluster_odd_leaves", v, f32(-0.5), f32(0.5), 7))
    t = rng.random(50_000, dtype=f32) - f32(0.5)
    t[::3] = f32(0.125)                                   # a third of the particles tie on one value
    out.append(("ties", t, f32(-0.5), f32(0.5), 4))
    t2 = np.full(30_000, f32(0.25))
    out.append(("all_equal", t2, f32(-0.5), f32(0.5), 2))
    z = rng.random(20_000, dtype=f32) - f32(0.5)
    z[:5000] = f32(0.0)
    z[5000:9000] = f32(-0.0)
    z[9000:9100] = f32(1e-42)
    out.append(("signed_zero_denormal", z, f32(-0.5), f32(0.5), 2))
    o = (rng.random(10_000, dtype=f32) * f32(3) - f32(1.5))
    out.append(("outside_box", o, f32(-0.5), f32(0.5), 2))
    out.append(("degenerate_box", rng.random(3000, dtype=f32), f32(0.3), f32(0.3), 2))
    out.append(("narrow_box", (f32(0.3) + rng.random(5000, dtype=f32) * f32(1e-6)).astype(f32), f32(0.3), f32(0.3000011), 2))
    big = rng.random(1 << 22, dtype=f32) - f32(0.5)       # float(cnt) granularity > 3 never happens below 2^24; still big
    out.append(("big", big, f32(-0.5), f32(0.5), 4096))
    sub = np.sort(rng.random(100_000, dtype=f32))[20_000:30_000] - f32(0.5)   # cell of a deeper level: narrow box
    out.append(("deep_cell", sub, sub.min(), f32(np.nextafter(sub.max(), f32(1)))
                , 3))
    return out


@pytest.mark.parametrize("name,v,L,R,nleaf", cases(), ids=[c[0] for c in cases()])
def test_select_equals_literal_bisection(name, v, L, R, nleaf):
    want = literal_bisection(v, L, R, v.size, nleaf)
    got = select_bisection(v, L, R, v.size, nleaf)
    if got is None:
        assert name in ("all_equal", "degenerate_box", "ties", "narrow_box"), "unexpected fallback"
        return
    assert same(got, want), (got, want)


def test_small_capacities_fall_back_or_agree():
    """Tiny candidate / ambiguity caps exercise the brute-force and fallback branches."""
    rng = np.random.default_rng(9)
    v = rng.random(40_000, dtype=f32) - f32(0.5)
    v[::7] = f32(-0.2)
    want = literal_bisection(v, f32(-0.5), f32(0.5), v.size, 2)
    got = select_bisection(v, f32(-0.5), f32(0.5), v.size, 2, nb1=16, nb2=32, amb_cap=4)
    assert got is not None and same(got, want)
    assert select_bisection(v, f32(-0.5), f32(0.5), v.size, 2, nb1=16, cand_cap=100) is None


def test_global_total_larger_than_local():
    """Multi-rank shape: the decision uses the global total, this rank only sees part of the counts.  The model is
    single-rank (counts are global); here just the monotonicity the algorithm relies on."""
    prod = make_prod(1 << 27, 2)
    d = [diff_of(c, prod) for c in range((1 << 26) - 40, (1 << 26) + 40)]
    assert all(a <= b for a, b in zip(d, d[1:]))


def zoom_range(v, lo, hi, nb, total, nleaf):
    """One zoom step of k_sel_percell: the interval of the candidate bins, widened by 1/16 of its width on each side."""
    prod = make_prod(total, nleaf)
    lo1, s1 = bin_params(lo, hi, nb)
    b = sel_bin(v, lo1, s1, nb)
    p1 = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=nb))])
    f1, l1 = ambiguous_range(p1, 0, prod)
    a = f32(lo1 + f32(f32(f1) / s1))
    e = f32(lo1 + f32(f32(l1 + 1) / s1))
    w = f32(e - a)
    return f32(a - w * f32(0.0625)), f32(e + w * f32(0.0625)), int(p1[l1 + 1] - p1[f1])


def test_zoomed_bin_function_gives_the_same_result():
    """A dense clump inside a wide cell: the first histogram leaves too many candidates; zooming the bin function onto
    the candidate bins (any monotone bin function over the whole cell is valid) isolates the median."""
    rng = np.random.default_rng(21)
    v = np.concatenate([rng.normal(0.31, 0.0004, 180_000), rng.random(20_000) - 0.5]).clip(-0.5, 0.5).astype(f32)
    L, R = f32(-0.5), f32(0.5)
    for nleaf in (2, 5, 64):
        want = literal_bisection(v, L, R, v.size, nleaf)
        assert select_bisection(v, L, R, v.size, nleaf, nb1=512, cand_cap=8192) is None      # too dense for one round
        lo, hi, ncand = L, R, v.size
        for _ in range(2):
            lo, hi, ncand = zoom_range(v, lo, hi, 512 if (lo, hi) == (L, R) else 2048, v.size, nleaf)
            got = select_bisection(v, L, R, v.size, nleaf, nb1=2048, cand_cap=8192, bins=(lo, hi))
            if got is not None:
                break
        assert got is not None and same(got, want), (got, want)


def test_sharded_histograms_and_local_left_count():
    """Several ranks: the histogram rows add up to the global row, the candidates of all ranks are the global
    candidates, and a rank's left count at the final cut is #{own particles below the candidate bins} +
    #{own candidates < cut} (k_selx_resolve / k_selmr_finish)."""
    rng = np.random.default_rng(33)
    v = rng.random(200_000, dtype=f32) - f32(0.5)
    v[::11] = f32(0.003)
    L, R, nleaf, nb1 = f32(-0.5), f32(0.5), 6, 512
    want = literal_bisection(v, L, R, v.size, nleaf)
    got = select_bisection(v, L, R, v.size, nleaf, nb1=nb1)
    assert got is not None and same(got, want)
    cutf = mid_cut(got[0], got[1])
    prod = make_prod(v.size, nleaf)
    lo1, s1 = bin_params(L, R, nb1)
    shards = np.array_split(v, 3)
    rows = [np.bincount(sel_bin(sh, lo1, s1, nb1), minlength=nb1) for sh in shards]
    glob = np.sum(rows, axis=0)
    assert np.array_equal(glob, np.bincount(sel_bin(v, lo1, s1, nb1), minlength=nb1))
    f1, l1 = ambiguous_range(np.concatenate([[0], np.cumsum(glob)]), 0, prod)
    tot = 0
    for sh, row in zip(shards, rows):
        b = sel_bin(sh, lo1, s1, nb1)
        own = sh[(b >= f1) & (b <= l1)]
        nleft_l = int(row[:f1].sum()) + int(np.count_nonzero(own < cutf))
        assert nleft_l == int(np.count_nonzero(sh < cutf))
        tot += nleft_l
    assert tot == want[4]


def test_randomised_inputs_and_capacities():
    """Seeded fuzz: distributions with clumps, lattices (many duplicates), ties, narrow ranges; random bin counts and
    capacities; with and without a zoomed bin function.  Whenever the method answers, it answers like the literal loop."""
    rng = np.random.default_rng(20261017)
    answered = 0
    for _ in range(160):
        n = int(rng.choice([3, 17, 200, 5000, 30000]))
        kind = int(rng.integers(0, 6))
        if kind == 0:
            v = rng.random(n, dtype=f32) - f32(0.5)
        elif kind == 1:
            v = rng.normal(rng.uniform(-0.4, 0.4), 10 ** rng.uniform(-6, -1), n).clip(-0.5, 0.5).astype(f32)
        elif kind == 2:
            v = rng.random(n, dtype=f32) - f32(0.5)
            for _k in range(int(rng.integers(1, 4))):
                v[rng.random(n) < rng.uniform(0.01, 0.4)] = f32(rng.uniform(-0.5, 0.5))
        elif kind == 3:
            v = (np.round((rng.random(n) - 0.5) * 2 ** int(rng.integers(3, 12))) / 2 ** 12).astype(f32)
        elif kind == 4:
            v = np.concatenate([rng.normal(0.1, 1e-4, n // 2), rng.random(n - n // 2) - 0.5]).clip(-0.5, 0.5).astype(f32)
        else:
            v = (rng.random(n, dtype=f32) * f32(1e-3) + f32(0.2)).astype(f32)
        lo, hi = f32(-0.5), f32(0.5)
        if rng.random() < 0.3:
            lo, hi = f32(v.min()), f32(np.nextafter(v.max(), f32(1)))
        nleaf = int(rng.integers(2, 40))
        nb1, nb2 = int(rng.choice([16, 64, 512, 2048])), int(rng.choice([32, 256, 2048]))
        amb, cap = int(rng.choice([4, 64, 2048])), int(rng.choice([50, 1000, 40960]))
        want = literal_bisection(v, lo, hi, v.size, nleaf)
        got = select_bisection(v, lo, hi, v.size, nleaf, nb1=nb1, nb2=nb2, amb_cap=amb, cand_cap=cap)
        if got is not None:
            answered += 1
            assert same(got, want), (kind, n, nleaf, nb1, nb2, amb, cap, got, want)
        zl, zh, _ = zoom_range(v, lo, hi, nb1, v.size, nleaf)
        if np.isfinite(zl) and np.isfinite(zh) and zh > zl:
            got = select_bisection(v, lo, hi, v.size, nleaf, nb1=2048, nb2=nb2, amb_cap=amb, cand_cap=cap, bins=(zl, zh))
            if got is not None:
                answered += 1
                assert same(got, want), ("zoom", kind, n, nleaf, nb1, got, want)
    assert answered > 150


# ---------------------------------------------------------------------------------------------------------------
# Sampled rows (csrc/orb_select.cuh: SelSampleEst, sel_sample_crit, the proof in k_sel_finish / k_sel_percell) and the
# value bounds that replace the bin test in the gathering pass (sel_bin_threshold, sel_value_bounds).
# ---------------------------------------------------------------------------------------------------------------
def sample_bound(p, ns, n_tot, z, upper):
    """SelSampleEst::bound: bound on the cell's exact prefix count given the sample's prefix count p."""
    if ns <= 0:
        return n_tot if upper else 0
    sc = f32(n_tot) / f32(ns)
    pf = f32(p)
    sd = sc * f32(math.sqrt(max(float(pf * (f32(ns) - pf)), 0.0) / float(ns)))
    m = f32(z) * sd + f32(2) * sc + f32(8)
    val = pf * sc + m if upper else pf * sc - m
    return int(min(max(float(val), 0.0), float(n_tot)))


def sampled_select(v, L, R, nleaf, stride, z, nb1=1024, piece=128):
    """HIST on every stride-th piece of the cell, RESOLVE with margins, exact counts from the gathering pass, proof,
    search on the candidates.  Returns (result or None, proven)."""
    n = v.size
    prod = make_prod(n, nleaf)
    lo1, s1 = bin_params(L, R, nb1)
    idx = np.arange(n)
    smp = v[(idx // piece) % stride == 0]
    ns = smp.size
    bs = sel_bin(smp, lo1, s1, nb1) if ns else np.zeros(0, np.int64)
    ps = np.concatenate([[0], np.cumsum(np.bincount(bs, minlength=nb1))])
    # critical sample prefixes (sel_sample_crit), then two integer compares per bin
    pA = next(p for p in range(ns + 1) if diff_of(sample_bound(p, ns, n, z, True), prod) > -3)
    pB = next(p for p in range(ns, -1, -1) if diff_of(sample_bound(p, ns, n, z, False), prod) < 3)
    firsts = [b for b in range(nb1) if ps[b + 1] >= pA]
    lasts = [b for b in range(nb1) if ps[b] <= pB]
    if not firsts or not lasts or firsts[0] > lasts[-1]:
        return None, False
    f1, l1 = firsts[0], lasts[-1]
    # the gathering pass reads every particle: exact count below the candidate bins, exact candidates
    b = sel_bin(v, lo1, s1, nb1)
    base = int(np.count_nonzero(b < f1))
    cand = v[(b >= f1) & (b <= l1)]
    proven = (f1 == 0 or diff_of(base, prod) <= -3) and (l1 == nb1 - 1 or diff_of(base + cand.size, prod) >= 3)
    if not proven:
        return None, False
    # replay with exact counts base + #{cand < cut}; cuts outside the candidate bins are decided by the bins
    Lc, Rc, it, found, nleft = L, R, 0, False, None
    while it < MAX_ITER:
        cut = mid_cut(Lc, Rc)
        c1 = int(sel_bin(cut, lo1, s1, nb1))
        dec = -1 if c1 < f1 else (1 if c1 > l1 else 0)
        it += 1
        if dec == 0:
            cnt = base + int(np.count_nonzero(cand < cut))
            d = diff_of(cnt, prod)
            if abs(d) < 3:
                found, nleft = True, cnt
                break
            dec = 1 if d > 0 else -1
        if dec > 0:
            Rc = cut
        else:
            Lc = cut
    if not found:
        cut = mid_cut(Lc, Rc)
        c1 = int(sel_bin(cut, lo1, s1, nb1))
        if c1 < f1 or c1 > l1:
            return None, True
        nleft = base + int(np.count_nonzero(cand < cut))
    return (Lc, Rc, it, found, nleft), True


@pytest.mark.parametrize("stride,z", [(8, 5.0), (4, 5.0), (16, 6.0), (8, 1.0), (8, 0.0)])
def test_sampled_rows_never_change_the_result(stride, z):
    """Whatever the sample suggests, a PROVEN bracket gives the literal loop's result; an unproven one is reported (the
    GPU then searches again with exact rows).  With the default margin random-order cells are always proven."""
    rng = np.random.default_rng(23)
    proven_n = total_n = 0
    for trial in range(12):
        n = int(rng.integers(20_000, 200_000))
        kind = trial % 3
        if kind == 0:
            v = rng.random(n, dtype=f32) - f32(0.5)
        elif kind == 1:
            v = rng.normal(0.1, 0.03, n).clip(-0.5, 0.5).astype(f32)
        else:
            v = np.concatenate([rng.normal(-0.3, 0.001, n // 2), rng.random(n - n // 2) - 0.5]).astype(f32)
            rng.shuffle(v)
        nleaf = int(rng.choice([2, 3, 7, 64]))
        want = literal_bisection(v, f32(-0.5), f32(0.5), v.size, nleaf)
        got, proven = sampled_select(v, f32(-0.5), f32(0.5), nleaf, stride, z)
        total_n += 1
        proven_n += bool(proven)
        if got is not None:
            assert same(got, want), (trial, got, want)
    if z >= 5.0:
        assert proven_n == total_n


def test_sampled_rows_on_sorted_cell_are_rejected_not_wrong():
    """Positional sample of a sorted cell: the estimate is biased, the proof fails (or holds by luck) - never a wrong tree."""
    rng = np.random.default_rng(3)
    v = np.sort(rng.random(100_000, dtype=f32) - f32(0.5))
    want = literal_bisection(v, f32(-0.5), f32(0.5), v.size, 2)
    for piece in (128, 4096):
        got, proven = sampled_select(v, f32(-0.5), f32(0.5), 2, 8, 5.0, piece=piece)
        assert got is None or same(got, want)


def _key(x):
    u = np.asarray(x, f32).view(np.uint32).astype(np.uint64)
    return np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)


def _unkey(k):
    k = np.uint64(k)
    u = (k & np.uint64(0x7FFFFFFF)) if (k & np.uint64(0x80000000)) else ((~k) & np.uint64(0xFFFFFFFF))
    return np.array([u], np.uint64).astype(np.uint32).view(f32)[0]


def bin_threshold(b, lo, scale, nb):
    """sel_bin_threshold: smallest float (in the total order) whose bin is >= b, by bisection over the ordered keys."""
    ninf, pinf = f32(-np.inf), f32(np.inf)
    if int(sel_bin(ninf, lo, scale, nb)) >= b:
        return ninf
    a, e = int(_key(ninf)), int(_key(pinf))
    while e - a > 1:
        m = a + ((e - a) >> 1)
        if int(sel_bin(_unkey(m), lo, scale, nb)) >= b:
            e = m
        else:
            a = m
    return _unkey(e)


def test_value_bounds_equal_the_bin_test():
    """first <= sel_bin(v) <= last  <=>  vLo <= v < vHi for the bisected thresholds - on values around every bin edge,
    signed zeros, denormals, particles outside the box and a degenerate box."""
    rng = np.random.default_rng(11)
    for trial in range(60):
        nb = int(rng.choice([256, 512, 2048, 8192]))
        L = f32(rng.uniform(-0.5, 0.4))
        R = f32(L + f32(rng.choice([1e-6, 1e-3, 0.05, 0.9])))
        if trial % 10 == 9:
            R = L
        lo, scale = bin_params(L, R, nb)
        first = int(rng.integers(0, nb))
        last = int(rng.integers(first, nb))
        v_lo = f32(-np.inf) if first == 0 else bin_threshold(first, lo, scale, nb)
        v_hi = None if last + 1 >= nb else bin_threshold(last + 1, lo, scale, nb)
        edges = (lo + (np.arange(nb + 1, dtype=f32) / max(scale, f32(1e-30)))).astype(f32) if scale > 0 else np.array([L], f32)
        pts = np.concatenate([edges, np.nextafter(edges, f32(-np.inf)), np.nextafter(edges, f32(np.inf)),
                              rng.uniform(float(L) - 0.2, float(R) + 0.2, 4000).astype(f32),
                              np.array([0.0, -0.0, 1e-42, -1e-42, np.inf, -np.inf, 1e30, -1e30], f32)])
        b = sel_bin(pts, lo, scale, nb)
        by_bin = (b >= first) & (b <= last)
        by_val = pts >= v_lo
        if v_hi is not None:
            by_val &= ~(pts >= v_hi)
        assert np.array_equal(by_bin, by_val), (trial, nb, L, R, first, last)
        # (a bin that not even +inf reaches gets the threshold +inf: a particle AT +inf would then be neither low nor a
        #  candidate - such a bin is empty and never becomes `first`, sel_resolve_cell only picks occupied prefixes)
        fin = np.isfinite(pts)
        assert np.array_equal((b < first)[fin], (pts < v_lo)[fin])


# ---------------------------------------------------------------------------------------------------------------
# Parallel FINISH of the sampled streaming levels (csrc/orb_select.cuh: k_sel_resolve publishes the FINE bin function
# over the candidate bins' interval, k_sel_fine bins the candidates, k_sel_fin_a proves the bracket and scans the fine
# histogram, k_sel_gather collects the values of the ambiguous fine bins, k_sel_fin_b searches with the fine bins as
# the only outer histogram).
# ---------------------------------------------------------------------------------------------------------------
def parallel_finish(v, L, R, nleaf, stride, z, nb1=1024, nb2=2048, amb_cap=2048, piece=128):
    """Returns (result or None, why): None + "unproven" / "ties" is what the kernels flag for the iterative search."""
    n = v.size
    prod = make_prod(n, nleaf)
    lo1, s1 = bin_params(L, R, nb1)
    idx = np.arange(n)
    smp = v[(idx // piece) % stride == 0]
    ns = smp.size
    ps = np.concatenate([[0], np.cumsum(np.bincount(sel_bin(smp, lo1, s1, nb1), minlength=nb1))])
    pA = next(p for p in range(ns + 1) if diff_of(sample_bound(p, ns, n, z, True), prod) > -3)
    pB = next(p for p in range(ns, -1, -1) if diff_of(sample_bound(p, ns, n, z, False), prod) < 3)
    firsts = [b for b in range(nb1) if ps[b + 1] >= pA]
    lasts = [b for b in range(nb1) if ps[b] <= pB]
    if not firsts or not lasts or firsts[0] > lasts[-1]:
        return None, "unproven"
    f1, l1 = firsts[0], lasts[-1]
    # k_sel_resolve: fine bin function over [L + f1 / s1, L + (l1 + 1) / s1) - NOT the exact value bounds of the candidate
    # bins; candidates a rounding step outside it clamp to fine bin 0 / nb2 - 1, which keeps the function monotone
    if s1 > 0:
        a, bnd = f32(L + f32(f32(f1) / s1)), f32(L + f32(f32(l1 + 1) / s1))
    else:
        a, bnd = f32(0), f32(0)
    lo2, s2 = bin_params(a, bnd, nb2)
    # COMPACT: exact count below the candidate bins, the candidates themselves (membership by bin = by value bounds)
    b1 = sel_bin(v, lo1, s1, nb1)
    base = int(np.count_nonzero(b1 < f1))
    cand = v[(b1 >= f1) & (b1 <= l1)]
    K = cand.size
    # k_sel_fin_a: proof with the exact numbers, scan of the fine histogram
    if not ((f1 == 0 or diff_of(base, prod) <= -3) and (l1 == nb1 - 1 or diff_of(base + K, prod) >= 3)):
        return None, "unproven"
    b2 = sel_bin(cand, lo2, s2, nb2)
    p2 = np.concatenate([[0], np.cumsum(np.bincount(b2, minlength=nb2))])
    f2 = next((b for b in range(nb2) if diff_of(base + p2[b + 1], prod) > -3), nb2)
    l2 = next((b for b in range(nb2 - 1, -1, -1) if diff_of(base + p2[b], prod) < 3), -1)
    if f2 > l2 or p2[l2 + 1] - p2[f2] > amb_cap:
        return None, "ties"
    base2 = base + int(p2[f2])
    amb = cand[(b2 >= f2) & (b2 <= l2)]          # k_sel_gather (by the value bounds of [f2, l2]; same set)
    # k_sel_fin_b: replay; a cut is decided by its FINE bin alone unless that bin is ambiguous
    Lc, Rc, it, found, nleft = L, R, 0, False, None
    while it < MAX_ITER:
        cut = mid_cut(Lc, Rc)
        c2 = int(sel_bin(cut, lo2, s2, nb2))
        dec = -1 if c2 < f2 else (1 if c2 > l2 else 0)
        it += 1
        if dec == 0:
            cnt = base2 + int(np.count_nonzero(amb < cut))
            d = diff_of(cnt, prod)
            if abs(d) < 3:
                found, nleft = True, cnt
                break
            dec = 1 if d > 0 else -1
        if dec > 0:
            Rc = cut
        else:
            Lc = cut
    if not found:
        cut = mid_cut(Lc, Rc)
        c2 = int(sel_bin(cut, lo2, s2, nb2))
        if c2 < f2 or c2 > l2:
            return None, "final cut outside the ambiguous bins"
        nleft = base2 + int(np.count_nonzero(amb < cut))
    return (Lc, Rc, it, found, nleft), "ok"


@pytest.mark.parametrize("stride,z", [(8, 5.0), (16, 6.0), (8, 0.5)])
def test_parallel_finish_never_changes_the_result(stride, z):
    """The fine bins alone decide every cut outside [f2, l2]: correct because the proof puts diff(base) <= -3 below the
    candidate range and diff(base + K) >= 3 above it, so a cut that leaves the candidate range on either side has the
    decision of the clamped fine bin.  Whatever the sample, a result that is returned equals the literal loop's."""
    rng = np.random.default_rng(41)
    returned = 0
    for trial in range(14):
        n = int(rng.integers(30_000, 250_000))
        kind = trial % 4
        if kind == 0:
            v = rng.random(n, dtype=f32) - f32(0.5)
        elif kind == 1:
            v = rng.normal(-0.2, 0.02, n).clip(-0.5, 0.5).astype(f32)
        elif kind == 2:
            v = np.concatenate([rng.normal(0.3, 0.0005, n // 3), rng.random(n - n // 3) - 0.5]).astype(f32)
            rng.shuffle(v)
        else:                                   # a lattice: ties inside the ambiguous bins
            v = (np.floor((rng.random(n) - 0.5) * 4096) / 4096).astype(f32)
        nleaf = int(rng.choice([2, 3, 5, 64, 4097]))
        L, R = f32(-0.5), f32(0.5)
        if trial % 5 == 4:                      # margins narrower than the data (a cell deep in the tree)
            L, R = f32(-0.25), f32(0.375)
        want = literal_bisection(v, L, R, v.size, nleaf)
        got, why = parallel_finish(v, L, R, nleaf, stride, z)
        if got is not None:
            returned += 1
            assert same(got, want), (trial, why, got, want)
    assert returned >= (8 if z >= 5.0 else 1)
