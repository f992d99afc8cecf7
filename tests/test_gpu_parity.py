"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Integer / index / byte work => the bar is bit-exact everywhere."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def level_cells(oracle, heap, l):
    a = (1 << (l - 1)) - 1
    return heap[a:a + (1 << (l - 1))].copy()


def make_level(orb, oracle, n, d, l, seed_shift=0, gen="uniform"):
    """Particles partitioned by the oracle down to level l (so the level's cells tile the array)."""
    if gen == "uniform":
        x, y, z = orb.generate_uniform(n, skip=seed_shift)
    else:
        x, y, z = orb.generate_clustered(n, gen, skip=seed_shift)
    if l == 1:
        cells = orb.root_cell(d)
        rng = np.array([[0, n]], np.uint32)
        return x, y, z, cells, rng
    r = oracle.build(x, y, z, 1 << (l - 1), ties=oracle.TIES_CANONICAL, full_levels=True)
    # oracle built l-1 full levels with d' = 2^(l-1); its leaves are the level-l cells; fix nLeafCells for d
    cells = level_cells(oracle, r["heap"], l)
    cells["nLeafCells"] = d >> (l - 1)
    a = (1 << (l - 1)) - 1
    rng = r["ranges"][0][a:a + (1 << (l - 1))]
    return r["x"], r["y"], r["z"], cells, rng


class Staged:
    """GPU context holding particles already partitioned to a level: replays the oracle's splits
    through orb.partition so the device range map matches."""

    def __init__(self, orb, oracle, n, d, l, gen="uniform"):
        self.orb, self.oracle = orb, oracle
        self.n, self.d, self.l = n, d, l
        if gen == "uniform":
            x, y, z = orb.generate_uniform(n)
        else:
            x, y, z = orb.generate_clustered(n, gen)
        self.x0, self.y0, self.z0 = x, y, z
        self.ctx = orb.Orb(n, d)
        self.ctx.upload(x, y, z)
        self.ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=True)
        for ll in range(1, l):
            self.ctx.partition(level_cells(oracle, self.ref["heap"], ll))

    def close(self):
        self.ctx.close()


@pytest.mark.parametrize("n,d,l", [(1 << 12, 16, 1), (1 << 14, 64, 3), (50_000, 256, 6), (1 << 16, 1 << 10, 9), (4099, 8, 2)])
def test_count_left_matches_oracle(orb, oracle, n, d, l):
    s = Staged(orb, oracle, n, d, l)
    try:
        cells = level_cells(oracle, s.ref["heap"], l)
        # trial cuts: fresh margins (first iteration of the level)
        for c in cells:
            c["foundCut"] = 0
        cells["cutMarginLeft"] = [c["lower"][c["cutAxis"]] for c in cells]
        cells["cutMarginRight"] = [c["upper"][c["cutAxis"]] for c in cells]
        got = s.ctx.count_left(cells)
        gx, gy, gz = s.ctx.download()
        rng = s.ctx.ranges()
        cols = (gx, gy, gz)
        for i, c in enumerate(cells):
            b, e = rng[c["id"]]
            want = oracle.count_left(cols[c["cutAxis"]], int(b), int(e), oracle.get_cut(c))
            assert got[i] == want, (i, got[i], want)
        # ServiceCount
        cnt = s.ctx.count(cells)
        assert np.array_equal(cnt, (rng[cells["id"], 1] - rng[cells["id"], 0]).astype(np.uint32))
        # found cells keep their entry (countLeft.cpp:19-21)
        cells["foundCut"][::2] = 1
        out = np.full(cells.size, 0xDEADBEEF, np.uint32)
        s.ctx.count_left(cells, out)
        assert (out[::2] == 0xDEADBEEF).all() and np.array_equal(out[1::2], got[1::2])
    finally:
        s.close()


@pytest.mark.parametrize("n,d", [(1 << 10, 4), (1 << 12, 16), (1 << 14, 64), (1 << 16, 1 << 8), (1 << 18, 1 << 10), (100_003, 32), (3000, 1 << 10)])
@pytest.mark.parametrize("m", [1, 2, 3])
def test_build_bit_exact_vs_oracle(orb, oracle, n, d, m):
    """Whole build: heap cells (cut positions as bits, margins, foundCut, axes), per-cell ranges and the
    particle order itself (the partition is stable, like the oracle's canonical mode) are identical."""
    x, y, z = orb.generate_uniform(n)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.set_trial_depth(m)
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.n_levels == ref["stats"].n_levels
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))
    assert np.array_equal(gy.view(np.uint32), ref["y"].view(np.uint32))
    assert np.array_equal(gz.view(np.uint32), ref["z"].view(np.uint32))
    assert st.active_passes > 0


@pytest.mark.parametrize("n,d", [(1 << 16, 1 << 8), (1 << 18, 1 << 10), (100_003, 32), (3000, 1 << 10), (1 << 21, 1 << 6)])
def test_build_bit_exact_with_partition_built_rows(orb, oracle, n, d, monkeypatch):
    """ORB_PREFUSE=1: from level 2 on the selection search takes its histogram rows from the previous level's
    partition (NextHist) - lean tiles, boundary tiles, one-block-per-cell levels, ragged tail."""
    monkeypatch.setenv("ORB_PREFUSE", "1")
    x, y, z = orb.generate_uniform(n)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert st.search_fallback_cells == 0
    # one read of the cut-axis column at the levels whose rows fit the histogram buffer, two elsewhere
    assert st.passes[0] == 2 and st.passes[1] == 1 and all(p <= 2 for p in st.passes[:st.n_levels]), list(st.passes[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("gen", ["gaussian", "plummer"])
@pytest.mark.parametrize("full", [False, True])
def test_build_clustered_and_full_levels(orb, oracle, gen, full, monkeypatch):
    """Also forces the partition-built histogram rows (ORB_PREFUSE=1; by default they are only used from 2^25 particles
    per GPU or with several ranks) in the full-levels runs."""
    n, d = 1 << 17, 1 << 9
    x, y, z = orb.generate_clustered(n, gen)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=full)
    if full:
        monkeypatch.setenv("ORB_PREFUSE", "1")
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build(full_levels=full)
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.n_levels == ref["stats"].n_levels
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_signed_zero_coordinates(orb, oracle):
    """-0.0 / +0.0 coordinates against cuts that are exactly +0.0 (the root's first cut is 0.0): the count kernel's
    sign-of-difference compare must agree with the CPU's `x < cut` (countLeft.cpp:35)."""
    n, d = 1 << 14, 8
    rng = np.random.default_rng(3)
    x = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
    x[::3] = np.float32(-0.0)
    x[1::7] = np.float32(0.0)
    y = np.where(rng.random(n) < 0.5, np.float32(-0.0), np.float32(0.0)).astype(np.float32)
    z = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=True)
    for m in (1, 2, 3):
        with orb.Orb(n, d) as ctx:
            ctx.set_trial_depth(m)
            ctx.upload(x, y, z)
            heap, st = ctx.build(full_levels=True)
            gx, gy, gz = ctx.download()
            rng_ = ctx.ranges()
        assert heap.tobytes() == ref["heap"].tobytes()
        assert np.array_equal(rng_, ref["ranges"][0])
        for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("stride,z,par", [("1", "6", "1"), ("8", "6", "1"), ("8", "5", "0"), ("4", "5", "1"), ("8", "0.25", "1"), ("8", "0.25", "0"), ("16", "0", "1")])
@pytest.mark.parametrize("n,d,gen", [(1 << 22, 1 << 11, "uniform"), (1 << 21, 1 << 6, "uniform"), (1 << 22, 1 << 10, "gaussian"),
                                     (1 << 21, 1 << 9, "plummer"), (3_000_017, 1 << 7, "uniform"), (1 << 23, 1 << 8, "uniform")])
def test_build_bit_exact_with_sampled_rows(orb, oracle, n, d, gen, stride, z, par, monkeypatch):
    """Sampled histogram rows (ORB_SAMPLE_STRIDE > 1, the default): HIST bins a sample, RESOLVE widens the candidate bins by
    z standard deviations, the gathering pass counts exactly and the bracket must be proven by the exact numbers.  The
    tree never depends on the sample: stride 1 (exact rows), the default, and margins far too small (z = 0.25, z = 0:
    brackets fail, k_sel_percell searches again with exact rows, k_sel_finish leaves the cell to the iterative search)
    all give the oracle's tree, ranges and particle order - streaming levels (private candidate regions, finished in
    parallel or by one block per cell) and one-block-per-cell levels."""
    monkeypatch.setenv("ORB_SAMPLE_STRIDE", stride)
    monkeypatch.setenv("ORB_SAMPLE_Z", z)
    monkeypatch.setenv("ORB_PAR_FINISH", par)                # 1: fine histogram + k_sel_fin_a / gather / fin_b, 0: k_sel_finish<true>
    monkeypatch.setenv("ORB_SAMPLE_MIN_LOCAL", "0")          # (by default only builds of >= 2^25 particles per GPU sample,
    monkeypatch.setenv("ORB_SAMPLE_MAX_AVG", str(1 << 30))   #  and only cells of <= 2^25 particles)
    x, y, z_ = orb.generate_uniform(n) if gen == "uniform" else orb.generate_clustered(n, gen)
    ref = oracle.build(x, y, z_, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z_)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    print(f"stride={stride} z={z}: passes {list(st.passes[:st.n_levels])} fallback cells {st.search_fallback_cells} ms {st.ms_total:.3f}")
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if float(z) >= 5 and gen == "uniform":
        assert st.search_fallback_cells == 0


def test_sampled_rows_on_presorted_particles(orb, oracle, monkeypatch):
    """Particles sorted along x: a positional sample of a cell is no random sample of its coordinates, the brackets
    fail, the level falls back and the build switches to exact rows - same tree as the oracle's."""
    n, d = 1 << 21, 1 << 8
    monkeypatch.setenv("ORB_SAMPLE_MIN_LOCAL", "0")
    monkeypatch.setenv("ORB_SAMPLE_MAX_AVG", str(1 << 30))
    x, y, z = orb.generate_uniform(n)
    o = np.argsort(x, kind="stable")
    x, y, z = (np.ascontiguousarray(a[o]) for a in (x, y, z))
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    print(f"presorted: passes {list(st.passes[:st.n_levels])} fallback cells {st.search_fallback_cells}")
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("n,d", [(1 << 20, 8), (1 << 16, 32)])
def test_search_fallback_on_massive_ties(orb, oracle, n, d):
    """Thousands of particles tie exactly where the cut has to go: the selection-based search cannot isolate the
    median (more ambiguous values than it keeps), flags those cells and the iterative bisection finishes them.
    n = 2^20 exercises the streaming passes (HIST / COMPACT / FINISH), n = 2^16 the cell-in-shared-memory kernel."""
    rng = np.random.default_rng(17)
    cols = []
    for a in range(3):
        v = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
        tie = rng.random(n) < 0.2                         # a fifth of the particles share one value near the median
        v[tie] = np.float32(0.01 * (a + 1))
        cols.append(v)
    x, y, z = cols
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=True)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build(full_levels=True)
        gx, gy, gz = ctx.download()
        rng_ = ctx.ranges()
    assert st.search_fallback_cells > 0
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng_, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_search_zooms_on_dense_clumps(orb, oracle):
    """Many tight clumps in a sparse background: at the levels where one block searches a whole cell (>= 64 cells), a
    cell is much wider than the clump its median falls in, so the first histogram (the row the partition built, or the
    block's own) leaves more candidates than the block stages; k_sel_percell then zooms its bin function onto the
    candidate bins instead of leaving the cell to the iterative search.  Tree, ranges and particle order stay exact,
    with the partition-built rows (default) and without (ORB_PREFUSE=0)."""
    n, d = 1 << 22, 1 << 8
    rng = np.random.default_rng(41)
    centres = rng.uniform(-0.45, 0.45, (256, 3))
    which = rng.integers(0, 256, n)
    pos = centres[which] + rng.normal(0.0, 2e-5, (n, 3))
    bg = rng.random(n) < 0.3
    pos[bg] = rng.uniform(-0.5, 0.5, (int(bg.sum()), 3))
    x, y, z = (np.ascontiguousarray(pos[:, a].clip(-0.5, 0.5).astype(np.float32)) for a in range(3))
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    for prefuse in ("1", "0"):
        os.environ["ORB_PREFUSE"] = prefuse
        try:
            with orb.Orb(n, d) as ctx:
                ctx.upload(x, y, z)
                heap, st = ctx.build()
                gx, gy, gz = ctx.download()
                rng_ = ctx.ranges()
        finally:
            del os.environ["ORB_PREFUSE"]
        assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
        assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
        assert heap.tobytes() == ref["heap"].tobytes()
        assert np.array_equal(rng_, ref["ranges"][0])
        for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        print(f"prefuse={prefuse}: passes {list(st.passes[:st.n_levels])} not_found {list(st.not_found[:st.n_levels])} "
              f"fallback cells {st.search_fallback_cells}")


def test_search_without_fallback_on_plain_inputs(orb, oracle):
    """Uniform particles never need the iterative path; ORB_SELECT=0 (iterative search everywhere) gives the same tree."""
    n, d = 1 << 19, 1 << 7
    x, y, z = orb.generate_uniform(n)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
    assert st.search_fallback_cells == 0
    assert list(st.passes[:st.n_levels]) == [2] * 3 + [1] * 3 or all(p <= 2 for p in st.passes[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    os.environ["ORB_SELECT"] = "0"
    try:
        with orb.Orb(n, d) as ctx:
            ctx.upload(x, y, z)
            heap2, st2 = ctx.build()
    finally:
        del os.environ["ORB_SELECT"]
    assert heap2.tobytes() == ref["heap"].tobytes()
    assert max(st2.passes[:st2.n_levels]) > 2


def test_degenerate_inputs(orb, oracle):
    """All-equal coordinates (never converges: 32-iteration cap + extra count), duplicates on the cut."""
    n, d = 1 << 13, 16
    x = np.full(n, 0.25, np.float32)
    y = np.linspace(-0.5, 0.5, n, dtype=np.float32)
    z = np.repeat(np.float32([-0.3, 0.0, 0.0, 0.3]), n // 4)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=True)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build(full_levels=True)
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert list(st.iters[:st.n_levels]) == list(ref["stats"].iters[:st.n_levels])
    assert list(st.not_found[:st.n_levels]) == list(ref["stats"].not_found[:st.n_levels])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("n,d,l", [(1 << 14, 64, 4), (70_001, 512, 8)])
def test_find_cuts_and_partition_services(orb, oracle, n, d, l):
    """Fused per-level bisection + the partition service, against the oracle level by level."""
    s = Staged(orb, oracle, n, d, l)
    try:
        cells = level_cells(oracle, s.ref["heap"], l)
        fresh = cells.copy()
        fresh["foundCut"] = 0
        fresh["cutMarginLeft"] = [c["lower"][c["cutAxis"]] for c in cells]
        fresh["cutMarginRight"] = [c["upper"][c["cutAxis"]] for c in cells]
        s.ctx.count(fresh)
        out, iters, passes = s.ctx.find_cuts(fresh)
        assert iters == s.ref["stats"].iters[l - 1]
        assert out.tobytes() == cells.tobytes()     # oracle heap holds the post-bisection margins / foundCut
        s.ctx.partition(out)
        kids = level_cells(oracle, s.ref["heap"], l + 1)
        rng = s.ctx.ranges()
        assert np.array_equal(rng[kids["id"]], s.ref["ranges"][0][kids["id"]])
    finally:
        s.close()


def test_bbox_matches_numpy(orb, oracle):
    n, d, l = 60_000, 128, 6
    s = Staged(orb, oracle, n, d, l, gen="gaussian")
    try:
        cells = level_cells(oracle, s.ref["heap"], l)
        got = s.ctx.bbox(cells)
        gx, gy, gz = s.ctx.download()
        rng = s.ctx.ranges()
        for i, c in enumerate(cells):
            b, e = (int(v) for v in rng[c["id"]])
            want = oracle.bbox(gx, gy, gz, b, e)
            assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (i, got[i], want)
    finally:
        s.close()


def test_tight_box_mode_vs_oracle(orb, oracle):
    n, d = 1 << 15, 64
    x, y, z = orb.generate_clustered(n, "gaussian")
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, tight_box=True)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build(tight_box=True)
        rng = ctx.ranges()
        gx, gy, gz = ctx.download()
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))


def test_range_contract_violation_fails_loudly(orb):
    n, d = 4096, 8
    x, y, z = orb.generate_uniform(n)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        cells = np.zeros(2, orb.CELL_DTYPE)
        cells["id"] = [1, 2]          # level-2 cells whose ranges were never produced by a partition
        cells["nLeafCells"] = 4
        cells["lower"] = -0.5
        cells["upper"] = 0.5
        cells["cutMarginLeft"] = -0.5
        cells["cutMarginRight"] = 0.5
        with pytest.raises(orb.OrbError):
            ctx.count_left(cells)


def test_full_size_properties_c2(orb, oracle):
    """BASELINE config 2 size (2^24 particles, 2^12 leaf cells) through size-independent properties:
    children partition parents, left < cut <= right on the cut axis, particle multiset conserved
    (checksum of per-particle hashes), level counts match the reference's published iteration pattern."""
    n, d = 1 << 24, 1 << 12
    x, y, z = orb.generate_uniform(n)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    # reference trace of `orbit 24 12 0`: iterations per level (identical in canonical mode
    # until the first tie at level 4 changes child sizes by one particle - the counts stay equal here)
    assert st.n_levels == 11
    assert list(st.iters[:3]) == [21, 22, 20]
    assert oracle.set_hash(gx, gy, gz) == oracle.set_hash(x, y, z)
    cols = (gx, gy, gz)
    for l in range(1, st.n_levels + 1):
        a = (1 << (l - 1)) - 1
        for c in heap[a:a + (1 << (l - 1))][:: max(1, (1 << (l - 1)) // 64)]:
            lid, rid = 2 * (c["id"] + 1) - 1, 2 * (c["id"] + 1)
            assert rng[lid][0] == rng[c["id"]][0] and rng[lid][1] == rng[rid][0] and rng[rid][1] == rng[c["id"]][1]
            cut = oracle.get_cut(c)
            col = cols[c["cutAxis"]]
            assert (col[rng[lid][0]:rng[lid][1]] < cut).all()
            assert (col[rng[rid][0]:rng[rid][1]] >= cut).all()
    leaves = heap[(1 << st.n_levels) - 1:(1 << (st.n_levels + 1)) - 1]
    sizes = rng[leaves["id"], 1] - rng[leaves["id"], 0]
    assert sizes.sum() == n and sizes.min() >= 8188 - 8 and sizes.max() <= 8196 + 8
    _check_known_answers("c2_r1", heap, st, rng, gx, gy, gz)


def _check_known_answers(key, heap, st, rng, gx, gy, gz):
    """bit-exact digests of the CPU oracle's build of the same workload (tests/golden/known_answers.json): iterations,
    capped cells, the whole cell heap, leaf ranges, per-leaf particle sets and the particle order"""
    import digests

    rec = digests.load_known()[key]
    L = st.n_levels
    res = digests.compare(rec, 0, iters=st.iters[:L], not_found=st.not_found[:L], heap=heap, ranges=rng, n_levels=L, x=gx, y=gy, z=gz)
    assert res["ok"], (key, res["mismatch"])


def _check_build_properties(orb, oracle, x, y, z, d, heap, rng, gx, gy, gz, n_levels, sample=48):
    n = x.size
    assert oracle.set_hash(gx, gy, gz) == oracle.set_hash(x, y, z)          # multiset of particles conserved
    cols = (gx, gy, gz)
    for l in range(1, n_levels + 1):
        a = (1 << (l - 1)) - 1
        cells = heap[a:a + (1 << (l - 1))]
        for c in cells[:: max(1, cells.size // sample)]:
            lid, rid = 2 * (c["id"] + 1) - 1, 2 * (c["id"] + 1)
            assert rng[lid][0] == rng[c["id"]][0] and rng[lid][1] == rng[rid][0] and rng[rid][1] == rng[c["id"]][1]
            cut = oracle.get_cut(c)
            col = cols[c["cutAxis"]]
            assert (col[rng[lid][0]:rng[lid][1]] < cut).all()
            assert (col[rng[rid][0]:rng[rid][1]] >= cut).all()
    leaves = heap[(1 << n_levels) - 1:(1 << (n_levels + 1)) - 1]
    sizes = rng[leaves["id"], 1].astype(np.int64) - rng[leaves["id"], 0]
    assert sizes.sum() == n
    return sizes


def test_full_size_properties_c3(orb, oracle):
    """BASELINE config 2 per-GPU size class: 2^27 uniform particles, 2^16 leaf cells (beyond the reference's MAX_CELLS):
    conservation, child partition of parents, left < cut <= right, iteration pattern of SURVEY.md Appendix B."""
    n, d = 1 << 27, 1 << 16
    x, y, z = orb.generate_uniform(n)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.n_levels == 15
    # known answers from the canonical-mode oracle run here on the CPU (`oracle/orb_oracle 27 16 --ties=canonical`,
    # 42 s): iterations per level, one level-4 cell hits the 32-iteration cap (455 tie particles: the generator's
    # values sit on a 2^-25 grid, duplicates make |difference| < 3 unreachable there), range hash of the leaf level.
    # (The verbatim-Hoare run of SURVEY.md Appendix B has 23 iterations at level 4: its tie particles fall differently.)
    assert list(st.iters[:15]) == [17, 25, 23, 32, 22, 21, 20, 20, 19, 18, 17, 17, 16, 15, 15]
    assert list(st.not_found[:15]) == [0, 0, 0, 1] + [0] * 11
    h = 1469598103934665603
    for i in range((1 << 15) - 1, (1 << 16) - 1):
        h = ((h ^ int(rng[i][1])) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert h == 0xD4F00F38C84CA6F5
    _check_build_properties(orb, oracle, x, y, z, d, heap, rng, gx, gy, gz, st.n_levels)
    _check_known_answers("c3_r1", heap, st, rng, gx, gy, gz)


@pytest.mark.parametrize("kind", ["gaussian", "plummer"])
def test_full_size_properties_c4_clustered(orb, oracle, kind):
    """BASELINE config 3: 2^26 clustered particles, 2^14 leaf cells: skewed cells, many iterations, empty regions."""
    n, d = 1 << 26, 1 << 14
    x, y, z = orb.generate_clustered(n, kind)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.n_levels == 13
    sizes = _check_build_properties(orb, oracle, x, y, z, d, heap, rng, gx, gy, gz, st.n_levels)
    # the reference's |difference| < 3 rule balances every split to within a few particles unless a cell hit the cap
    if sum(st.not_found[:13]) == 0:
        assert sizes.max() - sizes.min() <= 64
    # ... and bit for bit what the CPU oracle builds from the same 2^26 clustered particles
    _check_known_answers("c4g_r1" if kind == "gaussian" else "c4p_r1", heap, st, rng, gx, gy, gz)


def tie_columns(n, seed=17):
    """A fifth of the particles share one value near the median on every axis: the selection search must flag cells."""
    rng = np.random.default_rng(seed)
    cols = []
    for a in range(3):
        v = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
        tie = rng.random(n) < 0.2
        v[tie] = np.float32(0.01 * (a + 1))
        cols.append(v)
    return cols


@pytest.mark.parametrize("n,d,gen", [(1 << 16, 1 << 8, "uniform"), (1 << 22, 1 << 13, "uniform"), (1 << 20, 1 << 14, "uniform"),
                                     (100_003, 1 << 11, "uniform"), (1 << 21, 1 << 12, "plummer"), (1 << 20, 1 << 11, "ties"),
                                     (1 << 23, 1 << 4, "uniform")])
def test_multi_rank_protocol_against_itself(orb, oracle, n, d, gen, monkeypatch):
    """ORB_MR_SELF=1: one rank runs the whole multi-rank protocol of orb_exchange.cuh (REDUCE, owner FINISH, APPLY; streaming,
    block-per-cell and warp-per-cell levels) with itself as the only peer - every kernel of the exchange path on one GPU."""
    monkeypatch.setenv("ORB_MR_SELF", "1")
    if gen == "uniform":
        x, y, z = orb.generate_uniform(n)
    elif gen == "ties":
        x, y, z = tie_columns(n)
    else:
        x, y, z = orb.generate_clustered(n, gen)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    L = st.n_levels
    assert list(st.iters[:L]) == list(ref["stats"].iters[:L])
    assert list(st.not_found[:L]) == list(ref["stats"].not_found[:L])
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # uniform inputs never fall back; massive ties always do; dense Plummer cores may (a bin of the per-cell regimes can
    # hold more candidates than the owner stages - the result is exact either way)
    if gen == "uniform":
        assert st.search_fallback_cells == 0
    elif gen == "ties":
        assert st.search_fallback_cells > 0


def test_multi_rank_owner_search_in_global_scratch(orb, oracle, monkeypatch):
    """Cells with more candidates than an owner block stages in shared memory (the root of a 2^30-particle build: 131072
    particles per bin) are searched in global scratch; here the shared-memory capacity is turned down to force that path."""
    monkeypatch.setenv("ORB_MR_SELF", "1")
    monkeypatch.setenv("ORB_X_CAND_CAP", "1024")
    n, d = 1 << 22, 1 << 5
    x, y, z = orb.generate_uniform(n)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.search_fallback_cells == 0
    assert heap.tobytes() == ref["heap"].tobytes() and np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("n,d", [(1 << 21, 1 << 6), (100_003, 32)])
def test_partition_bulk_copy_path(orb, oracle, n, d, monkeypatch):
    """ORB_PART_BULK=1: the cooperative partition loads its tiles with cp.async.bulk + mbarrier (the default is per-thread
    cp.async, which measured faster - profiles/r02h_partition_bulk_ab.txt); same bits either way."""
    monkeypatch.setenv("ORB_PART_BULK", "1")
    x, y, z = orb.generate_uniform(n)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    with orb.Orb(n, d) as ctx:
        ctx.upload(x, y, z)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert heap.tobytes() == ref["heap"].tobytes() and np.array_equal(rng, ref["ranges"][0])
    for a, b in ((gx, ref["x"]), (gy, ref["y"]), (gz, ref["z"])):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
