"""CPU tests: the oracle (oracle/orb_oracle.c) against golden traces produced by the UNMODIFIED
reference (tests/golden/make_golden.py), plus the survey's known answers.  Byte-exact."""
import gzip
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(GOLDEN.glob("ref_*.trace.gz"))


def parse_name(p):
    stem = p.name.split(".")[0]          # ref_10_3p
    _, x, y = stem.split("_")
    parts = y.endswith("p")
    return int(x), int(y.rstrip("p")), parts


@pytest.mark.parametrize("path", CASES, ids=[p.name for p in CASES])
def test_oracle_trace_equals_reference_golden(oracle, tmp_path, path):
    """Every service call the reference made (cells, margins, foundCut, counts per bisection iteration,
    child ranges, ordered particle hashes, particle dumps) is reproduced byte for byte."""
    x, y, parts = parse_name(path)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    out = tmp_path / "o.trace"
    oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_HOARE, trace_path=out, trace_particles=parts)
    want = gzip.open(path, "rb").read()
    got = out.read_bytes()
    assert len(got) == len(want)
    assert got == want


def test_generator_known_answers(oracle, orb):
    """first six outputs of xorshf96 (init.cu:11-25), product and oracle generators."""
    want = [0xBEFFFFED, 0xBEFFFF92, 0xBEFFFFC1, 0xBECB9B7B, 0xBE70705E, 0xBCD1F110]
    for gen in (oracle.generate_uniform, orb.generate_uniform):
        x, y, z = gen(2)
        got = [int(v.view(np.uint32)[i]) for i in range(2) for v in (x, y, z)]
        assert got == want
    # product generator with skip == slice of the single stream
    x, y, z = orb.generate_uniform(1000)
    xs, ys, zs = orb.generate_uniform(300, skip=700)
    assert np.array_equal(xs, x[700:]) and np.array_equal(ys, y[700:]) and np.array_equal(zs, z[700:])


def test_c1_known_answers(oracle):
    """`orbit 20 10 0` (BASELINE config 0): per-level iterations, first-cell cuts and left sizes, range hash."""
    x, y, z = oracle.generate_uniform(1 << 20)
    r = oracle.build(x, y, z, 1 << 10, ties=oracle.TIES_HOARE)
    st = r["stats"]
    assert st.n_levels == 9
    assert list(st.iters[:9]) == [18, 17, 16, 15, 14, 14, 14, 13, 12]
    assert st.tie_particles == 0 and sum(st.not_found[:9]) == 0
    assert abs(st.active_passes / (1 << 20) - 116.8164) < 1e-3
    cuts = [0xB9240000, 0x39980000, 0x3ABE0000, 0xBE7EC1A1, 0xBE801506, 0xBE7FD485, 0xBEC0512E, 0xBEBF3501, 0xBEC14A4E]
    lefts = [524288, 262143, 131069, 65536, 32767, 16385, 8194, 4098, 2051]
    for l in range(1, 10):
        c = r["heap"][(1 << (l - 1)) - 1]
        assert int(oracle.get_cut(c).view(np.uint32)) == cuts[l - 1]
        lid = 2 * (c["id"] + 1) - 1
        assert int(r["ranges"][0][lid][1] - r["ranges"][0][lid][0]) == lefts[l - 1]
    h = 1469598103934665603
    for i in range((1 << 9) - 1, (1 << 10) - 1):
        h = ((h ^ int(r["ranges"][0][i][1])) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert h == 0x192258D9865BFADD
    # zero ties => the canonical (stable x<cut) mode the GPU implements gives the same cells and ranges
    rc = oracle.build(x, y, z, 1 << 10, ties=oracle.TIES_CANONICAL)
    assert rc["heap"].tobytes() == r["heap"].tobytes()
    assert np.array_equal(rc["ranges"], r["ranges"])


def test_sharded_counts_are_shard_invariant(oracle):
    """Cuts/cells are identical for every shard count (integer sums are order independent) and every
    thread count; per-shard ranges add up to the single-shard child sizes."""
    n, d = 1 << 16, 1 << 6
    x, y, z = oracle.generate_uniform(n)
    base = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    for shards, threads in ((2, 1), (4, 4), (8, 3)):
        r = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, n_shards=shards, n_threads=threads)
        assert r["heap"].tobytes() == base["heap"].tobytes()
        sizes = (r["ranges"][:, :, 1].astype(np.int64) - r["ranges"][:, :, 0]).sum(axis=0)
        bsz = base["ranges"][0][:, 1].astype(np.int64) - base["ranges"][0][:, 0]
        assert np.array_equal(sizes, bsz)
        assert list(r["stats"].iters[:6]) == list(base["stats"].iters[:6])


def test_canonical_partition_contract(oracle):
    """Left child strictly below the cut, right child >= cut, stable order, multiset conserved."""
    n, d = 20_000, 32
    rng = np.random.default_rng(7)
    x = rng.integers(-8, 8, n).astype(np.float32) / 16    # many exact ties on cut positions
    y = rng.random(n, dtype=np.float32) - 0.5
    z = rng.random(n, dtype=np.float32) - 0.5
    r = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL, full_levels=True)
    assert r["stats"].tie_particles > 0
    assert oracle.set_hash(r["x"], r["y"], r["z"]) == oracle.set_hash(x, y, z)
    cols = (r["x"], r["y"], r["z"])
    for c in r["heap"][: d - 1]:
        lid, rid = 2 * (c["id"] + 1) - 1, 2 * (c["id"] + 1)
        (lb, le), (rb, re) = r["ranges"][0][lid], r["ranges"][0][rid]
        cut = oracle.get_cut(c)
        assert (cols[c["cutAxis"]][lb:le] < cut).all() and (cols[c["cutAxis"]][rb:re] >= cut).all()
        assert le == rb and lb == r["ranges"][0][c["id"]][0] and re == r["ranges"][0][c["id"]][1]
