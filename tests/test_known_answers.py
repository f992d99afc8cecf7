"""The committed digest table tests/golden/known_answers.json (made by tests/golden/make_known_answers.py from the CPU
oracle) is what bench.py and the full-size GPU tests compare the CUDA path with.  Here: the oracle reproduces the small
entries, the table covers every workload bench.py runs by default, and the digests see what they are meant to see."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import digests


def test_oracle_reproduces_small_entries(orb, oracle):
    import make_known_answers as mk

    table = digests.load_known()
    for key in ("s18_r1", "s18_r2", "c1_r1", "c1_r1_hoare"):
        rec = mk.one(key)
        rec.pop("oracle_seconds")
        want = dict(table[key])
        want.pop("oracle_seconds")
        assert rec == want, key


def test_table_covers_the_default_bench_legs():
    sys.path.insert(0, str(ROOT))
    import bench

    table = digests.load_known()
    for n in (1, 2, 4, 8):
        for leg in bench.default_legs(n):
            rec = table[digests.known_key(leg, n)]
            cfg = bench.config_of(leg, n)
            assert (rec["x"], rec["y"], rec["dist"], rec["shards"]) == (cfg["x"], cfg["y"], cfg["dist"], n)
            assert len(rec["ranks"]) == n and rec["levels"] == cfg["y"] - 1 == len(rec["iters"])
    for leg in ("c2", "c3"):
        assert table[digests.known_key(leg, 1, "hoare")]["ties"] == "hoare"
    # the reference's own numbers of SURVEY.md App. B are in the table's C2 entries: 18 tie particles with Hoare
    assert table["c2_r1_hoare"]["tie_particles"] == 18 and table["c1_r1"]["tie_particles"] == 0
    # sharding changes neither cuts nor iterations, only the per-rank ranges
    assert table["c3_r8"]["heapHash"] == table["c3_r1"]["heapHash"] and table["c3_r8"]["iters"] == table["c3_r1"]["iters"]


def test_digests_see_sets_and_order(orb, oracle):
    x, y, z = orb.generate_uniform(1 << 14)
    ref = oracle.build(x, y, z, 1 << 5, ties=oracle.TIES_CANONICAL)
    L = ref["stats"].n_levels
    base = digests.rank_digests(ref["ranges"][0], L, ref["x"], ref["y"], ref["z"])
    assert base["leaves_tile_slice"]
    # permuting two particles inside one leaf keeps the set digest, changes the order digest
    ids = digests.leaf_ids(L)
    b, e = (int(v) for v in ref["ranges"][0][ids[3]])
    assert e - b >= 2
    px, py, pz = ref["x"].copy(), ref["y"].copy(), ref["z"].copy()
    for a in (px, py, pz):
        a[[b, b + 1]] = a[[b + 1, b]]
    perm = digests.rank_digests(ref["ranges"][0], L, px, py, pz)
    assert perm["leafSetHash"] == base["leafSetHash"] and perm["orderHash"] != base["orderHash"] and perm["rangeHash"] == base["rangeHash"]
    # moving a particle into the neighbouring leaf changes the set digest
    b2 = int(ref["ranges"][0][ids[4]][0])
    for a in (px, py, pz):
        a[[b, b2]] = a[[b2, b]]
    moved = digests.rank_digests(ref["ranges"][0], L, px, py, pz)
    assert moved["leafSetHash"] != base["leafSetHash"]
    # a range map that does not tile the slice is reported
    rng = ref["ranges"][0].copy()
    rng[ids[2], 1] += 1
    assert not digests.rank_digests(rng, L, ref["x"], ref["y"], ref["z"])["leaves_tile_slice"]
    # compare(): the record of this very build passes, a changed iteration count does not
    rec = {"iters": list(ref["stats"].iters[:L]), "not_found": list(ref["stats"].not_found[:L]), "heapHash": digests.heap_hash(ref["heap"]),
           "ranks": [base]}
    args = dict(iters=rec["iters"], not_found=rec["not_found"], heap=ref["heap"], ranges=ref["ranges"][0], n_levels=L, x=ref["x"], y=ref["y"], z=ref["z"])
    assert digests.compare(rec, 0, **args)["ok"]
    args["iters"] = rec["iters"][:-1] + [rec["iters"][-1] + 1]
    assert digests.compare(rec, 0, **args)["mismatch"] == ["iters"]
