"""GPU parity against the REAL reference: tests/golden/ref_20_10.trace.gz is the service-call trace of the unmodified
reference running `orbit 20 10 0` (BASELINE config 0, zero tie particles).  The CUDA build must reproduce, level by
level, the cells the reference's master() held after bisection (margins as bits, foundCut), the child ranges its
partition produced, and the particle multiset of every child (order-independent hashes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,x,y", [("ref_20_10.trace.gz", 20, 10), ("ref_16_8.trace.gz", 16, 8), ("ref_14_6.trace.gz", 14, 6)])
@pytest.mark.parametrize("m", [3, 1])
def test_gpu_build_equals_reference_trace(orb, oracle, name, x, y, m):
    _, levels = oracle.trace_levels(oracle.read_trace(oracle.GOLDEN / name))
    n, d = 1 << x, 1 << y
    px, py, pz = orb.generate_uniform(n)
    with orb.Orb(n, d) as ctx:
        ctx.set_trial_depth(m)
        ctx.upload(px, py, pz)
        heap, st = ctx.build()
        gx, gy, gz = ctx.download()
        rng = ctx.ranges()
    assert st.n_levels == len(levels)
    ph = oracle.particle_hash(gx, gy, gz)
    csum = np.concatenate([[np.uint64(0)], np.cumsum(ph, dtype=np.uint64)])     # wraps mod 2^64 like the reference tap
    for l, lev in enumerate(levels, start=1):
        a = (1 << (l - 1)) - 1
        ref_cells = lev["final_cells"]
        got = heap[a:a + ref_cells.size].copy()
        got["pad_"] = 0
        rc = ref_cells.copy()
        rc["pad_"] = 0
        assert got.tobytes() == rc.tobytes(), f"level {l}: cells differ from the reference"
        assert st.iters[l - 1] == len(lev["iters"]), f"level {l}: bisection iterations differ"
        # counts of the reference's last count-left call agree with the child sizes we produced (found cells)
        kids_l, kids_r = 2 * (ref_cells["id"] + 1) - 1, 2 * (ref_cells["id"] + 1)
        assert np.array_equal(rng[kids_l], lev["child_ranges"][:, 0, :]), f"level {l}: left child ranges differ"
        assert np.array_equal(rng[kids_r], lev["child_ranges"][:, 1, :]), f"level {l}: right child ranges differ"
        # particle sets per child: a child's multiset is the union of its final leaf ranges = its own final range
        with np.errstate(over="ignore"):
            for k, kids in enumerate((kids_l, kids_r)):
                b, e = rng[kids, 0].astype(np.int64), rng[kids, 1].astype(np.int64)
                assert np.array_equal(csum[e] - csum[b], lev["child_hashes"][:, k, 0]), f"level {l}: child particle sets differ"
