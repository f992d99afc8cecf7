"""Regenerates tests/golden/known_answers.json: digests of the CPU oracle's build (oracle/liborb_oracle.so, the
restatement pinned against the real reference by tests/test_oracle_vs_reference.py) for every workload bench.py and the
full-size GPU tests run: BASELINE.json configs C1..C5 at the rank counts they are sharded over, both tie modes where
the run is single-shard.  bench.py compares every build it times with this table ("parity" in its JSON line) and
tests/test_gpu_parity.py / test_gpu_multiproc.py assert the same digests.

    python tests/golden/make_known_answers.py [key ...]        # all, or only the named keys (e.g. c3_r8)

Digests (all uint64 arithmetic wraps; see tests/digests.py, shared with the checkers):
  iters, not_found   per level (the reference's j and the cells that hit the 32-iteration cap)
  heapHash           FNV-1a over the bytes of the whole Cell heap
  per shard r:
    rangeHash        SURVEY.md App. B: FNV-1a-style over the `end` index of every leaf range, id order
    leafSetHash      FNV-1a-style over (sum of particle hashes inside each leaf), id order: the north star's "final
                     particle-to-leaf assignment identical as a set per cell"
    orderHash        sum of particle_hash[i] * (2 i + 1): the particle ORDER (canonical ties: stable; hoare: 1 shard)
Particles: reference generator (uniform) or the clustered recipes of SURVEY.md §8(d), generated on the host by the
code bench.py uses (orb_generate_* in liborb_b200.so, plain C++, no GPU involved).
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import digests  # noqa: E402
import oracle_py as oracle  # noqa: E402

OUT = HERE / "known_answers.json"

# key -> (x_log2 of ALL particles, y_log2, dist, shards, ties)
WORKLOADS = {}
for ties in ("canonical", "hoare"):
    t = "" if ties == "canonical" else "_hoare"
    WORKLOADS[f"c1_r1{t}"] = (20, 10, "uniform", 1, ties)
    WORKLOADS[f"c2_r1{t}"] = (24, 12, "uniform", 1, ties)
    WORKLOADS[f"c3_r1{t}"] = (27, 16, "uniform", 1, ties)
    WORKLOADS[f"c4g_r1{t}"] = (26, 14, "gaussian", 1, ties)
    WORKLOADS[f"c4p_r1{t}"] = (26, 14, "plummer", 1, ties)
for r in (2, 4, 8):
    WORKLOADS[f"c3_r{r}"] = (27, 16, "uniform", r, "canonical")
    WORKLOADS[f"c4g_r{r}"] = (26, 14, "gaussian", r, "canonical")
    WORKLOADS[f"c4p_r{r}"] = (26, 14, "plummer", r, "canonical")
    WORKLOADS[f"c2w_r{r}"] = (24 + r.bit_length() - 1, 12, "uniform", r, "canonical")   # weak-scaled C2: 2^24 per rank
WORKLOADS["c5_r8"] = (30, 20, "uniform", 8, "canonical")
# small ones for the CPU test of this table and for smoke-sized GPU checks
WORKLOADS["s18_r1"] = (18, 8, "uniform", 1, "canonical")
WORKLOADS["s18_r2"] = (18, 8, "uniform", 2, "canonical")


def generate(x_log2, dist, shards):
    import orb_b200 as orb

    n = 1 << x_log2
    if dist == "uniform":
        return orb.generate_uniform(n)
    # one global stream: a rank slice drawn with skip = r * per equals this slice
    return orb.generate_clustered(n, dist)


def one(key):
    x_log2, y_log2, dist, shards, ties = WORKLOADS[key]
    t0 = time.time()
    x, y, z = generate(x_log2, dist, shards)
    ref = oracle.build(x, y, z, 1 << y_log2, ties=oracle.TIES_CANONICAL if ties == "canonical" else oracle.TIES_HOARE,
                       n_shards=shards, n_threads=min(shards, 8))
    del x, y, z
    st = ref["stats"]
    L = st.n_levels
    per = (1 << x_log2) // shards
    rec = {"x": x_log2, "y": y_log2, "dist": dist, "shards": shards, "ties": ties, "levels": L,
           "iters": list(st.iters[:L]), "not_found": list(st.not_found[:L]),
           "iter_particle_passes": int(st.active_passes), "tie_particles": int(st.tie_particles),
           "heapHash": digests.heap_hash(ref["heap"]), "ranks": []}
    for r in range(shards):
        sl = slice(r * per, (r + 1) * per)
        rec["ranks"].append(digests.rank_digests(ref["ranges"][r], L, ref["x"][sl], ref["y"][sl], ref["z"][sl]))
    rec["oracle_seconds"] = round(time.time() - t0, 1)
    return rec


def main():
    keys = sys.argv[1:] or list(WORKLOADS)
    table = json.loads(OUT.read_text()) if OUT.exists() else {}
    for k in keys:
        table[k] = one(k)
        print(k, table[k]["iters"], table[k]["heapHash"], f"{table[k]['oracle_seconds']} s", flush=True)
        OUT.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
