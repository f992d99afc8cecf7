"""Warm-up builds and measured builds of 2^x particles into 2^y leaf cells (for ncu launch lists and knob sweeps).
usage: build_once.py [x] [y] [reps] [dist]   (prints median / min build ms over `reps` builds after one warm-up)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orb_b200 as orb
x_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
y_log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dist = sys.argv[4] if len(sys.argv) > 4 else "uniform"
n, d = 1 << x_log2, 1 << y_log2
x, y, z = orb.generate_uniform(n) if dist == "uniform" else orb.generate_clustered(n, dist)
ctx = orb.Orb(n, d)
ms = []
for rep in range(1 + reps):
    ctx.upload(x, y, z)
    heap, st = ctx.build()
    if rep:
        ms.append(st.ms_total)
print("ms median %.4f min %.4f" % ((float(np.median(ms)), min(ms)) if ms else (0.0, 0.0)), "passes", list(st.passes[:st.n_levels]),
      "fallback_cells", st.search_fallback_cells)
