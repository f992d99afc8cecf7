"""Time orb_bbox (k_bbox: per-cell min/max of x,y,z, 12 B per particle) on the cells of a few levels of a built tree.
usage: bbox_once.py [x] [y]   - wall clock around the synchronous C-ABI call (launch + 24 B per cell back), best of 7"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import orb_b200 as orb
x_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 27
y_log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n, d = 1 << x_log2, 1 << y_log2
x, y, z = orb.generate_uniform(n)
ctx = orb.Orb(n, d)
ctx.upload(x, y, z)
heap, st = ctx.build()
for level in (1, 4, 8, 12, y_log2 - 1):
    if level > st.n_levels:
        continue
    a, cnt = (1 << (level - 1)) - 1, 1 << (level - 1)
    cells = heap[a:a + cnt].copy()
    best = 1e9
    for rep in range(8):
        t0 = time.perf_counter()
        bb = ctx.bbox(cells)
        dt = time.perf_counter() - t0
        if rep:
            best = min(best, dt)
    gbs = 12.0 * n / best / 1e9
    ok = bool((bb[:, 0] <= bb[:, 3]).all())
    print(f"level {level:2d} cells {cnt:6d}: {best * 1e6:8.1f} us  {gbs:7.1f} GB/s algorithmic (12 B/particle)  boxes ordered {ok}")
