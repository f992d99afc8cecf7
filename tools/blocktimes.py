import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import os
os.environ["ORB_DEBUG_TIMES"] = "2"
os.environ["ORB_DEBUG_TIMES_FILE"] = "gpurun_out/orb_block_times.bin"
import orb_b200 as orb
n, d = 1 << 24, 1 << 12
x, y, z = orb.generate_uniform(n)
ctx = orb.Orb(n, d)
for rep in range(3):
    ctx.upload(x, y, z)
    heap, st = ctx.build()
print("ms", st.ms_total, "passes", list(st.passes[:st.n_levels]))
raw = open("gpurun_out/orb_block_times.bin", "rb").read()
off = 0
while off < len(raw):
    lvl, g = struct.unpack_from("<II", raw, off); off += 8
    t = np.frombuffer(raw, "<u8", 12 * g * 4, off).reshape(12, g, 4).astype(np.int64); off += 12 * g * 4 * 8
    for p in range(12):
        a = t[p]
        if a[:, 0].max() == 0: break
        ok = a[:, 0] > 0
        t0 = a[ok, 0].min()
        r = (a[ok] - t0) / 1e3
        print(f"L{lvl} p{p} g{g}: start med {np.median(r[:,0]):.1f} max {r[:,0].max():.1f} | classified med {np.median(r[:,1]):.1f} max {r[:,1].max():.1f} | streamed med {np.median(r[:,2]):.1f} p90 {np.percentile(r[:,2],90):.1f} max {r[:,2].max():.1f} | flushed med {np.median(r[:,3]):.1f} p90 {np.percentile(r[:,3],90):.1f} max {r[:,3].max():.1f} argmax {int(np.argmax(r[:,3]))}")
    if lvl >= 4: break
