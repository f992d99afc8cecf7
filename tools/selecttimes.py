"""ORB_DEBUG_TIMES=2 dump of the selection search: per-block phase stamps of HIST / COMPACT, phases of FINISH block 0."""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ORB_DEBUG_TIMES"] = "2"
os.environ["ORB_DEBUG_TIMES_FILE"] = "gpurun_out/orb_block_times.bin"
import orb_b200 as orb
n, d = 1 << 24, 1 << 12
x, y, z = orb.generate_uniform(n)
ctx = orb.Orb(n, d)
for rep in range(3):
    ctx.upload(x, y, z)
    heap, st = ctx.build()
print("ms", st.ms_total)
raw = open("gpurun_out/orb_block_times.bin", "rb").read()
off = 0
while off < len(raw):
    lvl, g = struct.unpack_from("<II", raw, off); off += 8
    t = np.frombuffer(raw, "<u8", 12 * g * 4, off).reshape(12, g, 4).astype(np.int64); off += 12 * g * 4 * 8
    for p, name in ((0, "hist"), (1, "compact")):
        a = t[p]; ok = a[:, 0] > 0
        if not ok.any(): continue
        t0 = a[ok, 0].min(); r = (a[ok] - t0) / 1e3
        print(f"L{lvl} {name:7s} blocks {ok.sum()}: start max {r[:,0].max():.1f} | classified med {np.median(r[:,1]):.1f} max {r[:,1].max():.1f} | streamed med {np.median(r[:,2]):.1f} p90 {np.percentile(r[:,2],90):.1f} max {r[:,2].max():.1f} | flushed med {np.median(r[:,3]):.1f} p90 {np.percentile(r[:,3],90):.1f} max {r[:,3].max():.1f}")
    f = t[2].reshape(-1)[:8]
    if f[0]:
        hist_end = t[0][t[0][:, 0] > 0][:, 3].max(); comp = t[1][t[1][:, 0] > 0]
        print(f"L{lvl} gaps: hist end -> compact start {(comp[:,0].min()-hist_end)/1e3:.1f} us; compact end -> finish start {(f[0]-comp[:,3].max())/1e3:.1f} us")
        print(f"L{lvl} finish block0: " + " ".join(f"{(f[i]-f[0])/1e3:.1f}" for i in range(1, 8)) + "  (staged, minmax, hist2, scanned, gathered, replayed, end)")
    if lvl >= 9: break
