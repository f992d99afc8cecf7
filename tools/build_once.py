"""One warm-up build and one measured build of 2^x uniform particles into 2^y leaf cells (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orb_b200 as orb
x_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
y_log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n, d = 1 << x_log2, 1 << y_log2
x, y, z = orb.generate_uniform(n)
ctx = orb.Orb(n, d)
for rep in range(2):
    ctx.upload(x, y, z)
    heap, st = ctx.build()
print("ms", st.ms_total, "passes", list(st.passes[:st.n_levels]))
