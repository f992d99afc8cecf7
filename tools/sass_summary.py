"""Static evidence from the built library, no GPU needed: per kernel the ptxas resource line (registers, shared memory,
spills; from `make -B` with -Xptxas -v) and the SASS instruction mix that matters on this path (vector / async / bulk
loads, shared-memory atomics, ballots, barriers).   usage: python tools/sass_summary.py <ptxas log> <liborb_b200.so>"""
import collections
import re
import subprocess
import sys

log, lib = sys.argv[1], sys.argv[2]
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
short = lambda d: re.sub(r"\(.*", "", d).replace("void ", "").replace("orb::", "")

res = {}
for blk in open(log, errors="ignore").read().split("ptxas info    : Compiling entry function ")[1:]:
    name = blk.split("'")[1]
    regs = re.search(r"Used (\d+) registers", blk)
    smem = re.search(r"(\d+) bytes smem", blk)
    spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    res[short(demangle(name))] = (int(regs.group(1)) if regs else -1, int(smem.group(1)) if smem else 0,
                                  int(spill.group(1)) if spill else 0, int(spill.group(2)) if spill else 0)

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
mix = {}
cur = None
KEYS = [("LDG.E.128", r"\bLDG\.E\.128"), ("LDG (all)", r"\bLDG\."), ("LDGSTS", r"\bLDGSTS"), ("UBLKCP", r"\bUBLKCP"), ("STG.E.128", r"\bSTG\.E\.128"),
        ("STG (all)", r"\bSTG\."), ("ATOMS", r"\bATOMS"), ("ATOM/RED global", r"\b(ATOMG|RED|ATOM)\b|\bATOMG\.|\bRED\."), ("VOTE", r"\bVOTE"),
        ("MATCH", r"\bMATCH"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("SYNCS (mbarrier)", r"\bSYNCS"), ("instructions", r"^\s+/\*[0-9a-f]{4,}\*/")]
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = short(demangle(m.group(1)))
        mix[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, pat in KEYS:
        if re.search(pat, line):
            mix[cur][k] += 1

print(f"{'kernel':58s} {'regs':>4s} {'smem':>6s} {'spill st/ld':>11s} | " + " ".join(f"{k:>9s}" for k, _ in KEYS))
for name in sorted(set(res) | set(mix)):
    r = res.get(name, (-1, 0, 0, 0))
    c = mix.get(name, {})
    print(f"{name[:58]:58s} {r[0]:4d} {r[1]:6d} {r[2]:5d}/{r[3]:<5d} | " + " ".join(f"{c.get(k, 0):9d}" for k, _ in KEYS))
