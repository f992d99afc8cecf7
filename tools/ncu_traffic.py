"""DRAM traffic per launch of the two kernel groups of a build, from an ncu capture of that very config.

    # on the GPU box (one build after one warm-up build; -s skips the warm-up's launches):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/r02_traffic_c3.csv python tools/build_once.py 27 16 1
    # here:
    python tools/ncu_traffic.py c3 27 gpurun_out/r02_traffic_c3.csv      # -> profiles/r02_ncu_traffic.json["c3"]

Groups (the ones bench.py reports): k_count = every particle-streaming kernel of the cut search (k_sel_stream, k_sel_percell,
k_xd_*, k_count_*, k_level_persistent); k_partition = k_partition_*.  The LAST build's launches are used (the capture holds
warm-up + measured builds back to back; a build starts at the first launch after a k_partition_* whose successor is a
level-1 kernel - simpler: the last `launches_per_build` launches, taken from the build's own launch count).
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COUNT = ("k_sel_stream", "k_sel_percell", "k_xd_hist", "k_xd_compact", "k_count_stream", "k_count_cells", "k_level_persistent")
PART = ("k_partition_coop", "k_partition_cells", "k_partition_warp")


def main():
    cfg, x_log2, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    n_builds = int(sys.argv[4]) if len(sys.argv) > 4 else 2       # warm-up + measured
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    iN, iM, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        try:
            lid = int(r[iID])
        except ValueError:
            continue
        d = launches.setdefault(lid, {"name": r[iN]})
        try:
            d[r[iM]] = float(r[iV].replace(",", ""))
        except ValueError:
            pass
    ids = sorted(launches)
    per_build = len(ids) // n_builds
    last = [launches[i] for i in ids[-per_build:]]
    n = 1 << x_log2
    out = {}
    for group, names, alg in (("k_count", COUNT, 4 * n), ("k_partition", PART, 24 * n)):
        sel = [l for l in last if any(l["name"].startswith(k) or (" " + k) in l["name"] or l["name"].split("(")[0].endswith(k) or k in l["name"].split("<")[0] for k in names)]
        if not sel:
            continue
        rd = sum(l.get("dram__bytes_read.sum", 0.0) for l in sel) / len(sel)
        wr = sum(l.get("dram__bytes_write.sum", 0.0) for l in sel) / len(sel)
        us = sum(l.get("gpu__time_duration.sum", 0.0) for l in sel) / len(sel)
        out[group] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "launches_captured": len(sel),
                      "algorithmic_bytes_per_launch": alg, "traffic_over_algorithmic": (rd + wr) / alg,
                      "duration_us_under_ncu_mean": us / 1e3 if us > 1e4 else us,
                      "note": f"mean over the {len(sel)} {group} launches of one {cfg} build under ncu (dram__bytes_read.sum + dram__bytes_write.sum, "
                              f"--clock-control none; {Path(path).name}); algorithmic bytes of one launch: {alg}"}
    dst = ROOT / "profiles" / "r02_ncu_traffic.json"
    table = json.loads(dst.read_text()) if dst.exists() else {}
    table[cfg] = out
    dst.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
