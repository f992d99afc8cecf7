"""Memory-side view of an `ncu --set full` raw page (CSV): one line per launch with the achieved DRAM rate (ncu's own
counters over ncu's own duration: cold cache, one kernel at a time), DRAM / L2 / L1 / SM throughput as % of ncu's peaks and
the L2 / L1 sector hit rates.   usage: python tools/ncu_memory_view.py <raw.csv>"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    i = col[name]
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[i]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6,
              "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}.get(u, 1.0)
    return v


print(f"{'#':>3s} {'kernel':34s} {'us':>7s} {'DRAM GB/s':>9s} {'dram%':>6s} {'L2%':>6s} {'L1%':>6s} {'SM%':>6s} {'L2 hit%':>7s} {'L1 hit%':>7s}")
for k, r in enumerate(data):
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("orb::", "")
    t = val(r, "gpu__time_duration.sum")
    by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    print(f"{k:3d} {name[:34]:34s} {t * 1e6:7.1f} {by / t / 1e9:9.0f} {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {val(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {val(r, 'lts__t_sector_hit_rate.pct'):7.1f} {val(r, 'l1tex__t_sector_hit_rate.pct'):7.1f}")
