"""One line per captured launch of an `ncu --page raw --csv` export: time, DRAM bytes, issue rate, occupancy, shared-memory
atomics / bank conflicts, top stall reasons.   usage: python tools/ncu_table.py <raw.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rdMB", "dram__bytes_write.sum": "wrMB",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue%", "smsp__inst_executed.sum": "Minst",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%", "launch__grid_size": "grid", "launch__block_size": "blk",
        "launch__registers_per_thread": "regs", "launch__occupancy_limit_shared_mem": "occSm",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bankconf", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shwave",
        "smsp__inst_executed_op_shared_atom.sum": "shatom", "launch__shared_mem_per_block_dynamic": "dsmemKB"}
units = dict(zip(hdr, rows[1]))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "").split("(")[0].replace("void ", "")[:30]
    vals = []
    for k, a in want.items():
        v = d.get(k, "")
        try:
            f = float(v.replace(",", ""))
            if a in ("rdMB", "wrMB"):
                u = units.get(k, "")
                f = f * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
            if a == "Minst":
                f /= 1e6
            vals.append(f"{a}={f:.4g}")
        except ValueError:
            vals.append(f"{a}={v}")
    st = {k: float(v.replace(",", "")) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
    top = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")[:12], round(v, 1))
           for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]]
    print(d.get("ID"), name, " ".join(vals), top)
