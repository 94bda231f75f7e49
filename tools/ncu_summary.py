#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu -i rep --page raw --csv) into one line per kernel launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%")]
units = rows[1]
for r in rows[2:]:
    name = r[col["Kernel Name"]][:40]
    out = [name]
    for key, label in want:
        if key in col:
            v = r[col[key]]; u = units[col[key]]
            try:
                f = float(v.replace(",", ""))
                if label == "us":
                    f = f / 1000 if u == "ns" else (f * 1000 if u == "ms" else f)
                if label in ("rdMB", "wrMB"):
                    f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
                out.append(f"{label}={f:.4g}")
            except ValueError:
                out.append(f"{label}={v}")
    print("  ".join(out))
