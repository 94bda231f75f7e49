#!/usr/bin/env python
"""Turn an ncu report into the tracked summaries under profiles/:

    python tools/ncu_to_profile.py gpurun_out/prof.ncu-rep profiles/r01_step_ncu [--workload cityscapes_b8]

writes <out>.json (one record per kernel launch: duration, DRAM bytes, throughput percentages,
issue-slot utilisation, registers, L1/L2 hit rates) and <out>.txt (the same as a table).
bench.py reads the JSON to fill roofline.traffic for the dominant kernel."""
import csv
import json
import subprocess
import sys
import time

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
}
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    records = []
    for r in rows[2:]:
        rec = {"kernel": r[col["Kernel Name"]]}
        for key, name in KEYS.items():
            if key in col:
                try:
                    rec[name] = float(r[col[key]].replace(",", "")) * SCALE.get(units[col[key]], 1.0)
                except ValueError:
                    pass
        if "dram_read_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
        records.append(rec)
    blob = {"source": rep, "workload": workload, "how": "ncu --set full --clock-control none --import-source on",
            "created": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),  # bench.py reads the newest record
            "launches": records}
    with open(out + ".json", "w") as f:
        json.dump(blob, f, indent=1)
    with open(out + ".txt", "w") as f:
        f.write(f"# {rep}  workload={workload}\n")
        for rec in records:
            f.write(f"{rec['kernel'][:48]:48s} {rec.get('duration_us', 0):9.1f} us  dram {rec.get('dram_bytes', 0) / 1e6:8.1f} MB "
                    f"({rec.get('dram_pct_of_peak', 0):5.1f}% of peak)  sm {rec.get('sm_pct_of_peak', 0):5.1f}%  "
                    f"issue {rec.get('issue_slot_pct', 0):5.1f}%  occ {rec.get('achieved_occupancy_pct', 0):5.1f}%  "
                    f"regs {int(rec.get('registers_per_thread', 0)):3d}  L2hit {rec.get('l2_hit_pct', 0):5.1f}%  "
                    f"L1hit {rec.get('l1_hit_pct', 0):5.1f}%  warp-inst {rec.get('warp_instructions', 0):.3e}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
