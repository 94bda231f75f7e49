#!/usr/bin/env python
"""Dynamic per-phase view of the tile kernel from an `ncu --page source --csv --print-source cuda,sass`
dump: SASS in address order is cut at every block barrier / mbarrier wait, and per segment the share
of executed warp instructions, of stall samples, the FMA-pipe cycles per issued instruction (packed
FFMA2 / FADD2 / FMUL2 count two) and the opcode mix are printed.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:cdp_photo > src.csv
       python tools/ncu_phase_breakdown.py src.csv
"""
import collections
import csv
import re
import sys

csv.field_size_limit(1 << 30)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ia = hdr.index("Address"); isrc = ia + 1
    iex = hdr.index("Instructions Executed"); ist = hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [(c[len("stall_"):], i) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    ins = {}
    for r in rows[h + 1:]:
        if len(r) <= iex or not re.fullmatch(r"(0x)?[0-9a-fA-F]+", r[ia] or ""):
            continue
        if r[ia] in ins:
            continue
        try:
            ins[r[ia]] = (r[isrc].strip(), int(r[iex]), int(r[ist]), [int(r[i] or 0) for _, i in stall_cols])
        except ValueError:
            continue
    seg, segs = 0, collections.defaultdict(lambda: [0, 0, collections.Counter(), 0, [0] * len(stall_cols)])
    for a in sorted(ins, key=lambda a: int(a, 16)):
        t, n, s, sr = ins[a]
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", t)
        op = m.group(2) if m else "?"
        g = segs[seg]
        g[0] += n; g[1] += s; g[2][op] += n
        g[4] = [x + y for x, y in zip(g[4], sr)]
        if op in ("FFMA2", "FADD2", "FMUL2"):
            g[3] += 2 * n
        elif op in ("FFMA", "FADD", "FMUL", "IMAD", "HFMA2"):
            g[3] += n
        if "BAR.SYNC" in t or "SYNCS.PHASECHK" in t:
            seg += 1
    ti = sum(g[0] for g in segs.values()); ts = sum(g[1] for g in segs.values())
    print(f"total warp instructions {ti:.4g}, stall samples {ts}")
    for k, g in segs.items():
        if not g[0]:
            continue
        print(f"seg {k:2d}: inst {g[0] / ti * 100:5.1f}%  samples {g[1] / ts * 100:5.1f}%  fma-cycles/inst {g[3] / g[0]:.2f}  "
              + ", ".join(f"{o} {n / g[0] * 100:.0f}" for o, n in g[2].most_common(9)))
        tot = sum(g[4])
        if tot:
            top = sorted(zip(g[4], (n for n, _ in stall_cols)), reverse=True)[:7]
            print("         stalls: " + ", ".join(f"{n} {v / tot * 100:.0f}%" for v, n in top))


if __name__ == "__main__":
    main()
