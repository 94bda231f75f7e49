#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source file / function.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > src.csv
       python tools/ncu_line_breakdown.py src.csv [--top 25]

Prints, per source file, the share of executed warp instructions and of stall samples, then the
share per enclosing function (found by scanning the file for `CDP_HD ... name(`), then the hottest
lines.  Only the current checkout's sources are scanned, so run it on a profile of the same code.
"""
import argparse
import collections
import csv
import re
import sys

csv.field_size_limit(1 << 30)


def functions_of(path):
    """[(first_line, name)] for every function definition found in a source file."""
    out = []
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return out
    pat = re.compile(r"^(?:template.*>\s*)?(?:CDP_HD|static|__global__|__device__|extern \"C\")[^;=]*?\b(\w+)\(")
    for i, line in enumerate(lines, 1):
        m = pat.match(line)
        if m and not line.rstrip().endswith(";"):
            out.append((i, m.group(1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    per_line = collections.defaultdict(lambda: [0, 0, ""])  # (file, line) -> [inst, samples, text]
    cur = None
    cols = None
    for row in csv.reader(open(a.csv)):
        if not row:
            continue
        if row[0] == "File Path":
            cur = row[1]
        elif row[0] == "Line No":
            cols = row
        elif cur and cols and row[0].isdigit():
            inst = row[cols.index("Instructions Executed")]
            smp = row[cols.index("# Samples")]
            rec = per_line[(cur, int(row[0]))]
            rec[0] += int(inst) if inst.isdigit() else 0
            rec[1] += int(smp) if smp.isdigit() else 0
            rec[2] = row[1].strip()
    tot_i = sum(v[0] for v in per_line.values()) or 1
    tot_s = sum(v[1] for v in per_line.values()) or 1
    files = collections.defaultdict(lambda: [0, 0])
    funcs = collections.defaultdict(lambda: [0, 0])
    fn_cache = {}
    for (f, ln), (i, s, _) in per_line.items():
        files[f][0] += i
        files[f][1] += s
        if f not in fn_cache:
            fn_cache[f] = functions_of(f)
        name = "?"
        for first, nm in fn_cache[f]:
            if first <= ln:
                name = nm
        funcs[(f.split("/")[-1], name)][0] += i
        funcs[(f.split("/")[-1], name)][1] += s
    print(f"total warp instructions {tot_i:.4g}, stall samples {tot_s}")
    print("\n-- per file: inst% samples%")
    for f, (i, s) in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"{100*i/tot_i:6.2f} {100*s/tot_s:6.2f}  {f}")
    print("\n-- per function: inst% samples%")
    for k, (i, s) in sorted(funcs.items(), key=lambda kv: -kv[1][0]):
        if i / tot_i > 0.002 or s / tot_s > 0.002:
            print(f"{100*i/tot_i:6.2f} {100*s/tot_s:6.2f}  {k[0]}:{k[1]}")
    print(f"\n-- top {a.top} lines by samples: inst% samples% file:line text")
    for (f, ln), (i, s, t) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{100*i/tot_i:6.2f} {100*s/tot_s:6.2f}  {f.split('/')[-1]}:{ln}  {t[:110]}")


if __name__ == "__main__":
    sys.exit(main())
