#!/usr/bin/env python
"""Static SASS view of one kernel of a built library: instruction counts per block-barrier segment
(= per phase of the tile kernel) and an opcode histogram per segment.

    python tools/sass_segments.py codeps_b200/libcodeps_photo.so cdp_photo_kernelILb1ELb0E [--dump SEG]
"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    dump = int(sys.argv[sys.argv.index("--dump") + 1]) if "--dump" in sys.argv else None
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = next(f for f in funcs if f.split("\n", 1)[0].find(pat) >= 0)
    seg, segs = 0, collections.defaultdict(list)
    for line in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(2).strip()
        segs[seg].append((m.group(1), ins))
        if "BAR.SYNC" in ins:
            seg += 1
    for k, v in segs.items():
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for _, i in v)
        print(f"segment {k}: {len(v)} instructions; " + ", ".join(f"{o} {n}" for o, n in ops.most_common(14)))
    if dump is not None:
        for a, i in segs[dump]:
            print(a, i)


if __name__ == "__main__":
    main()
