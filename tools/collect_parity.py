#!/usr/bin/env python
"""Turn gpurun_out/parity_records.jsonl (written by the -m gpu tests, tests/helpers.py) into the
tracked summary profiles/rNN_parity.json.

    python tools/collect_parity.py gpurun_out/parity_records.jsonl profiles/r02_parity.json
"""
import collections
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cases = collections.OrderedDict()
    for line in open(src):
        line = line.strip()
        if not line:
            continue
        r = json.loads(line)
        key = f"{r.get('build', 'default')}::{r.get('case', '?')}"
        what = r.pop("what", "?")
        cases.setdefault(key, collections.OrderedDict())[what] = {k: v for k, v in r.items() if k not in ("case", "build")}
    out = {"tolerances": {"loss_rel": 1e-5, "grad_rel_of_max_abs": 1e-4, "argmin": "bit-exact where the fp64 top-2 gap > 1e-6"},
           "columns": "per gradient: elements, excluded (discrete switches, from the fp64 oracle), excluded_frac, "
                      "q999 (99.9 % quantile of |got - fp64| / max|fp64| over the kept elements), worst (kept), "
                      "worst_unmasked (all elements), fp32_reference_worst (the reference algorithm in fp32 vs its fp64 run)",
           "cases": cases}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(cases)} cases -> {dst}")


if __name__ == "__main__":
    main()
