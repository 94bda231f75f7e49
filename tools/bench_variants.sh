#!/bin/bash
# run bench.py for each variant library: prints photo kernel ms and value
cd "$(dirname "$0")/.."
for f in codeps_b200/variants/lib_*.so; do
  CODEPS_B200_LIB=$PWD/$f python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    print('$f', 'value=%.0f'%d['value'], 'ms_per_step=%.3f'%d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['kernel_ms_all'].items() if v})
"
done
