#!/bin/bash
# Quick GPU check of a kernel change (run through gpurun): parity tests, then the short bench twice
# and every tuning variant under codeps_b200/variants/ (tools/build_variants.sh).
#   tools/gpu_check.sh [pytest targets...]
cd "$(dirname "$0")/.."
python -m pytest ${@:-tests/test_gpu_parity.py tests/test_heads.py} -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    print('main', 'value=%.0f'%d['value'], 'ms_per_step=%.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['kernel_ms_all'].items() if v})
"
done
ls codeps_b200/variants/lib_*.so >/dev/null 2>&1 && tools/bench_variants.sh
exit 0
