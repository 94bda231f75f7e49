#!/bin/bash
cd /root/repo
for s in 3 12; do python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --shift-px $s 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    print('shift $s', 'value=%.0f'%d['value'], 'photo=%.4f'%d['roofline']['kernel_ms_all']['photo'])
"; done
