#!/bin/bash
# One evidence pass on the GPU box (run through gpurun):  tools/evidence.sh r02_9
#   full GPU test suite (parity records), ncu --set full of one step's kernels (+ tracked summary and
#   per-phase view of the tile kernel), the launch list of the bench command, the full bench line.
# Everything lands in gpurun_out/; the summaries also in profiles/ of the box copy (so that bench.py
# picks the new ncu record up) -- copy them to profiles/ here afterwards.
name=${1:-evidence}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/${name}_gputests.log 2>&1; tail -2 gpurun_out/${name}_gputests.log
tools/ncu_step.sh ${name}_step > gpurun_out/${name}_ncu.log 2>&1
python tools/ncu_to_profile.py gpurun_out/${name}_step.ncu-rep profiles/${name}_step_ncu --workload cityscapes_b8 > /dev/null
cp profiles/${name}_step_ncu.json profiles/${name}_step_ncu.txt gpurun_out/
ncu -i gpurun_out/${name}_step.ncu-rep --page source --csv --print-source sass --kernel-name regex:cdp_photo > /tmp/src.csv 2>/dev/null
python tools/ncu_phase_breakdown.py /tmp/src.csv > gpurun_out/${name}_photo_phases.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${name}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
python bench.py > gpurun_out/${name}_bench.json 2> gpurun_out/${name}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${name}_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'path_frac', d['roofline'].get('path_frac'), 'clocks', d['clocks'])
print({k: (v.get('value') if isinstance(v, dict) else v) for k, v in d['extras'].get('workloads', {}).items()})
"
