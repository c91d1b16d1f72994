#!/bin/bash
# parity tests + the SCST workload's per-kernel table -- usage under gpurun: bash tools/scst_check.sh [pytest -k expression]
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -3
python bench.py --workload scst --batch 512 --steps 10 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms_per_step'],3), v.get('tflops') and round(v['tflops'])) for k,v in d['kernels'].items()}, d['clocks'])"
