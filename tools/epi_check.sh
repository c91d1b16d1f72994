#!/bin/bash
# parity tests, then the per-kernel tables of the workloads an epilogue change touches -- usage under gpurun: bash tools/epi_check.sh
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms_per_step'],3), v.get('tflops') and round(v['tflops'])) for k,v in d['kernels'].items()}, d['clocks']['sm_mhz'], d.get('parity_sample'))"; }
python bench.py --workload scst --batch 512 --steps 10 --no-cpu-baseline --no-extras 2>/dev/null | show scst
python bench.py --workload nic --steps 10 --no-cpu-baseline --no-extras 2>/dev/null | show nic
python bench.py --steps 20 --no-cpu-baseline --no-extras 2>/dev/null | show headline
python bench.py --batch 16 --steps 200 --no-cpu-baseline --no-extras 2>/dev/null | show b16
