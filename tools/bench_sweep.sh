#!/bin/bash
# All selectable workloads at one GPU: bash tools/bench_sweep.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
for w in nic nic_images butd_spatial aoa aoa_bu scst; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --cpu-images 2 > gpurun_out/bench_${tag}_${w}.json 2> gpurun_out/bench_${tag}_${w}.err
  echo "$w rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${tag}_${w}.json"))
    print("  ", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_sample"))
except Exception as e:
    print("   parse failed:", e)
PY
  tail -2 gpurun_out/bench_${tag}_${w}.err
done
