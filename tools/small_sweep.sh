#!/bin/bash
# Small-batch sweep of the default workload (captions/s and ms per batch) -- usage under gpurun: bash tools/small_sweep.sh <tag> [batches...]
tag=${1:-s}; shift || true
mkdir -p gpurun_out
for b in ${@:-1 4 16 32 42}; do
  python bench.py --batch $b --steps 200 --no-extras --cpu-images 4 > gpurun_out/small_${tag}_b$b.json 2> gpurun_out/small_${tag}_b$b.err || tail -3 gpurun_out/small_${tag}_b$b.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/small_${tag}_b$b.json"))
    print("B=$b", round(d["value"]), "cap/s", round(d["ms_per_step"],4), "ms/batch  e2e", round(d["e2e"]["value"]), d.get("parity_sample"), {k: round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("B=$b failed", e)
PY
done
