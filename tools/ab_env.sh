#!/bin/bash
# A/B an environment switch on the default bench: bash tools/ab_env.sh VAR v1 v2 v3 ...   (prints value, ms/step, attention ms)
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 30 --no-cpu-baseline 2>/dev/null > /tmp/ab.json
  python - "$var=$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
k = d["kernels"]
print(sys.argv[1], round(d["value"]), "captions/s", round(d["ms_per_step"], 3), "ms; attention", round(k["attention"]["ms_per_step"], 3),
      "lstm", round(k["gemm_lstm"]["ms_per_step"], 3), "logits", round(k["gemm_logits"]["ms_per_step"], 3), "sm_mhz", d["clocks"]["sm_mhz"])
PY
done
