# Per-launch DRAM traffic of the current build for the bench line, then the bench itself -- usage under gpurun:
#   bash tools/traffic_round.sh <tag>
set -u
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_${tag}.log 2>&1 || { tail -5 gpurun_out/plain_${tag}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'gemm2_kernel|attention|beam_step' -s 380 -c 6 -o gpurun_out/prof_${tag} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
python tools/ncu_traffic.py gpurun_out/prof_${tag}.ncu-rep ${tag} 2>&1 | tail -5
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_${tag}.json
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"], d["clocks"])
PY
