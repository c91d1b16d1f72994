#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu full capture of the dominant kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [bench args...]
set -u
tag=${1:-r01}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_${tag}.log
tail -3 gpurun_out/pytest_${tag}.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/plain_${tag}.log 2>&1 || { tail -5 gpurun_out/plain_${tag}.log; exit 1; }
# one decode step = 6 launches (TD-LSTM, dec_att, attention, LM-LSTM, logits, beam_step); skip warm-up decodes
ncu --set full --clock-control none --import-source on -k regex:'gemm2_kernel|attention|beam_step' -s 380 -c 6 -o gpurun_out/prof_${tag} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
# per-launch DRAM traffic of THIS build's kernels for the bench line (stamped with the kernel sources' hash)
python tools/ncu_traffic.py gpurun_out/prof_${tag}.ncu-rep ${tag} > /dev/null 2>&1 && cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_${tag}.json
python bench.py "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/ncu_list_${tag}.log 2>&1
echo "ncu list rc=$?"
ls -la gpurun_out | tail -8
# the one-time projection GEMM (EPI_STORE, 55296 x 1024 x 2048): 22 store-epilogue launches per decode (projection, hoisted mean
# term, 20 x dec_att); skip the warm-up decodes
# (ncu matches -k on the base name: 1 gemm2_kernel launch at load + 82 per decode -> the 4th decode's first three launches)
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 247 -c 3 -o gpurun_out/prof_proj_${tag} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras "$@" > gpurun_out/ncu_proj_${tag}.log 2>&1
echo "ncu projection rc=$?"
# agreement with the reference's own captions (5000 BUTD + 1000 AoA images, both math modes)
python tests/tools/agreement.py --set butd --images 5000 --out gpurun_out/agreement_butd_5000_${tag}.json 2>&1 | tail -2
python tests/tools/agreement.py --set aoa_bu --images 1000 --out gpurun_out/agreement_aoa_bu_1000_${tag}.json 2>&1 | tail -2
