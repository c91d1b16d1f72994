#!/bin/bash
# ncu full capture of the top-k logit GEMM (gemm2_kernel<EPI_TOPK, 4>) in a workload -- usage under gpurun:
#   bash tools/ncu_logits_epi.sh <tag> <workload> [skip]
tag=${1:-t}; wl=${2:-nic}; skip=${3:-50}
mkdir -p gpurun_out
args="--workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
python bench.py $args > gpurun_out/plain_logits_${tag}.log 2>&1 || { tail -5 gpurun_out/plain_logits_${tag}.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm2_kernelILi3 -s $skip -c 2 \
    -o gpurun_out/prof_logits_${tag} -f python bench.py $args > gpurun_out/ncu_logits_${tag}.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_logits_${tag}.log | cut -c1-200
