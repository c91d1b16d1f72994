#!/bin/bash
# ncu full capture of the sampling-epilogue logit GEMM (gemm2_kernel<EPI_SAMPLE>) inside the SCST workload -- usage under gpurun:
#   bash tools/ncu_sample_epi.sh <tag>
tag=${1:-s}
mkdir -p gpurun_out
args="--workload scst --batch 512 --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
python bench.py $args > gpurun_out/plain_sample_${tag}.log 2>&1 || { tail -5 gpurun_out/plain_sample_${tag}.log; exit 1; }
tail -c 1500 gpurun_out/plain_sample_${tag}.log
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm2_kernelILi4 -s 50 -c 2 \
    -o gpurun_out/prof_sample_${tag} -f python bench.py $args > gpurun_out/ncu_sample_${tag}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_sample_${tag}.log
