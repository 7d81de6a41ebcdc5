#!/bin/bash
# ncu full capture of one kernel (run on the GPU box): scripts_ncu_probe.sh <kernel regex> <out name> [bench args]
mkdir -p gpurun_out
K=${1:-k_probe}; O=${2:-probe}; shift; shift
NCU_CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-other-tier $*"
$NCU_CMD > gpurun_out/plain_$O.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 3 -c 1 -o gpurun_out/$O $NCU_CMD > gpurun_out/ncu_$O.log 2>&1
tail -2 gpurun_out/ncu_$O.log
