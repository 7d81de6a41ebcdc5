#!/bin/bash
# ncu capture of the probe kernel (run on the GPU box): launch list + full set with source
mkdir -p gpurun_out
NCU_CMD="python bench.py --steps 2 --warmup 1 --genomes 20 --no-cpu"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu1.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${1:-k_probe}" -s 3 -c 2 -o gpurun_out/${2:-probe} $NCU_CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
