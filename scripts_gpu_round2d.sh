#!/bin/bash
# round-2 (final state) evidence run on the GPU box: tests, smoke, both bench arms, ncu launch list of the match / pack /
# ingest / decode passes, one full capture each of k_clean, k_scan_ranges and k_probe (each only after the same command
# has exited 0 without ncu); usage: scripts_gpu_round2b.sh [tag]
TAG=${1:-r02d}
mkdir -p gpurun_out
make -s -C oracle oracle
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err | cut -c1-300
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
NCU_CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-other-tier --no-collection --no-cli --long-seconds 0"
$NCU_CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $NCU_CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
python profiles/ncu_launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_summary.txt; head -40 gpurun_out/launches_${TAG}_summary.txt
for K in k_clean; do
  ncu --set full --clock-control none --import-source on -k regex:"$K" -s 1 -c 1 -o gpurun_out/ncu_${K}_$TAG $NCU_CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  ncu -i gpurun_out/ncu_${K}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${K}_${TAG}_raw.csv 2>/dev/null
  python profiles/ncu_raw_summary.py gpurun_out/ncu_${K}_${TAG}_raw.csv > gpurun_out/ncu_${K}_${TAG}_raw_summary.txt 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
