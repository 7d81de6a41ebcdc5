#!/bin/bash
# helper run on the GPU box: tests, smoke, bench (both arms), ncu launch list; usage: scripts_gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
make -s -C oracle oracle
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 1500 gpurun_out/bench_ref_$TAG.json
NCU_CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-other-tier"
$NCU_CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $NCU_CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
python profiles/ncu_launch_summary.py gpurun_out/launches_$TAG.csv | head -30
