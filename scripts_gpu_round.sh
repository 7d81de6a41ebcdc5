#!/bin/bash
# helper run on the GPU box: tests, smoke, bench (both arms); usage: scripts_gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
make -s -C oracle oracle
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --tier div --no-cpu > gpurun_out/bench_${TAG}_div.json 2> gpurun_out/bench_${TAG}_div.err; echo "bench div rc=$?"; tail -c 3500 gpurun_out/bench_${TAG}_div.json; tail -5 gpurun_out/bench_${TAG}_div.err
