#!/bin/bash
# helper run on the GPU box: tests, smoke, bench, ncu launch list + full capture of the probe kernel
mkdir -p gpurun_out
make -s -C oracle oracle
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json
NCU_CMD="python bench.py --steps 2 --warmup 1 --genomes 20 --no-cpu"
$NCU_CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu1.log 2>&1
$NCU_CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_probe -s 2 -c 2 -o gpurun_out/probe $NCU_CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
