#!/bin/bash
# Runs on the GPU box under gpurun: default bench (band engine), then the ncu launch list and one
# full capture of the dominant kernel on shortened bench commands (same code path).
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 3500 gpurun_out/bench_n1.json
SHORT="python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu --e2e-steps 0"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
SHORT2="python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu --e2e-steps 0"
$SHORT2 > gpurun_out/plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:diag_min -s 40 -c 4 \
    -o gpurun_out/prof_diag -f $SHORT2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
