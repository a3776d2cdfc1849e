#!/bin/bash
# Runs on the GPU box under gpurun: default bench, then the ncu launch list and one full capture
# of the dominant kernel on a shortened bench command (same code path, smaller query batch).
set -u
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_n1.json
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 --batch 16384"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/plain_short2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:allpairs_min -s 3 -c 1 \
    -o gpurun_out/prof_allpairs -f $SHORT > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
