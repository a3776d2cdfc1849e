#!/bin/bash
# round-2 experiment 28: partition index in tiles of 2048 / 4096 entries: parity, timing at cfg4 size, ncu --set full
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_seed.py -m gpu -q -x -k "partition" 2>&1 | tail -5
timeout 200 python tools/index_build_probe.py > gpurun_out/index_build_probe2.jsonl 2> gpurun_out/index_build_probe2.err; echo "probe rc=$?"
cat gpurun_out/index_build_probe2.jsonl; tail -3 gpurun_out/index_build_probe2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"seed_part|seed_scan" -c 6 -o gpurun_out/r02_prof_index_partition -f \
   python tools/index_build_probe.py --builds 1,2 --index-only > gpurun_out/ncu_index_partition.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_index_partition.log
