#!/bin/bash
# round-2 experiment 29: partition index variants (shared-memory count pass, three-word extraction, persistent pass B):
# parity, timing at cfg4 size, ncu --set full of the combination; join kernel shapes with 5 CTAs per SM
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_seed.py -m gpu -q -x -k "partition" 2>&1 | tail -5
timeout 200 python tools/index_build_probe.py > gpurun_out/index_build_probe3.jsonl 2> gpurun_out/index_build_probe3.err; echo "probe rc=$?"
cat gpurun_out/index_build_probe3.jsonl; tail -3 gpurun_out/index_build_probe3.err
timeout 200 python tools/join_variants_probe.py > gpurun_out/join_variants2.jsonl 2> gpurun_out/join_variants2.err; echo "join probe rc=$?"
cat gpurun_out/join_variants2.jsonl; tail -3 gpurun_out/join_variants2.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"seed_part|seed_count" -c 3 -o gpurun_out/r02_prof_index_partition_v29 -f \
   python tools/index_build_probe.py --builds 29 --index-only > gpurun_out/ncu_index_partition_v29.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_index_partition_v29.log
