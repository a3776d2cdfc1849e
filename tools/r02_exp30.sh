#!/bin/bash
# round-2 experiment 30: persistent pass A + reserve atomics in flight during the grouping: parity, timing at cfg4 size
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_seed.py -m gpu -q -x -k "partition" 2>&1 | tail -5
timeout 200 python tools/index_build_probe.py --builds 0,29,61,1,29,61 > gpurun_out/index_build_probe4.jsonl 2> gpurun_out/index_build_probe4.err; echo "probe rc=$?"
cat gpurun_out/index_build_probe4.jsonl; tail -3 gpurun_out/index_build_probe4.err
