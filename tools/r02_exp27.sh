#!/bin/bash
# round-2 experiment 27: seed index by two partition passes (K4B_SEED_INDEX=1): parity, then timing at cfg4 size
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_seed.py -m gpu -q -x -k "partition" 2>&1 | tail -15
timeout 200 python tools/index_build_probe.py > gpurun_out/index_build_probe.jsonl 2> gpurun_out/index_build_probe.err; echo "probe rc=$?"
cat gpurun_out/index_build_probe.jsonl; tail -3 gpurun_out/index_build_probe.err
