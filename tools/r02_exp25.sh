#!/bin/bash
# round-2 experiment 25: join kernel shapes x phase schedule on cfg4
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seed.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/join_variants_probe.py 2>&1 | tee gpurun_out/join_variants.jsonl | tail -20
