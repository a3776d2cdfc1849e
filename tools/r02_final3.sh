#!/bin/bash
# round-2 final check at HEAD (1 GPU) after the partition index build: GPU suite, smoke, cfg4 bench line
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/final3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final3_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/final3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final3_smoke.log
timeout 80 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/final3_bench_cfg4.json 2> gpurun_out/final3_bench_cfg4.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/final3_bench_cfg4.json
