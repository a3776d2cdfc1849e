#!/bin/bash
# round-2 experiment 17: Crick launches on a side stream + half-length segments for small launches
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_diag.py -m gpu -q -x 2>&1 | tail -2
for w in cfg2 cfg1 cfg5k32; do for one in 1 0; do
  if [ $one = 1 ]; then export K4B_DIAG_ONE_STREAM=1; else unset K4B_DIAG_ONE_STREAM; fi
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --configs none > gpurun_out/bench17_${w}_one$one.json 2> gpurun_out/bench17_${w}_one$one.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench17_${w}_one$one.json").read().strip().splitlines()[-1])
    print("$w one_stream=$one", round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"], d["parity"]["ok"])
except Exception as e: print("$w parse failed", e)
PY
done; done
