#!/bin/bash
# round-2 experiment 12: auto choice of the two-word kernel, 8193-row segments
set -u
mkdir -p gpurun_out
for w in cfg5k64 cfg5k96; do for dw in 1 2; do
  K4B_DIAG_DW=$dw python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --configs none > gpurun_out/bench12_${w}_dw$dw.json 2> gpurun_out/bench12_${w}_dw$dw.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench12_${w}_dw$dw.json").read().strip().splitlines()[-1])
    print("$w dw=$dw", round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"], d["parity"]["ok"])
except Exception as e: print("$w dw=$dw parse failed", e)
PY
done; done
for w in cfg2 cfg1 cfg5k16 cfg5k128; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --configs none > gpurun_out/bench12_${w}_auto.json 2> gpurun_out/bench12_${w}_auto.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench12_${w}_auto.json").read().strip().splitlines()[-1])
    print("$w auto", round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"], d["parity"]["ok"])
except Exception as e: print("$w auto parse failed", e)
PY
done
python -m pytest tests/test_gpu_diag.py -m gpu -q -x 2>&1 | tail -3
