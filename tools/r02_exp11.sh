#!/bin/bash
# round-2 experiment 11: two-diagonal-words-per-thread band kernel: parity (whole band suite under K4B_DIAG_DW=2) and speed
set -u
mkdir -p gpurun_out
K4B_DIAG_DW=2 python -m pytest tests/test_gpu_diag.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest DW=2 rc=$?"; tail -3 gpurun_out/pytest_gpu11.log
for w in cfg2 cfg5k16 cfg5k128 cfg1; do for dw in 1 2; do
  K4B_DIAG_DW=$dw python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --configs none > gpurun_out/bench11_${w}_dw$dw.json 2> gpurun_out/bench11_${w}_dw$dw.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench11_${w}_dw$dw.json").read().strip().splitlines()[-1])
    print("$w dw=$dw", round(d["value"]), d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"], d["parity"]["ok"])
except Exception as e: print("$w dw=$dw parse failed", e)
PY
done; done
SHORT2="python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --configs none"
K4B_DIAG_DW=2 ncu --set full --clock-control none --import-source on -k regex:diag_min -s 40 -c 4 \
    -o gpurun_out/r02_prof_diag_dw2 -f $SHORT2 > gpurun_out/ncu_full11.log 2>&1
echo "ncu full rc=$?"
