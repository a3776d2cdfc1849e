#!/bin/bash
# round-2 experiment 1: GPU suite, band kernel variants (window table vs row table, 4096 vs 8192 rows), ncu
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0"
for v in "0 4096" "1 4096" "1 8192" "0 8192"; do
  set -- $v
  K4B_DIAG_EWIN=$1 K4B_DIAG_ROWS=$2 $B > gpurun_out/bench_ewin$1_rows$2.json 2> gpurun_out/bench_ewin$1_rows$2.err
  echo "ewin=$1 rows=$2 rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ewin$1_rows$2.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"])
except Exception as e: print("parse failed", e)
PY
done
K4B_DIAG_EWIN=1 K4B_DIAG_ROWS=8192 python bench.py --workload cfg5k32 --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_cfg5k32_ewin1.json 2>&1
K4B_DIAG_EWIN=0 python bench.py --workload cfg5k32 --steps 2 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_cfg5k32_ewin0.json 2>&1
tail -c 400 gpurun_out/bench_cfg5k32_ewin1.json | head -c 10 >/dev/null
SHORT2="python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu --e2e-steps 0"
ncu --set full --clock-control none --import-source on -k regex:diag_min -s 40 -c 4 \
    -o gpurun_out/r02_prof_diag_ewin -f $SHORT2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
python tools/pack_probe.py > gpurun_out/pack_probe.log 2>&1; echo "pack probe rc=$?"; cat gpurun_out/pack_probe.log
ncu --set full --clock-control none -k regex:"pack_kernel|valid_kernel|revcomp_planes_kernel" -c 12 \
    -o gpurun_out/r02_prof_pack -f python tools/pack_probe.py > gpurun_out/ncu_pack.log 2>&1
echo "ncu pack rc=$?"
ls -la gpurun_out | tail -20
