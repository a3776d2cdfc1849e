#!/bin/bash
# round-2 experiment 15: warp-specialised join kernel: parity + cfg4 timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seed.py -m gpu -q -x > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu15.log
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench15_cfg4.json 2> gpurun_out/bench15_cfg4.err; echo "cfg4 rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench15_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seed_join -s 3 -c 1 -o gpurun_out/r02_prof_join_ws -f \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_join_ws.log 2>&1; echo "ncu rc=$?"
