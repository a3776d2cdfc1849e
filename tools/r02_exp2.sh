#!/bin/bash
# round-2 experiment 2: GPU suite + smoke with the code-array band kernel, bench cfg2 short, host-path trace of cfg4
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
python __graft_entry__.py smoke > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke2.log
python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench2_cfg2.json 2> gpurun_out/bench2_cfg2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench2_cfg2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["result_checksum"])
PY
K4B_TRACE=1 python tools/cfg4_probe.py 1.0 > gpurun_out/cfg4_trace.log 2>&1; echo "cfg4 probe rc=$?"; tail -40 gpurun_out/cfg4_trace.log
