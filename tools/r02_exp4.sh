#!/bin/bash
# round-2 experiment 4: GPU suite + smoke after the host-path / seed-index rewrite; cfg4 bench + trace; full bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu4.log
python __graft_entry__.py smoke > gpurun_out/smoke4.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke4.log
K4B_TRACE=1 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench4_cfg4.json 2> gpurun_out/bench4_cfg4.err; echo "cfg4 rc=$?"
tail -30 gpurun_out/bench4_cfg4.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench4_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", d["value"], d["ms_per_step"], d["e2e"], d["parity"], d["roofline"]["kernel_ms"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_cfg4.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench4_full.json 2> gpurun_out/bench4_full.err; echo "bench rc=$?"
tail -5 gpurun_out/bench4_full.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench4_full.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"])
print("cpu", d["cpu_baseline"])
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"))
PY
