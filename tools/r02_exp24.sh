#!/bin/bash
# round-2 experiment 24: seed join with a two-level signature bound (plane 0 first), 2048-entry tiles, explicit shared addressing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seed.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench24.json 2> gpurun_out/bench24.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench24.json").read().strip().splitlines()[-1])
print("cfg4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"], d["roofline"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4_join2.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4_join2.log 2>&1
grep -E "seed_join|seed_scan" gpurun_out/r02_launches_cfg4_join2.csv | tail -3 | awk -F'","' '{print $5, $(NF)}' | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seed_join" -s 3 -c 1 -o gpurun_out/r02_prof_join_2level -f \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_join_2level.log 2>&1; echo "ncu full rc=$?"
