#!/bin/bash
# round-2 experiment 26: bucket-per-CTA join + phase schedule as the default: parity, cfg4 bench, launch list, ncu --set full of the join
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seed.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench26.json 2> gpurun_out/bench26.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench26.json").read().strip().splitlines()[-1])
print("cfg4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4_bucket_join.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4_bj.log 2>&1
grep -E "seed_join|seed_scan|seed_item" gpurun_out/r02_launches_cfg4_bucket_join.csv | tail -12 | awk -F'","' '{print $5, $(NF)}' | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seed_join" -s 12 -c 4 -o gpurun_out/r02_prof_join_bucket -f \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_join_bucket.log 2>&1; echo "ncu full rc=$?"
