#!/bin/bash
# round-2 experiment 13: join kernel with pre-masked tile copies + CTA-aggregated index passes: parity, cfg4 timing, launch lists
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_seed.py -m gpu -q -x > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu13.log
K4B_SEED_AGG=1 python -m pytest tests/test_gpu_seed.py -m gpu -q -x > gpurun_out/pytest_gpu13b.log 2>&1; echo "pytest AGG=1 rc=$?"; tail -3 gpurun_out/pytest_gpu13b.log
for agg in 0 1; do
K4B_SEED_AGG=$agg python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench13_cfg4_agg$agg.json 2> gpurun_out/bench13_cfg4_agg$agg.err; echo "cfg4 agg=$agg rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench13_cfg4_agg$agg.json").read().strip().splitlines()[-1])
print("cfg4 agg=$agg", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_cfg4_b.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4b.log 2>&1; echo "ncu cfg4 rc=$?"
grep -E "seed_join|seed_scan|seed_agg" gpurun_out/r02_launches_cfg4_b.csv | tail -4 | cut -c1-200
