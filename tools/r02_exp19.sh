#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seed.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for s8 in 0 1; do
K4B_SEED_SCAN8=$s8 timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench19_s8$s8.json 2> gpurun_out/bench19_s8$s8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench19_s8$s8.json").read().strip().splitlines()[-1])
print("cfg4 scan8=$s8", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_cfg4_c.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4c.log 2>&1
grep -E "seed_join|seed_scan" gpurun_out/r02_launches_cfg4_c.csv | tail -3 | awk -F'","' '{print $5, $(NF)}' | cut -c1-200
timeout 300 python tools/pack_probe.py 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:"valid_kernel" -s 6 -c 2 -o gpurun_out/r02_prof_valid -f python tools/pack_probe.py > gpurun_out/ncu_valid.log 2>&1; echo "ncu valid rc=$?"
