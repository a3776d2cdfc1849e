#!/bin/bash
# round-2 experiment 14: join kernel variants under ncu
set -u
mkdir -p gpurun_out
for mk in 0 1; do
K4B_SEED_MASKED=$mk python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench14_cfg4_mk$mk.json 2> gpurun_out/bench14_cfg4_mk$mk.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench14_cfg4_mk$mk.json").read().strip().splitlines()[-1])
print("cfg4 masked=$mk", d["value"], d["ms_per_step"], d["parity"]["ok"], d["roofline"]["frac"])
PY
K4B_SEED_MASKED=$mk ncu --set full --clock-control none --import-source on -k regex:seed_join -s 3 -c 1 -o gpurun_out/r02_prof_join_mk$mk -f \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_join_mk$mk.log 2>&1; echo "ncu rc=$?"
done
python tools/pack_probe.py > gpurun_out/pack_probe14.log 2>&1; cat gpurun_out/pack_probe14.log
