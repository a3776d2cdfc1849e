#!/bin/bash
# round-2 final evidence run (1 GPU): GPU suite + smoke, the full default bench line, launch lists, ncu of the pack phase
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
tail -4 gpurun_out/final_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/final_bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"], d["ms_per_step"])
print("cpu", d["cpu_baseline"])
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"), (c.get("cpu_baseline") or {}).get("value"), ((c.get("cpu_baseline") or {}).get("file_parity") or {}).get("identical"))
PY
SHORT="python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 --configs none"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg2.csv $SHORT > gpurun_out/final_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/pack_probe.py > gpurun_out/final_pack_probe.log 2>&1; cat gpurun_out/final_pack_probe.log
timeout 600 ncu --set full --clock-control none -k regex:"pack_kernel|valid_kernel|revcomp_planes_kernel" -c 12 \
    -o gpurun_out/final_prof_pack -f python tools/pack_probe.py > gpurun_out/final_ncu_pack.log 2>&1; echo "ncu pack rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"seed_scan" -s 6 -c 2 \
    -o gpurun_out/final_prof_seed_scan -f python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/final_ncu_seed.log 2>&1; echo "ncu seed rc=$?"
