#!/bin/bash
# round-2 final evidence run at HEAD (1 GPU): GPU suite + smoke + the full default bench line (all configs, both reference legs)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final2_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final2_smoke.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 ) > gpurun_out/final2_bench_n1.json 2> gpurun_out/final2_bench_n1.err; echo "bench rc=$?"
tail -4 gpurun_out/final2_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/final2_bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"], d["ms_per_step"])
print("cpu", d["cpu_baseline"])
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"), (c.get("cpu_baseline") or {}).get("value"), ((c.get("cpu_baseline") or {}).get("file_parity") or {}).get("identical"))
PY
