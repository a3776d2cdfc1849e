#!/bin/bash
# round-2 experiment 3: the full default bench line (all configs) on one GPU, and the reference arm (short)
set -u
mkdir -p gpurun_out
( time python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench3_full.json 2> gpurun_out/bench3_full.err; echo "bench rc=$?"
tail -5 gpurun_out/bench3_full.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench3_full.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"])
print("cpu", d["cpu_baseline"])
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"))
PY
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench3_ref.json
