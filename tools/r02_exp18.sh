#!/bin/bash
set -u
mkdir -p gpurun_out
K4B_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --configs cfg5k32,cfg4 > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "rc=$?"
grep -B2 -A12 "H2D + pack" gpurun_out/bench18.err | tail -60
python - <<PY
import json
d=json.loads(open("gpurun_out/bench18.json").read().strip().splitlines()[-1])
for c in d["configs"]:
    print(c.get("workload","?")[:40], c.get("value"), c.get("e2e"))
PY
