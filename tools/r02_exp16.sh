#!/bin/bash
# round-2 experiment 16 (8 GPUs): the default bench line under torchrun, all configs + the in-process leg
set -u
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench16_n8.json 2> gpurun_out/bench16_n8.err; echo "bench rc=$?"
tail -6 gpurun_out/bench16_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench16_n8.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"], d["ms_per_step"])
print("inproc", d.get("e2e_inproc"))
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"))
PY
