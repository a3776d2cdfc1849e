#!/bin/bash
# round-2 experiment 9 (2 GPUs): multi-GPU tests, torchrun bench with distributed legs + in-process leg, reference arm under torchrun
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "multi or distributed or inproc or in_process" > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu9.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --configs cfg1,cfg5k32,cfg4,cfg3 ) > gpurun_out/bench9_n2.json 2> gpurun_out/bench9_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/bench9_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench9_n2.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"])
print("inproc", d.get("e2e_inproc"))
for c in d["configs"]:
    print(c.get("workload","?")[:60], c.get("value"), (c.get("e2e") or {}).get("value"), (c.get("roofline") or {}).get("frac"), (c.get("parity") or {}).get("ok"), c.get("leg_wall_seconds"), c.get("error"))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench9_ref.json
