#!/bin/bash
# round-2 experiment 23 (recorded and dropped; the code is not in the tree): seed index built by sorting (key pass, radix sort, offsets, gather) vs the count / cursor-fill build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_seed.py -m gpu -q -x 2>&1 | tail -3
for idx in 1 0; do
K4B_SEED_INDEX=$idx timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/bench20_idx$idx.json 2> gpurun_out/bench20_idx$idx.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench20_idx$idx.json").read().strip().splitlines()[-1])
print("cfg4 sorted_index=$idx", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity"]["ok"], d["roofline"]["frac"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg4_sorted.csv \
   python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cfg4_sorted.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02_launches_cfg4_sorted.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    tot.setdefault(r[ki][:70], []).append(v)
for k, v in tot.items():
    print("%-70s n=%3d last=%10.1f us" % (k, len(v), v[-1] / 1e3))
PY
