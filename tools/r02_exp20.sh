#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/pack_probe.py 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none -k regex:"pack_kernel|valid_kernel|revcomp_planes_kernel" -c 12 \
    -o gpurun_out/final_prof_pack -f python tools/pack_probe.py > gpurun_out/final_ncu_pack.log 2>&1; echo "ncu pack rc=$?"
