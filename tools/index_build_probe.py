"""Seed index build on the cfg4 workload (2 Mbp probes vs 500 Mbp, K=32, R=3): the one-by-one build
(K4B_SEED_INDEX=0) against the two partition passes (flags of k4b_kernels.cuh: 1 plain, 2 tiles of 4096 entries,
4 shared-memory count pass, 8 three-word extraction, 16 persistent second pass).  Per build: the engine's own CUDA-event time of a
call that joins ONE probe K-mer (= the index build) and of the whole job, and the checksum of the minima."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kit4b_b200 as k4b
from kit4b_b200 import hamm

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--builds", default="0,1,5,9,17,29,14,1,29")   # K4B_SEED_INDEX values, in this order
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--index-only", action="store_true")  # one index-only call per build (for ncu)
args = ap.parse_args()
scale = args.scale
K, R = 32, 3
rng = np.random.default_rng(41)
nchr, clen = 20, int(25_000_000 * scale)
parts = []
for c in range(nchr):
    parts += [rng.integers(0, 4, size=clen, dtype=np.uint8), np.array([7], dtype=np.uint8)]
target = np.ascontiguousarray(np.concatenate(parts))
rng = np.random.default_rng(42)
pl = int(1_000_000 * scale)
src = int(3.3 * clen) + 12345
copy = target[src:src + pl].copy()
idx = rng.choice(pl, size=int(0.03 * pl), replace=False)
copy[idx] = (copy[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
probes = np.ascontiguousarray(np.concatenate([copy, [7], rng.integers(0, 4, size=pl, dtype=np.uint8)]), dtype=np.uint8)
k4b.gpu_init(1)
t = hamm.Packed.from_host(target, K)
q = hamm.Packed.from_host(probes, K)
L = len(probes)
core = K // (R + 1); clamp = K // core
best = torch.empty(L, dtype=torch.int32, device="cuda")
sums = set()
for build in args.builds.split(","):
    os.environ["K4B_SEED_INDEX"] = build
    idx_ms, job_ms = [], []
    if args.index_only:
        hamm.best_init_device(best.data_ptr(), L, K)
        hamm.targeted_seed_device(q, t, True, clamp, core, 0, 1, best.data_ptr()); torch.cuda.synchronize()
        continue
    for rep in range(args.reps):
        hamm.best_init_device(best.data_ptr(), L, K)
        hamm.targeted_seed_device(q, t, True, clamp, core, 0, 1, best.data_ptr()); torch.cuda.synchronize()
        idx_ms.append(round(hamm.last_kernel_ms(), 2))
        hamm.best_init_device(best.data_ptr(), L, K)
        n = hamm.targeted_seed_device(q, t, True, clamp, core, 0, L, best.data_ptr()); torch.cuda.synchronize()
        job_ms.append(round(hamm.last_kernel_ms(), 2))
    cs = int(torch.clamp(best, max=clamp).sum().item())
    sums.add(cs)
    print(json.dumps({"build": int(build), "index_ms": idx_ms, "job_ms": job_ms, "launches": n, "checksum": cs,
                      "indexed_cores": hamm.last_seed_info()["indexed_cores"]}), flush=True)
print(json.dumps({"all_checksums_equal": len(sums) == 1}))
