"""Join kernel shapes and the phase schedule of the seed engine on the cfg4 workload (BASELINE configs[3]): one
data set, then K4B_SEED_JOIN_VARIANT x K4B_SEED_PHASES (read by the library at every launch); prints the kernel
time of the whole seed-engine call (index + item keys + sorts + join) and a checksum of the minima."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kit4b_b200 as k4b
from kit4b_b200 import hamm

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
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
for phases in ("0", "1"):
    for variant in ("0", "1", "2"):
        os.environ["K4B_SEED_PHASES"] = phases
        os.environ["K4B_SEED_JOIN_VARIANT"] = variant
        ms = []
        for rep in range(4):
            hamm.best_init_device(best.data_ptr(), L, K); torch.cuda.synchronize()
            hamm.targeted_seed_device(q, t, True, clamp, core, 0, L, best.data_ptr()); torch.cuda.synchronize()
            ms.append(round(hamm.last_kernel_ms(), 2))
        chk = int(torch.clamp(best, max=clamp).to(torch.int64).sum().item())
        sums.add(chk)
        print(json.dumps({"phases": int(phases), "variant": int(variant), "kernel_ms": ms, "checksum": chk}), flush=True)
print(json.dumps({"all_checksums_equal": len(sums) == 1}))
