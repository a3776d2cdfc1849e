"""Phase timing of the seed-and-verify engine on the cfg4 workload (device-level API): pack,
first (cold allocator) and second (warm) engine call, finalize + D2H."""
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
def sync(): torch.cuda.synchronize()
res = {}
t0 = time.perf_counter(); t = hamm.Packed.from_host(target, K); sync(); res["pack_target_s"] = round(time.perf_counter() - t0, 3)
t0 = time.perf_counter(); q = hamm.Packed.from_host(probes, K); sync(); res["pack_probes_s"] = round(time.perf_counter() - t0, 3)
L = len(probes)
core = K // (R + 1); clamp = K // core
best = torch.empty(L, dtype=torch.int32, device="cuda")
out = torch.empty(L, dtype=torch.int16, device="cuda")
for rep in ("cold", "warm", "warm2"):
    hamm.best_init_device(best.data_ptr(), L, K); sync()
    t0 = time.perf_counter()
    hamm.targeted_seed_device(q, t, True, clamp, core, 0, L, best.data_ptr()); sync()
    res["seed_%s_s" % rep] = round(time.perf_counter() - t0, 3)
    res["seed_%s_kernel_ms" % rep] = round(hamm.last_kernel_ms(), 1)
t0 = time.perf_counter()
hamm.targeted_finalize_device(q, best.data_ptr(), clamp, out.data_ptr()); h = out.cpu().numpy(); res["finalize_d2h_s"] = round(time.perf_counter() - t0, 3)
res["hist"] = {int(v): int((h == v).sum()) for v in range(0, 6)}
print(json.dumps(res))
