"""BASELINE configs[3] at full size: K=32 K-mers of a 2 Mbp probe set (1 Mbp copied from the assembly
with 3 % substitutions + 1 Mbp fresh random) vs a 500 Mbp synthetic assembly (20 x 25 Mbp), both
strands, R=3, through the host-buffer C ABI (k4b_hamm_targeted -> band engine).  Prints timing,
the result distribution and a cross-check of sampled probes against the POPC engine."""
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
    parts.append(rng.integers(0, 4, size=clen, dtype=np.uint8))
    parts.append(np.array([7], dtype=np.uint8))          # .sfx layout: every entry is followed by EOS
target = np.ascontiguousarray(np.concatenate(parts))
rng = np.random.default_rng(42)
pl = int(1_000_000 * scale)
src = int(3.3 * clen) + 12345
copy = target[src:src + pl].copy()
assert (copy != 7).all()
idx = rng.choice(pl, size=int(0.03 * pl), replace=False)
copy[idx] = (copy[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
probes = np.ascontiguousarray(np.concatenate([copy, [7], rng.integers(0, 4, size=pl, dtype=np.uint8)]), dtype=np.uint8)
k4b.gpu_init(1)
t0 = time.perf_counter()
out = k4b.targeted(target, probes, K, R, True)
dt = time.perf_counter() - t0
nq = int((out != 0xFF).sum())
nt = nchr * (clen - K + 1)
hist = {int(v): int((out == v).sum()) for v in range(0, 6)}
print(json.dumps({"target_bases": int(len(target)), "probe_kmers": nq, "seconds_e2e": round(dt, 2), "band_ms": round(hamm.last_kernel_ms(), 1),
                  "Gcmp_s_e2e": round(nq * nt * 2 / dt / 1e9, 1), "hist": hist}), flush=True)
# cross-check 3 windows of probes against the POPC engine
hamm.set_engine(hamm.ENGINE_POPC)
bad = 0
for b in (0, pl - 600, pl + 5000):
    sub = np.ascontiguousarray(probes[b:b + 256 + K - 1])
    if (sub == 7).any():
        continue
    want = k4b.targeted(target, sub, K, R, True)
    bad += int((want[:256] != out[b:b + 256]).sum())
print(json.dumps({"crosscheck_mismatches_vs_popc_engine": bad}), flush=True)
