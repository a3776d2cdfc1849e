"""BASELINE configs[4] on the band engine: FULL all-vs-all passes (both strands) over the 5 Mbp
synthetic genome at K = 16/32/64/96/128 (+25, 50, 100, 200, 500): seconds per job, Gcmp/s, counter
planes used, and a cross-check of 3 x 4096 queries against the POPC engine.  One JSON object per K."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kit4b_b200 as k4b
from kit4b_b200 import hamm

k4b.gpu_init(1)
rng = np.random.default_rng(51)
G = 5_000_000
concat = np.ascontiguousarray(rng.integers(0, 4, size=G, dtype=np.uint8))
d_concat = torch.from_numpy(concat).cuda()
best = torch.empty(G, dtype=torch.int32, device="cuda")
out = torch.empty(G, dtype=torch.int16, device="cuda")
B = 4096
chk = torch.empty(B, dtype=torch.int16, device="cuda")
for K in [int(k) for k in (sys.argv[1:] or [16, 25, 32, 50, 64, 96, 100, 128, 200, 500])]:
    g = hamm.Packed.from_device(d_concat.data_ptr(), G, K)
    secs = []
    for rep in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hamm.best_init_device(best.data_ptr(), G, K)
        ev[0].record()
        hamm.diag_bootstrap_device(g, True, 0, G, best.data_ptr())
        ev[1].record()
        n = hamm.diag_bands_device(g, True, 0, 1, best.data_ptr())
        ev[2].record()
        hamm.best_finalize_device(g, best.data_ptr(), out.data_ptr())
        torch.cuda.synchronize(); secs.append(time.perf_counter() - t0)
    boot_ms, bands_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    info = hamm.last_diag_info()
    nv = G - K + 1
    bad = 0
    for b in (0, G // 3, G - B):
        hamm.allpairs_min_device(g, g, True, True, b, b + B, chk.data_ptr())
        torch.cuda.synchronize()
        bad += int((chk != out[b:b + B]).sum().item())
    print(json.dumps({"K": K, "seconds": round(secs[1], 3), "boot_ms": round(boot_ms, 1), "bands_ms": round(bands_ms, 1), 
                      "Gcmp_s": round(nv * nv * 2 / ((boot_ms + bands_ms) * 1e-3) / 1e9, 1), "launches": n, "counter_planes": info,
                      "crosscheck_mismatches_vs_popc": bad}), flush=True)
    g.free()
