"""Times full exhaustive passes with the diagonal-band engine (device-resident) on the bench
workloads; optional cross-check against the POPC engine on a query sample."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kit4b_b200 as k4b
from kit4b_b200 import hamm

k4b.gpu_init(1)
for wl in sys.argv[1:] or ["cfg1", "cfg2"]:
    concat, chroms, K, both = bench.synth_genome(wl)
    L = len(concat)
    nv = bench.valid_count(chroms, K)
    d_concat = torch.from_numpy(concat).cuda()
    g = hamm.Packed.from_device(d_concat.data_ptr(), L, K)
    best = torch.empty(L, dtype=torch.int32, device="cuda")
    out = torch.empty(L, dtype=torch.int16, device="cuda")
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hamm.best_init_device(best.data_ptr(), L, K)
        n = hamm.exhaustive_diag_device(g, both, 0, 1, best.data_ptr())
        hamm.best_finalize_device(g, best.data_ptr(), out.data_ptr())
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"workload": wl, "K": K, "rep": rep, "seconds": round(dt, 3), "band_ms": round(hamm.last_kernel_ms(), 1),
                          "launches": n, "Gcmp_s": round(nv * nv * (2 if both else 1) / dt / 1e9, 1)}), flush=True)
    # cross-check a sample of queries against the POPC engine
    B = 4096
    chk = torch.empty(B, dtype=torch.int16, device="cuda")
    bad = 0
    for b in (0, L // 3, L - B):
        hamm.allpairs_min_device(g, g, both, True, b, b + B, chk.data_ptr())
        torch.cuda.synchronize()
        bad += int((chk != out[b:b + B]).sum().item())
    print(json.dumps({"workload": wl, "crosscheck_mismatches": bad}), flush=True)
    g.free()
