"""How much of a band pass is the flagged-cell slow path?  Runs the bands of a workload twice on
the same minima array: the first pass starts from the bootstrap (thresholds tighten slab by
slab), the second starts from the FINAL minima (thresholds as tight as they can be, almost no
flags).  The difference is what looser thresholds cost; env K4B_DIAG_SLABS / K4B_BOOT_TILES apply."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kit4b_b200 as k4b
from kit4b_b200 import hamm

k4b.gpu_init(1)
for wl in sys.argv[1:] or ["cfg2"]:
    concat, chroms, K, both = bench.synth_genome(wl)
    K = int(os.environ.get("K4B_PROBE_K", K))  # same genome, other K
    L = len(concat)
    d_concat = torch.from_numpy(concat).cuda()
    g = hamm.Packed.from_device(d_concat.data_ptr(), L, K)
    best = torch.empty(L, dtype=torch.int32, device="cuda")
    res = {"workload": wl, "K": K, "slabs": os.environ.get("K4B_DIAG_SLABS"), "boot": os.environ.get("K4B_BOOT_TILES")}
    for rep in range(2):
        hamm.best_init_device(best.data_ptr(), L, K)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        hamm.diag_bootstrap_device(g, both, 0, L, best.data_ptr())
        ev[1].record()
        hamm.diag_bands_device(g, both, 0, 1, best.data_ptr())
        torch.cuda.synchronize()
        first = hamm.last_kernel_ms()
        out = torch.empty(L, dtype=torch.int16, device="cuda")
        hamm.best_finalize_device(g, best.data_ptr(), out.data_ptr())
        chk = int(out.to(torch.int64).sum().item())  # finalized output (invalid starts masked)
        hamm.diag_bands_device(g, both, 0, 1, best.data_ptr())
        torch.cuda.synchronize()
        second = hamm.last_kernel_ms()
        res.update({"boot_ms": round(ev[0].elapsed_time(ev[1]), 1), "bands_ms": round(first, 1),
                    "bands_final_thresholds_ms": round(second, 1), "checksum": chk,
                    "unchanged_by_second_pass": None})
        hamm.best_finalize_device(g, best.data_ptr(), out.data_ptr())
        res["unchanged_by_second_pass"] = chk == int(out.to(torch.int64).sum().item())
    print(json.dumps(res), flush=True)
    g.free()
