"""BASELINE configs[4]: K sweep 16/32/64/96/128 (+25, 50, 100, 200) over a 5 Mbp synthetic genome:
device-resident Gcmp/s and fraction of the measured POPC roofline for a query batch vs all
targets, both strands.  Prints one JSON object per K."""
import json, os, sys
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
peak = hamm.microbench_intpipe(0, 4000)
B = int(os.environ.get("KSWEEP_BATCH", "131072"))
out = torch.empty(B, dtype=torch.int16, device="cuda")
for K in [int(k) for k in (sys.argv[1:] or [16, 25, 32, 50, 64, 96, 100, 128, 200])]:
    p = hamm.Packed.from_device(d_concat.data_ptr(), G, K)
    ms = []
    for it in range(4):
        hamm.allpairs_min_device(p, p, True, True, 1000 + it * B, 1000 + (it + 1) * B, out.data_ptr())
        ms.append(hamm.last_kernel_ms())
    t = float(np.mean(ms[1:]))
    nt = G - K + 1
    W = (K + 31) // 32
    gcmp = B * nt * 2 / (t * 1e-3) / 1e9
    print(json.dumps({"K": K, "W": W, "kernel_ms": round(t, 2), "Gcmp_s": round(gcmp, 1),
                      "Gwc_s": round(gcmp * W, 1), "frac_of_popc_peak": round(gcmp * W / peak, 4),
                      "popc_peak": round(peak, 1)}), flush=True)
    p.free()
