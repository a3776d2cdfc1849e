"""Diagnostic: host-buffer (e2e) calls of the exhaustive shard API with per-call kernel time and
SM clocks sampled during the call."""
import sys, time, threading, numpy as np
sys.path.insert(0, '.')
import bench, kit4b_b200 as k4b, torch, pynvml
from kit4b_b200 import hamm
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
concat, chroms, K, both = bench.synth_genome('cfg2')
k4b.gpu_init(1)
pinned = torch.from_numpy(concat).pin_memory().numpy()
out = np.full(len(concat), K+1, np.uint16)
samples = []; stop = threading.Event()
def samp():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h)/1000.0))
        time.sleep(0.02)
th = threading.Thread(target=samp, daemon=True); th.start()
for i in range(6):
    t0 = time.perf_counter()
    hamm.exhaustive_shard(pinned, K, both, 1000+i*131072, 1000+(i+1)*131072, out)
    t1 = time.perf_counter()
    cl = [c for (t, c, p) in samples if t0 <= t <= t1]; pw = [p for (t, c, p) in samples if t0 <= t <= t1]
    print('call %d: %.3f s kernel_ms=%.1f clocks min/med/max=%s/%s/%s power max=%.0f W' % (i, t1-t0, hamm.last_kernel_ms(), min(cl), int(np.median(cl)), max(cl), max(pw)), flush=True)
stop.set()
