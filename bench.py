#!/usr/bin/env python
"""bench.py - K-mer comparisons/s of the `hammings` hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] - `hammings -m1 -K50 -c` all-vs-all over a
10 Mbp synthetic bacterial-scale multifasta: 2e14 K-mer comparisons.  One *step* is that WHOLE
job on the diagonal-band engine (about 5 s on one B200).  With N GPUs the pair matrix of the
same job is partitioned over the ranks (strong scaling) after ONE NCCL broadcast of the packed
sequence set; the per-rank minima meet in all_reduce(MIN).  `--engine popc` benchmarks the
XOR/fold/POPC all-pairs kernel alone on query batches (a full pass would take minutes).

Printed JSON (one line, rank 0):
  value     Gcmp/s with inputs resident in HBM (CUDA events around exactly K steps, max over
            ranks); comparisons = valid queries x valid targets x strands
  e2e       same metric through the host-buffer C ABI / distributed host API: every step
            copies the 1-byte/base concat host->device, packs, (broadcasts), compares, and
            reads the minima back
  roofline  integer-pipe roofline of the dominant kernel: bands - ALU-pipe thread-ops (SHF/LOP3 per
            32-cell row step, counted from SASS) per second over the band launches' own CUDA-event
            time vs the LOP3 rate measured live; popc - word-compares per second vs the POPC rate
            measured live (1 POPC per 32-base word-compare)
  cpu_baseline  the reference's own CPU engine (oracle/_ref, unmodified sources) timed on this
            host on a bounded sample of the same workload
`--impl reference` times only that CPU engine, with all host threads, on the same config.
`--workload cfg4` runs BASELINE configs[3] (targeted mode, -m0 -I) on the seed-and-verify engine:
step = index of the 500 Mbp assembly + all 2e6 probe K-mers; its roofline entry is the HBM stream
of the index (warp-per-item schedule) or the POPC pipe (bucket-major join).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (chromosome lengths, K, both strands, seed)   [SURVEY.md 8d table]
    "cfg1": ([1_000_000], 25, True, 12),
    "cfg2": ([9_200_000, 400_000, 200_000, 100_000, 100_000], 50, True, 21),
    "cfg3": ([5_000_000] * 10, 100, True, 31),
    "cfg5k32": ([5_000_000], 32, True, 51),
}
WORKLOAD_DESCR = {
    "cfg1": "BASELINE configs[0]: hammings -m1 -K25 -c, 1 Mbp synthetic genome",
    "cfg2": "BASELINE configs[1]: hammings -m1 -K50 -c all-vs-all, 10 Mbp synthetic bacterial-scale multifasta (1x9.2 Mbp + 4 plasmid-like 0.1-0.4 Mbp)",
    "cfg3": "BASELINE configs[2]: hammings -m1 -K100 -c, 50 Mbp synthetic genome (10 x 5 Mbp)",
    "cfg5k32": "BASELINE configs[4]: hammings -m1 -K32 -c, 5 Mbp synthetic genome",
}


def synth_genome(name):
    lens, K, both, seed = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    parts, chroms, pos = [], [], 0
    for i, n in enumerate(lens):
        parts.append(rng.integers(0, 4, size=n, dtype=np.uint8))
        chroms.append(("chr%d" % (i + 1), pos, n))
        pos += n
        if i + 1 < len(lens):
            parts.append(np.array([7], dtype=np.uint8))
            pos += 1
    return np.ascontiguousarray(np.concatenate(parts)), chroms, K, both


def synth_targeted(scale=1.0):
    """BASELINE configs[3]: assembly of 20 x 25 Mbp (seed 41, every entry followed by
    EOS as in the .sfx sequence area), probes = 1 Mbp copied from the assembly with 3 % substitutions
    + 1 Mbp fresh random (seed 42); K=32, R=3, both strands."""
    rng = np.random.default_rng(41)
    nchr, clen = 20, int(25_000_000 * scale)
    parts = []
    for _ in range(nchr):
        parts += [rng.integers(0, 4, size=clen, dtype=np.uint8), np.array([7], dtype=np.uint8)]
    target = np.ascontiguousarray(np.concatenate(parts))
    rng = np.random.default_rng(42)
    pl = int(1_000_000 * scale)
    src = int(3.3 * clen) + 12345
    copy = target[src:src + pl].copy()
    idx = rng.choice(pl, size=int(0.03 * pl), replace=False)
    copy[idx] = (copy[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
    probes = np.ascontiguousarray(np.concatenate([copy, [7], rng.integers(0, 4, size=pl, dtype=np.uint8)]), dtype=np.uint8)
    return target, probes, 32, 3, True, nchr * (clen - 32 + 1), 2 * (pl - 32 + 1)


def valid_count(chroms, K, b=None, e=None):
    """number of valid K-mer starts (inside one chromosome) in flat range [b,e)"""
    tot = 0
    for _, start, n in chroms:
        lo, hi = start, start + max(0, n - K + 1)
        if b is not None:
            lo, hi = max(lo, b), min(hi, e)
        tot += max(0, hi - lo)
    return tot


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi fields of the profiling recipe, via NVML)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, threads=None, target_seconds=15.0, sweeps=None):
    """Times the reference's own CPU engine on a bounded sample of the workload: the leading
    sweep offsets of the exhaustive run (-b1 -B<n>), all host threads.  Returns a dict with
    Gcmp/s computed from the logical comparisons those sweeps cover."""
    from oracle import hamm_oracle as ho
    concat, chroms, K, both = synth_genome(workload)
    L = len(concat)
    cores = os.cpu_count() or 1
    T = min(128, threads or cores)  # reference caps worker threads at 128 (libkit4b/commdefs.h:191)
    N = valid_count(chroms, K)
    # reference cost: ~(L-s) Watson + ~L Crick cells per sweep offset at ~1.1e8 cells/s/core
    if sweeps is None:
        per_sweep_core_s = (2.0 * L if both else 1.0 * L) / 1.1e8
        sweeps = max(T, int(target_seconds * T / per_sweep_core_s))
        sweeps = min(sweeps, L - K)
    ref = ho.ref_binary(nosleep=True)

    def logical_cmps(n):
        s = np.arange(1, n + 1, dtype=np.float64)
        w = 2.0 * np.maximum(0.0, N - s).sum()          # each Watson cell serves (i,j) and (j,i)
        c = float(n) * N if both else 0.0                # each Crick cell serves one ordered pair
        return w + c

    if ref:
        with tempfile.TemporaryDirectory() as td:
            seq = os.path.join(td, "g.seq")
            entries = [(nm, concat[st:st + n]) for nm, st, n in chroms]
            ho.write_bioseq(seq, entries, title=workload)
            args = [ref, "hammings", "-m1", "-K%d" % K, "-T%d" % T, "-b1", "-B%d" % sweeps, "-i", seq]
            if both:
                args.insert(3, "-c")
            t0 = time.perf_counter()
            p = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
            dt = time.perf_counter() - t0
            if p.returncode != 0:
                raise RuntimeError("reference run failed: " + p.stdout.decode("latin-1")[-400:])
        kind = "reference"
        how = ("unmodified reference hammings (oracle/_ref, libc sleep() interposed so its fixed 10 s "
               "main-thread sleep is not billed) -m1 -K%d %s-T%d -b1 -B%d on the same genome: %d of %d "
               "sweep offsets incl. genome load, per-thread array init and min-merge"
               % (K, "-c " if both else "", T, sweeps, sweeps, L - K))
    else:
        t0 = time.perf_counter()
        total = (L - K) + (2 * (L - K) + 1 if both else 0)
        num = sweeps * (3 if both else 1)
        ho.exhaustive_sliding_sample(concat, K, both, T, num, total)
        dt = time.perf_counter() - t0
        kind = "port"
        how = "oracle C port (sliding diagonals), leading %d of %d diagonals, %d threads" % (num, total, T)
    cmps = logical_cmps(sweeps)
    return {"value": cmps / dt / 1e9, "unit": "Gcmp/s", "cores": T, "kind": kind, "sample": how,
            "seconds": dt, "sweeps": sweeps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(args.workload, target_seconds=args.cpu_seconds)
        if i >= args.warmup:
            vals.append((last["value"], last["seconds"]))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals])) * 1e3
    _, _, K, both = synth_genome(args.workload)
    line = {
        "impl": "reference", "metric": "kmer_comparisons_per_sec", "value": v, "unit": "Gcmp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_DESCR[args.workload], "K": K, "both_strands": both,
                   "step": "bounded sample: leading sweep offsets of the exhaustive run, all host threads"},
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": "Gcmp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def _setup(args):
    import torch
    import torch.distributed as dist
    import kit4b_b200 as k4b
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kit4b_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k4b.gpu_init(1, [local])
    return torch, dist, k4b, world, rank, local, dev


def _traffic(key):
    tr_path = os.path.join(ROOT, "profiles", "dram_bytes.json")
    try:
        return json.load(open(tr_path)).get(key)
    except Exception:
        return None


def run_ours_bands(args):
    """Default: one step = the WHOLE all-vs-all job of the workload on the diagonal-band engine
    (sharded bootstrap -> all_reduce(MIN) -> this rank's part of the pair matrix, all_reduce(MIN) after every slab
    -> finalize).  N GPUs split the same job: strong scaling."""
    torch, dist, k4b, world, rank, local, dev = _setup(args)
    from kit4b_b200 import hamm
    from kit4b_b200.dist import CudaEngine, bands_slabwise, exhaustive_distributed_bands, shard_bounds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    concat, chroms, K, both = synth_genome(args.workload)
    L = len(concat)
    S = 2 if both else 1
    Nv = valid_count(chroms, K)
    engine = CudaEngine(dev)
    # ---- setup (untimed): rank 0 packs, ONE NCCL broadcast of the packed sequence set ----
    if rank == 0:
        image, packed, non_acgt = engine.pack(concat, K)
        flag = torch.tensor([int(non_acgt)], dtype=torch.int64, device=dev)
    else:
        image = engine.empty_image(L)
        flag = torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(flag, src=0)
        dist.broadcast(image, src=0)
    if rank != 0:
        packed = engine.adopt(image, L, K, bool(flag.item()))
    torch.cuda.synchronize()

    best = torch.empty(L, dtype=torch.int32, device=dev)
    out = torch.empty(L, dtype=torch.int16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    qb, qe = shard_bounds(0, L, world)[rank]

    def step():
        n = 1
        hamm.best_init_device(best.data_ptr(), L, K, stream.cuda_stream)
        hamm.diag_bootstrap_device(packed, both, qb, qe, best.data_ptr(), stream.cuda_stream)
        n += 1
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MIN)
        n += bands_slabwise(engine, packed, both, rank, world, best)  # all_reduce(MIN) after every slab
        if rank == 0:
            hamm.best_finalize_device(packed, best.data_ptr(), out.data_ptr(), stream.cuda_stream)
            n += 1
        return n

    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, band_ms = 0, []
    ev0.record(stream)
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations (inside the bracket: ~0.1 ms)
        launches += step() + 1
        band_ms.append(hamm.last_kernel_ms())  # CUDA events around this rank's band launches
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    cmps_per_step = float(Nv) * float(Nv) * S
    value = cmps_per_step * args.steps / (ms_max * 1e-3) / 1e9
    checksum = int(out.to(torch.int64).sum().item()) if rank == 0 else 0

    # roofline of the dominant kernel (this rank's band launches of one step)
    k_ms = float(np.mean(band_ms))
    info = hamm.last_diag_info()                    # counter widths the last step actually ran with
    fr = info["narrow_slabs"] / max(1, info["slabs"])
    np_planes = info["np_small"] * fr + info["np_full"] * (1.0 - fr) if info["np_small"] else float(info["np_full"])
    ops_per_rowstep = 9.0 + np_planes              # per-thread ALU-pipe instructions per 32-cell row step (SASS: a row
                                                   # PAIR is 8 SHF + (2*NP+9) LOP3 + 1 ISETP on the ALU pipe; its 8 IMAD
                                                   # run on the FMA pipe, its 4 LDS.128 on the LSU)
    cells = cmps_per_step / 2.0 / world            # every cell serves the two K-mers of a pair
    achieved = cells / 32.0 * ops_per_rowstep / (k_ms * 1e-3) / 1e9
    peak = hamm.microbench_intpipe(1, 4000)        # measured LOP3 thread-ops/s on this GPU
    roofline = {"bound": "int_pipe(alu: lop3/shf)", "achieved": achieved, "peak": peak, "unit": "Gop/s",
                "frac": achieved / peak, "traffic": _traffic(args.workload + "_bands"),
                "kernel": "diag_min_kernel<NP=%d|%d,P=2>: %d of %d slabs ran the narrow-counter instance (all band "
                          "launches of one step on this rank)" % (info["np_small"], info["np_full"], info["narrow_slabs"],
                                                                 info["slabs"]),
                "kernel_ms": k_ms, "kernel_share_of_step": k_ms * args.steps / ms if ms > 0 else None,
                "ops_model": "%.1f ALU-pipe thread-ops per 32-cell row step (per row PAIR: 8 SHF window cuts + 2*NP+9 "
                             "LOP3 [pair-min signed-delta ripple] + 1 ISETP; the 8 broadcast XORs of a pair run as IMAD "
                             "on the FMA pipe, row bits come from a shared-memory table); cells = valid pairs, each "
                             "serving 2 comparisons; block prologues, warm-up rows and the flagged-cell slow path are "
                             "NOT counted as useful work" % ops_per_rowstep,
                "peak_source": "measured live: register-resident LOP3 microbenchmark (k4b_microbench_intpipe), "
                               "nominal 64 lanes/clk/SM x 148 SMs x 1.965 GHz = 18614",
                "hbm_note": "planes (%.1f MB) are L2 resident; HBM is not the bound" % (hamm.packed_image_bytes(L) / 1e6)}

    # ---- end-to-end through the host-buffer API ----
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else args.steps
    pinned = torch.from_numpy(concat).pin_memory()
    h_concat = pinned.numpy()

    def e2e_step():
        if world == 1:
            return hamm.exhaustive(h_concat, K, both)          # k4b_hamm_exhaustive, host buffers
        return exhaustive_distributed_bands(h_concat if rank == 0 else None, K, both, engine=engine)

    e2e_sum = None
    for i in range(args.warmup if e2e_steps else 0):
        e2e_step()
        engine.keep.clear()
    barrier()
    e2e_sampler = ClockSampler(local)  # the e2e leg follows ~40 s of sustained load: record its clocks too
    e2e_sampler.start()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        r = e2e_step()
        engine.keep.clear()
        if rank == 0:
            e2e_sum = int(r.astype(np.int64).sum())
    barrier()
    dt = time.perf_counter() - t0
    e2e_clocks = e2e_sampler.stop()
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = cmps_per_step * e2e_steps / float(tt.item()) / 1e9 if e2e_steps else None
    e2e = {"value": e2e_val, "unit": "Gcmp/s", "h2d_bytes_per_step": int(L), "d2h_bytes_per_step": int(2 * L),
           "steps": e2e_steps, "result_checksum_equals_resident_run": (e2e_sum == checksum) if e2e_steps else None,
           "clocks": e2e_clocks,
           "api": "k4b_hamm_exhaustive (host buffers)" if world == 1 else
                  "kit4b_b200.dist.exhaustive_distributed_bands (rank-0 host buffer, NCCL broadcast + all_reduce MIN after the bootstrap and every slab)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_sample(args.workload, target_seconds=args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "kmer_comparisons_per_sec", "value": value, "unit": "Gcmp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESCR[args.workload], "K": K, "both_strands": both,
                       "genome_bases": int(L), "kmers": int(Nv),
                       "step": "the whole all-vs-all job: every K-mer vs every K-mer, %s (%.3g comparisons)"
                               % ("both strands" if both else "Watson only", cmps_per_step),
                       "engine": "diagonal bands (bit-sliced sliding counters) bootstrapped by the POPC all-pairs kernel",
                       "parallelism": "pair-matrix partition x%d + all_reduce(MIN) per slab" % world,
                       "l2": "256 MB flush write between timed steps", "result_checksum": checksum},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    k4b.gpu_shutdown()


def run_ours_targeted(args):
    """--workload cfg4: targeted mode (-m0 -I) on the seed-and-verify engine.  One step = the WHOLE
    job (bucket index of the assembly + every probe K-mer, both strands); N GPUs split the probes
    (each builds the index from the broadcast planes; minima meet in one all_reduce MIN)."""
    torch, dist, k4b, world, rank, local, dev = _setup(args)
    from kit4b_b200 import hamm
    from kit4b_b200.dist import CudaEngine, shard_bounds

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    target, probes, K, R, both, Nt, Nq = synth_targeted(args.scale)
    core = K // (R + 1)
    clamp = K // core
    engine = CudaEngine(dev)
    imgs = []
    for concat in (target, probes):  # rank 0 packs, ONE broadcast of each packed set
        if rank == 0:
            image, packed, non_acgt = engine.pack(concat, K)
            flag = torch.tensor([int(non_acgt)], dtype=torch.int64, device=dev)
        else:
            image = engine.empty_image(len(concat))
            flag = torch.zeros(1, dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(flag, src=0)
            dist.broadcast(image, src=0)
        if rank != 0:
            packed = engine.adopt(image, len(concat), K, bool(flag.item()))
        imgs.append((image, packed))
    t_img, q_img = imgs[0][1], imgs[1][1]
    torch.cuda.synchronize()
    L = len(probes)
    best = torch.empty(L, dtype=torch.int32, device=dev)
    out = torch.empty(L, dtype=torch.int16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    qb, qe = shard_bounds(0, L, world)[rank]

    def step():
        hamm.best_init_device(best.data_ptr(), L, K, stream.cuda_stream)
        n = 1 + hamm.targeted_seed_device(q_img, t_img, both, clamp, core, qb, qe, best.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(best, op=dist.ReduceOp.MIN)
        if rank == 0:
            hamm.targeted_finalize_device(q_img, best.data_ptr(), clamp, out.data_ptr(), stream.cuda_stream)
            n += 1
        return n

    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, k_ms = 0, []
    ev0.record(stream)
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        launches += step() + 1
        k_ms.append(hamm.last_kernel_ms())
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    cmps = float(Nq) * float(Nt) * 2.0
    value = cmps * args.steps / (ms_max * 1e-3) / 1e9
    checksum = int(out.to(torch.int64).sum().item()) if rank == 0 else 0

    info = hamm.last_seed_info()
    kms = float(np.mean(k_ms))
    # algorithmic bytes of one step on this rank: query = 12 B per streamed bucket entry; index =
    # planes read twice (2 x 3/8 B per base) + 12 B written per indexed core
    alg_bytes = 12.0 * info["occurrences"] + 0.75 * len(target) + 12.0 * info["indexed_cores"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6552.3))
    bits = min(2 * core, 22)
    join = os.environ.get("K4B_SEED_JOIN", str(int((len(target) >> bits) >= 256))) != "0"
    if join:
        # bucket-major join: entries are staged once per 64 items in shared memory, so HBM is no longer the
        # bound; every entry test is 3 LOP3 + 1 POPC, and POPC (XU pipe, 16 lanes/clk/SM) is the scarcer pipe
        ppeak = hamm.microbench_intpipe(0, 4000)
        ach = info["occurrences"] / (kms * 1e-3) / 1e9
        roofline = {"bound": "int_pipe(xu: popc)", "achieved": ach, "peak": ppeak, "unit": "Gop/s", "frac": ach / ppeak,
                    "traffic": _traffic("cfg4_seed_join"),
                    "kernel": "seed_join_kernel (+ item keys, radix sort, index build) of one step on this rank",
                    "kernel_ms": kms, "kernel_share_of_step": kms * args.steps / ms if ms > 0 else None,
                    "ops_model": "1 POPC per entry test (%d tests: every item against every entry of its core's bucket); "
                                 "index build, item keys and the radix sort of the items are counted as overhead"
                                 % info["occurrences"],
                    "hbm_equivalent_GBps": alg_bytes / (kms * 1e-3) / 1e9,
                    "hbm_note": "the warp-per-item kernel streams 12 B per test from HBM (0.75 of the %.0f GB/s peak, "
                                "profiles/r01_bench_n1_cfg4_seed.json); the join reads each entry once per 64 items" % peak,
                    "peak_source": "measured live: register-resident POPC microbenchmark (k4b_microbench_intpipe)"}
    else:
        roofline = {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg_bytes / (kms * 1e-3) / 1e9 / peak, "traffic": _traffic("cfg4_seed"),
                    "kernel": "seed_query_kernel (+ seed_count/seed_fill index build) of one step on this rank",
                    "kernel_ms": kms, "kernel_share_of_step": kms * args.steps / ms if ms > 0 else None,
                    "bytes_model": "12 B per bucket entry streamed by the query kernel (%d entries) + index build: planes read "
                                   "twice, 12 B written per indexed core (%d cores)" % (info["occurrences"], info["indexed_cores"]),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (driver-written copy bandwidth)" if peaks else
                                   "fallback 6552.3 GB/s (MEASURED_PEAKS.json absent)"}

    e2e_steps = args.steps if args.e2e_steps is None else args.e2e_steps
    e2e_val, e2e_ok = None, None
    if e2e_steps and world == 1:
        k4b.targeted(target, probes, K, R, both)  # warm-up of the host path
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h = k4b.targeted(target, probes, K, R, both)
        dt = time.perf_counter() - t0
        e2e_val = cmps * e2e_steps / dt / 1e9
        e2e_ok = int(h[h != 0xFF].astype(np.int64).sum()) == int(out.cpu().numpy().view(np.uint16)[h != 0xFF].astype(np.int64).sum())
    if rank == 0:
        line = {
            "metric": "kmer_comparisons_per_sec", "value": value, "unit": "Gcmp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(1, args.steps),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: hammings -m0 -K32 -r3 -c -I probes: %d probe K-mers (1 Mbp mutated "
                                   "copy + 1 Mbp random, x%.2f) vs a %d-base synthetic assembly (20 entries)" % (Nq, args.scale, len(target)),
                       "K": K, "R": R, "both_strands": both, "probe_kmers": int(Nq), "target_kmers": int(Nt),
                       "step": "the whole targeted job: index of the assembly + every probe K-mer (%.3g logical comparisons)" % cmps,
                       "engine": "seed-and-verify (pigeonhole cores of %d bases, bucket index with flank signatures, %s)"
                                 % (core, "bucket-major join" if join else "warp per item"),
                       "parallelism": "probe shards x%d + all_reduce(MIN)" % world,
                       "l2": "256 MB flush write between timed steps", "result_checksum": checksum},
            "roofline": roofline, "cpu_baseline": None,
            "cpu_baseline_note": "the reference needs its suffix-array index for this mode; timed beside this engine by "
                                 "tools/cfg4_reference.py (profiles/r01_cfg4_reference*.log)",
            "e2e": {"value": e2e_val, "unit": "Gcmp/s", "h2d_bytes_per_step": int(len(target) + len(probes)),
                    "d2h_bytes_per_step": int(2 * L), "steps": e2e_steps if world == 1 else 0,
                    "result_checksum_equals_resident_run": e2e_ok, "api": "k4b_hamm_targeted (host buffers)"},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    k4b.gpu_shutdown()


def run_ours_popc(args):
    """--engine popc: the XOR/fold/POPC all-pairs kernel alone.  A full pass takes minutes, so a
    step is one query batch per GPU against all targets (per-query work is uniform); weak scaling."""
    torch, dist, k4b, world, rank, local, dev = _setup(args)
    from kit4b_b200 import hamm
    from kit4b_b200.dist import CudaEngine, exhaustive_distributed, shard_bounds
    hamm.set_engine(hamm.ENGINE_POPC)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    concat, chroms, K, both = synth_genome(args.workload)
    L = len(concat)
    S = 2 if both else 1
    W = (K + 31) // 32
    Nt = valid_count(chroms, K)
    B = min(args.batch, L)
    engine = CudaEngine(dev)
    if rank == 0:
        image, packed, non_acgt = engine.pack(concat, K)
        flag = torch.tensor([int(non_acgt)], dtype=torch.int64, device=dev)
    else:
        image = engine.empty_image(L)
        flag = torch.zeros(1, dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(flag, src=0)
        dist.broadcast(image, src=0)
    if rank != 0:
        packed = engine.adopt(image, L, K, bool(flag.item()))
    torch.cuda.synchronize()
    lo, hi = shard_bounds(0, L, world)[rank]

    def batch_range(i):
        span = max(1, hi - lo - B)
        b = lo + (i * B) % span
        return b, min(b + B, hi)

    out = torch.empty(B, dtype=torch.int16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def step(i):
        b, e = batch_range(i)
        n = engine.compute(packed, both, b, e, out)
        return n, valid_count(chroms, K, b, e)

    for i in range(args.warmup):
        flush.fill_(i & 0xFF)
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, nq_total, kernel_ms = 0, 0, []
    ev0.record(stream)
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        n, nq = step(args.warmup + i)
        launches += n + 1
        nq_total += nq
        kernel_ms.append(hamm.last_kernel_ms())
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    q = torch.tensor([float(nq_total)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(q, op=dist.ReduceOp.SUM)
    ms_max, nq_all = float(t.item()), float(q.item())
    value = nq_all * Nt * S / (ms_max * 1e-3) / 1e9
    k_ms = float(np.mean(kernel_ms))
    achieved = (nq_total / max(1, args.steps)) * Nt * S * W / (k_ms * 1e-3) / 1e9
    peak = hamm.microbench_intpipe(0, 4000)
    roofline = {"bound": "int_pipe(popc)", "achieved": achieved, "peak": peak, "unit": "Gwc/s",
                "frac": achieved / peak, "traffic": _traffic(args.workload + "_popc"),
                "kernel": "allpairs_min_kernel<W=%d>" % W, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * args.steps / ms if ms > 0 else None,
                "peak_source": "measured live: register-resident POPC microbenchmark, 1 POPC per 32-base "
                               "word-compare; nominal 16 lanes/clk/SM x 148 SMs x 1.965 GHz = 4654"}
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else args.steps
    pinned = torch.from_numpy(concat).pin_memory()
    h_concat = pinned.numpy()
    host_out = np.full(L, K + 1, dtype=np.uint16)

    def e2e_step(i):
        if world == 1:
            b, e = batch_range(i)
            hamm.exhaustive_shard(h_concat, K, both, b, e, host_out)
            return valid_count(chroms, K, b, e), (e - b) * 2
        gb = (i * world * B) % max(1, L - world * B)
        ge = min(gb + world * B, L)
        exhaustive_distributed(h_concat if rank == 0 else None, K, both, gb, ge, engine=engine)
        engine.keep.clear()
        return valid_count(chroms, K, gb, ge), (ge - gb) * 2

    for i in range(args.warmup if e2e_steps else 0):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    nq_e2e, d2h = 0, 0
    for i in range(e2e_steps):
        nq, nb = e2e_step(args.warmup + i)
        nq_e2e += nq
        d2h = nb
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e = {"value": (nq_e2e * Nt * S) / float(tt.item()) / 1e9 if e2e_steps else None, "unit": "Gcmp/s",
           "h2d_bytes_per_step": int(L), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
           "api": "k4b_hamm_exhaustive_shard (host buffers)" if world == 1 else
                  "kit4b_b200.dist.exhaustive_distributed (rank-0 host buffer, NCCL broadcast, gather)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_sample(args.workload, target_seconds=args.cpu_seconds)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        line = {
            "metric": "kmer_comparisons_per_sec", "value": value, "unit": "Gcmp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESCR[args.workload], "K": K, "both_strands": both,
                       "genome_bases": int(L), "target_kmers": int(Nt),
                       "step": "query batch of %d K-mers per GPU vs all targets, both strands" % B,
                       "engine": "POPC all-pairs kernel only (--engine popc)",
                       "global_batch_queries": int(B * world), "parallelism": "query-shard x%d" % world,
                       "l2": "256 MB flush write between timed steps"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    k4b.gpu_shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--scale", type=float, default=1.0, help="--workload cfg4: size factor of assembly and probes")
    ap.add_argument("--engine", default="bands", choices=["bands", "popc"],
                    help="bands: diagonal-band engine, step = whole job (default); popc: all-pairs POPC kernel, step = query batch")
    ap.add_argument("--batch", type=int, default=131072, help="--engine popc: query K-mers per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.workload == "cfg4":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "cfg4: the reference needs its suffix-array index; "
                              "see tools/cfg4_reference.py"}))
            return
        run_ours_targeted(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.engine == "popc":
        run_ours_popc(args)
    else:
        run_ours_bands(args)


if __name__ == "__main__":
    main()
