#!/usr/bin/env python
"""bench.py - K-mer comparisons/s of the `hammings` hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs all|none|a,b,..]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (the top-level keys of the ONE JSON line rank 0 prints): BASELINE.json configs[1] -
`hammings -m1 -K50 -c` all-vs-all over a 10 Mbp synthetic bacterial-scale multifasta, 2e14 K-mer
comparisons.  One *step* is that WHOLE job on the diagonal-band engine.  With N GPUs the pair matrix
of the same job is partitioned over the ranks (strong scaling) after ONE NCCL broadcast of the packed
sequence set; the per-rank minima meet in all_reduce(MIN).

  value     Gcmp/s with inputs resident in HBM (CUDA events around exactly K steps, max over
            ranks); comparisons = valid queries x valid targets x strands
  e2e       same metric through the host-buffer C ABI (N=1: k4b_hamm_exhaustive) / the distributed
            host API (N>1): every step copies the 1-byte/base concat host->device from pinned
            memory, packs, (broadcasts), compares, and reads every minimum back
  roofline  integer-pipe roofline of the dominant kernel (diag_min_kernel): ALU-pipe thread-ops of
            the shipped inner loop per second over the band launches' own CUDA-event time vs the
            LOP3 rate measured live; `alu_pipe_busy_ncu` is the pipe counter of the committed
            ncu capture (profiles/)
  cpu_baseline  the reference's own CPU engine (oracle/_ref, unmodified sources) timed on this host
            on a bounded sample of the same workload (leading sweep offsets, -b1 -B<n>), fixed
            costs (load, per-thread array init, merge) measured by a 1-offset run and subtracted;
            the sample's output FILE is compared with `k4b_hammings -m1 -b1 -B<n>` on the same bioseq
  parity    the band result of the last timed step re-derived two other ways: >= 16 k sampled
            query K-mers by the POPC all-pairs kernel, a few by a NumPy brute force in this file
  configs   every other BASELINE config measured in the same process, each with value / e2e /
            roofline / result_checksum / parity: cfg1, cfg5 (K = 16,32,64,96,128), cfg4 (targeted,
            seed-and-verify engine), the POPC all-pairs kernel alone (SURVEY 8d roofline), cfg3
  e2e_inproc (N>1) rank 0 afterwards drives all N GPUs from ONE process through k4b_gpu_init(N) +
            k4b_hamm_exhaustive (the path `k4b_hammings --gpus N` uses), plus that CLI file to file

`--impl reference` times only the reference CPU engine, with all host threads, on the same config
(rank 0; other ranks exit 0).  `--workload cfg4` / `--engine popc` run those legs as the headline.
"""
import argparse
import concurrent.futures
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
    "cfg5k16": ([5_000_000], 16, True, 51),
    "cfg5k32": ([5_000_000], 32, True, 51),
    "cfg5k64": ([5_000_000], 64, True, 51),
    "cfg5k96": ([5_000_000], 96, True, 51),
    "cfg5k128": ([5_000_000], 128, True, 51),
}
WORKLOAD_DESCR = {
    "cfg1": "BASELINE configs[0]: hammings -m1 -K25 -c, 1 Mbp synthetic genome",
    "cfg2": "BASELINE configs[1]: hammings -m1 -K50 -c all-vs-all, 10 Mbp synthetic bacterial-scale multifasta (1x9.2 Mbp + 4 plasmid-like 0.1-0.4 Mbp)",
    "cfg3": "BASELINE configs[2]: hammings -m1 -K100 -c, 50 Mbp synthetic genome (10 x 5 Mbp)",
}
for _k in (16, 32, 64, 96, 128):
    WORKLOAD_DESCR["cfg5k%d" % _k] = "BASELINE configs[4]: hammings -m1 -K%d -c, 5 Mbp synthetic genome (K sweep)" % _k
ALL_CONFIGS = ["cfg1", "cfg5k16", "cfg5k32", "cfg5k64", "cfg5k96", "cfg5k128", "cfg4", "popc", "cfg3"]
CACHE = os.environ.get("K4B_BENCH_CACHE", os.path.join(tempfile.gettempdir(), "k4b_bench_cache"))


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


def valid_mask(concat, K):
    """bool[len]: a K-mer of one chromosome (no EOS in the window) starts here"""
    L = len(concat)
    eos = np.concatenate([[0], np.cumsum(concat == 7)])
    ok = np.zeros(L, dtype=bool)
    if L >= K:
        ok[:L - K + 1] = (eos[K:] - eos[:L - K + 1]) == 0
    return ok


def write_fasta(path, entries):
    """80-column FASTA of (name, codes 0..3) entries; NumPy only (500 Mbp in a few seconds)."""
    lut = np.frombuffer(b"ACGTNNN\n", dtype=np.uint8)
    with open(path, "wb") as f:
        for name, codes in entries:
            f.write((">%s\n" % name).encode())
            n = len(codes)
            full = n // 80 * 80
            if full:
                body = np.empty((full // 80, 81), dtype=np.uint8)
                body[:, :80] = lut[codes[:full]].reshape(-1, 80)
                body[:, 80] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(lut[codes[full:]].tobytes() + b"\n")


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
        return self

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the ONLY code in this file that touches oracle/)
# ------------------------------------------------------------------------------------------
def _ref_run(args, cwd=None):
    t0 = time.perf_counter()
    p = subprocess.run(args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + p.stdout.decode("latin-1")[-400:])
    return dt


class ReferenceSampler:
    """The reference's own CPU engine (oracle/_ref, unmodified sources; libc sleep() interposed so
    its fixed 10 s main-thread sleep is not billed) on a bounded sample of an exhaustive workload:
    the leading sweep offsets (-b1 -B<n>), all host threads.  The fixed costs of a run (process
    start, genome load, per-thread array initialisation, min-merge) are measured once by a
    1-offset run and subtracted, so `value` is the rate a full-size run would sustain."""

    def __init__(self, workload, threads=None):
        from oracle import hamm_oracle as ho
        self.ho = ho
        self.workload = workload
        self.concat, self.chroms, self.K, self.both = synth_genome(workload)
        self.L = len(self.concat)
        cores = os.cpu_count() or 1
        self.T = min(128, threads or cores)  # reference caps worker threads at 128 (libkit4b/commdefs.h:191)
        self.N = valid_count(self.chroms, self.K)
        self.ref = ho.ref_binary(nosleep=True)
        self.dir = tempfile.mkdtemp(prefix="k4b_ref_")
        self.seq = os.path.join(self.dir, "g.seq")
        ho.write_bioseq(self.seq, [(nm, self.concat[st:st + n]) for nm, st, n in self.chroms], title=workload)
        self.fixed_s = None

    def sweeps_for(self, target_seconds):
        # reference cost: ~(L-s) Watson + ~L Crick cells per sweep offset at ~1.1e8 cells/s/core
        per_sweep_core_s = (2.0 * self.L if self.both else 1.0 * self.L) / 1.1e8
        return int(min(max(self.T, int(target_seconds * self.T / per_sweep_core_s)), self.L - self.K))

    def logical_cmps(self, n):
        s = np.arange(1, n + 1, dtype=np.float64)
        w = 2.0 * np.maximum(0.0, self.N - s).sum()      # each Watson cell serves (i,j) and (j,i)
        c = float(n) * self.N if self.both else 0.0       # each Crick cell serves one ordered pair
        return w + c

    def args(self, sweeps, out_csv=None):
        a = [self.ref, "hammings", "-m1", "-K%d" % self.K] + (["-c"] if self.both else []) + \
            ["-T%d" % self.T, "-b1", "-B%d" % sweeps, "-i", self.seq]
        return a + (["-o", out_csv] if out_csv else [])

    def run(self, sweeps, out_csv=None):
        if self.ref:
            return _ref_run(self.args(sweeps, out_csv))
        t0 = time.perf_counter()  # oracle port (only when the reference binary is absent)
        total = (self.L - self.K) + (2 * (self.L - self.K) + 1 if self.both else 0)
        self.ho.exhaustive_sliding_sample(self.concat, self.K, self.both, self.T, sweeps * (3 if self.both else 1), total)
        return time.perf_counter() - t0

    def fixed(self):
        if self.fixed_s is None:
            self.fixed_s = min(self.run(1), self.run(1)) if self.ref else 0.0
        return self.fixed_s

    def sample(self, target_seconds, out_csv=None):
        sweeps = self.sweeps_for(target_seconds)
        dt = self.run(sweeps, out_csv)
        net = max(dt - self.fixed(), 1e-3)
        cmps = self.logical_cmps(sweeps)
        kind = "reference" if self.ref else "port"
        how = ("unmodified reference hammings (oracle/_ref) -m1 -K%d %s-T%d -b1 -B%d on the same genome: %d of %d sweep "
               "offsets in %.2f s, minus %.2f s of fixed cost (load, per-thread array init, merge: a 1-offset run)"
               % (self.K, "-c " if self.both else "", self.T, sweeps, sweeps, self.L - self.K, dt, self.fixed())
               if self.ref else "oracle C port (sliding diagonals), leading diagonals, %d threads" % self.T)
        return {"value": cmps / net / 1e9, "value_incl_fixed_costs": cmps / dt / 1e9, "unit": "Gcmp/s", "cores": self.T,
                "kind": kind, "sample": how, "seconds": dt, "fixed_seconds": self.fixed(), "sweeps": sweeps}

    def file_parity(self, target_seconds):
        """reference sample WITH -o, then `k4b_hammings -m1 -b1 -B<n>` (the product CLI, GPU) on the same
        bioseq: the north star's check - a file diff - at this workload's full size."""
        from kit4b_b200 import hostlib
        sweeps = self.sweeps_for(target_seconds)
        ref_csv, our_csv = os.path.join(self.dir, "ref.csv"), os.path.join(self.dir, "ours.csv")
        if not self.ref:
            return {"ran": False, "why": "reference binary absent"}
        self.run(sweeps, ref_csv)
        a = [hostlib.cli_path(), "hammings", "-m1", "-K%d" % self.K] + (["-c"] if self.both else []) + \
            ["-b1", "-B%d" % sweeps, "-i", self.seq, "-o", our_csv]
        t0 = time.perf_counter()
        p = subprocess.run(a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            return {"ran": False, "why": "k4b_hammings failed: " + p.stdout.decode("latin-1")[-300:]}
        same = open(ref_csv, "rb").read() == open(our_csv, "rb").read()
        res = {"ran": True, "identical": bool(same), "bytes": os.path.getsize(ref_csv), "sweeps": sweeps,
               "ours_cli_seconds": round(dt, 2),
               "what": "reference `hammings -m1 -K%d %s-b1 -B%d -o ref.csv` vs `k4b_hammings` with the same flags on the "
                       "same bioseq: files compared byte for byte" % (self.K, "-c " if self.both else "", sweeps)}
        for f in (ref_csv, our_csv):
            os.unlink(f)
        return res

    def close(self):
        import shutil
        shutil.rmtree(self.dir, ignore_errors=True)


def cfg4_reference(scale, target_seconds=8.0, parity=True):
    """BASELINE configs[3] on the reference: `index` of the synthetic assembly once (cached under
    K4B_BENCH_CACHE), then `hammings -m0 -K32 -r3 -c -I <sample of the probes>` with all host threads;
    fixed costs (suffix-array load) measured by a 1-K-mer probe run and subtracted.  Optionally the
    product CLI on the same files and a byte compare of the outputs."""
    from oracle import hamm_oracle as ho
    from kit4b_b200 import hostlib
    ref = ho.ref_binary(nosleep=True)
    if not ref:
        return None
    target, probes, K, R, both, Nt, Nq = synth_targeted(scale)
    cores = os.cpu_count() or 1
    T = min(64, cores)
    d = os.path.join(CACHE, "cfg4_x%g" % scale)
    os.makedirs(d, exist_ok=True)
    sfx = os.path.join(d, "asm.sfx")
    index_s = None
    if not os.path.exists(sfx):
        nchr = 20
        clen = (len(target) // nchr) - 1
        write_fasta(os.path.join(d, "asm.fa"), [("chr%d" % (i + 1), target[i * (clen + 1):i * (clen + 1) + clen]) for i in range(nchr)])
        index_s = _ref_run([ref, "index", "-i", "asm.fa", "-o", "asm.sfx.tmp", "-r", "asm", "-T%d" % T], cwd=d)
        os.replace(os.path.join(d, "asm.sfx.tmp"), sfx)
        os.unlink(os.path.join(d, "asm.fa"))
    pl = (len(probes) - 1) // 2
    # reference cost measured on the GPU box: ~2.4 ms per probe K-mer of this sample on 16 threads (the mutated-copy half
    # descends the whole cascade; round 1's full job averaged 0.27 ms over all probes)
    n_half = int(max(200, min(pl - K, target_seconds * T / 16.0 / 2.4e-3 * min(1.0, 1.0 / max(scale, 1e-3)) / 2)))
    def probe_files(tag, n):
        fa, seq = os.path.join(d, tag + ".fa"), os.path.join(d, tag + ".seq")
        write_fasta(fa, [("mutated_copy", probes[:n + K - 1]), ("unrelated", probes[pl + 1:pl + 1 + n + K - 1])])
        _ref_run([ref, "genbioseq", "-i", fa, "-o", seq, "-r", tag])
        return seq
    seq_big, seq_one = probe_files("probes_sample", n_half), probe_files("probes_one", 1)
    base = [ref, "hammings", "-m0", "-K%d" % K, "-r%d" % R, "-c", "-T%d" % T, "-i", sfx]
    ref_csv, our_csv = os.path.join(d, "ref.csv"), os.path.join(d, "ours.csv")
    fixed = _ref_run(base + ["-I", seq_one, "-o", os.path.join(d, "one.csv")])
    dt = _ref_run(base + ["-I", seq_big, "-o", ref_csv])
    nq = 2 * n_half
    cmps = float(nq) * Nt * 2.0
    res = {"value": cmps / max(dt - fixed, 1e-3) / 1e9, "value_incl_fixed_costs": cmps / dt / 1e9, "unit": "Gcmp/s", "cores": T,
           "kind": "reference", "seconds": dt, "fixed_seconds": fixed, "index_seconds": index_s,
           "sample": "unmodified reference (oracle/_ref): `index` of the %d-base assembly (%s), then hammings -m0 -K%d -r%d -c "
                     "-T%d -I <%d probe K-mers: the first %d of the mutated copy and of the unrelated part> in %.2f s, minus "
                     "%.2f s of fixed cost (suffix-array load: a 1-K-mer probe run)"
                     % (len(target), "%.1f s" % index_s if index_s else "cached", K, R, T, nq, n_half, dt, fixed)}
    if parity:
        a = [hostlib.cli_path(), "hammings", "-m0", "-K%d" % K, "-r%d" % R, "-c", "-i", sfx, "-I", seq_big, "-o", our_csv]
        t0 = time.perf_counter()
        p = subprocess.run(a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        res["file_parity"] = {"ran": p.returncode == 0, "ours_cli_seconds": round(time.perf_counter() - t0, 2),
                              "identical": p.returncode == 0 and open(ref_csv, "rb").read() == open(our_csv, "rb").read(),
                              "what": "reference vs `k4b_hammings -m0 -K32 -r3 -c -i asm.sfx -I probes_sample.seq`: output files compared byte for byte"}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "cfg4":
        vals = []
        for i in range(args.warmup + args.steps):
            r = cfg4_reference(args.scale, target_seconds=args.cpu_seconds, parity=False)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref reference binary absent"}))
                return
            if i >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        line = {"impl": "reference", "metric": "kmer_comparisons_per_sec", "value": v, "unit": "Gcmp/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean([r["seconds"] for r in vals])) * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "BASELINE configs[3]: hammings -m0 -K32 -r3 -c -I probes (x%.2f)" % args.scale,
                           "step": "bounded sample of the probe K-mers against the full index, all host threads"},
                "cpu_baseline": {k: vals[-1][k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": v, "unit": "Gcmp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    rs = ReferenceSampler(args.workload)
    try:
        vals, last = [], None
        for i in range(args.warmup + args.steps):
            last = rs.sample(args.cpu_seconds)
            if i >= args.warmup:
                vals.append(last)
        v = float(np.mean([r["value"] for r in vals]))
        line = {
            "impl": "reference", "metric": "kmer_comparisons_per_sec", "value": v, "unit": "Gcmp/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean([r["seconds"] for r in vals])) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESCR[args.workload], "K": rs.K, "both_strands": rs.both,
                       "step": "bounded sample: leading sweep offsets of the exhaustive run, all host threads; value = "
                               "logical comparisons of the sample / (wall time - fixed cost of a 1-offset run)"},
            "value_incl_fixed_costs": float(np.mean([r["value_incl_fixed_costs"] for r in vals])),
            "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": "Gcmp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
    finally:
        rs.close()


# ------------------------------------------------------------------------------------------
# our arm: context
# ------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx():
    import torch
    import torch.distributed as dist
    import kit4b_b200 as k4b
    from kit4b_b200 import hamm
    from kit4b_b200 import dist as kdist
    c = Ctx()
    c.torch, c.dist, c.k4b, c.hamm, c.kdist = torch, dist, k4b, hamm, kdist
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kit4b_b200 has no CPU fallback")
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    k4b.gpu_init(1, [c.local])
    c.engine = kdist.CudaEngine(c.dev)
    c.stream = torch.cuda.current_stream(c.dev)
    c.flush = torch.empty(256 << 20, dtype=torch.uint8, device=c.dev)  # > 126 MB L2
    c.peaks = {}
    try:
        c.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    c.lop3_peak = hamm.microbench_intpipe(1, 4000)  # measured LOP3 thread-ops/s on this GPU (Gop/s)
    c.popc_peak = hamm.microbench_intpipe(0, 4000)
    return c


def barrier(c):
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def max_over_ranks(c, x):
    t = c.torch.tensor([x], dtype=c.torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_packed(c, concat, K):
    """untimed setup of a resident leg: rank 0 packs, ONE NCCL broadcast of the packed set"""
    packed, _, _ = c.kdist._broadcast_packed(c.engine, concat if c.rank == 0 else None, K, c.rank, c.world, None)
    c.torch.cuda.synchronize()
    return packed


def _profile_note(key):
    """figures of the committed ncu --set full capture of a kernel (profiles/dram_bytes.json, written
    by tools/ncu_summary.py): per-launch DRAM traffic and pipe utilisation"""
    try:
        v = json.load(open(os.path.join(ROOT, "profiles", "dram_bytes.json"))).get(key)
    except Exception:
        return {}
    if isinstance(v, (int, float)):
        return {"dram_bytes_per_launch": v}
    return v or {}


# ------------------------------------------------------------------------------------------
# exhaustive mode on the band engine
# ------------------------------------------------------------------------------------------
def bands_resident(c, packed, L, K, both, steps, warmup):
    """W untimed + K timed whole-job steps with the packed set resident in HBM."""
    torch, hamm = c.torch, c.hamm
    best = torch.empty(L, dtype=torch.int32, device=c.dev)
    out = torch.empty(L, dtype=torch.int16, device=c.dev)
    qb, qe = c.kdist.shard_bounds(0, L, c.world)[c.rank]

    def step():
        n = 1
        hamm.best_init_device(best.data_ptr(), L, K, c.stream.cuda_stream)
        hamm.diag_bootstrap_device(packed, both, qb, qe, best.data_ptr(), c.stream.cuda_stream)
        n += 1
        if c.world > 1:
            c.dist.all_reduce(best, op=c.dist.ReduceOp.MIN)
        n += c.kdist.bands_slabwise(c.engine, packed, both, c.rank, c.world, best)
        if c.rank == 0:
            hamm.best_finalize_device(packed, best.data_ptr(), out.data_ptr(), c.stream.cuda_stream)
            n += 1
        return n

    for i in range(warmup):
        c.flush.fill_(i & 0xFF)
        step()
    barrier(c)
    sampler = ClockSampler(c.local).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, band_ms = 0, []
    ev0.record(c.stream)
    for i in range(steps):
        c.flush.fill_(i & 0xFF)  # L2 flush between timed iterations (inside the bracket: ~0.1 ms)
        launches += step() + 1
        band_ms.append(hamm.last_kernel_ms())  # CUDA events around this rank's band launches
    ev1.record(c.stream)
    barrier(c)
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    return {"ms": ms, "ms_max": max_over_ranks(c, ms), "launches": launches, "band_ms": float(np.mean(band_ms)),
            "clocks": clocks, "out": out, "info": hamm.last_diag_info()}


def bands_roofline(c, workload, res, cmps_per_step, steps, L, three_planes):
    info = res["info"]
    fr = info["narrow_slabs"] / max(1, info["slabs"])
    np_planes = info["np_small"] * fr + info["np_full"] * (1.0 - fr) if info["np_small"] else float(info["np_full"])
    # per-thread ALU-pipe instructions per 32-cell row step, counted in the SASS of the shipped inner loop:
    # a row PAIR of the two-plane window-table instance is 4 SHF (window cuts) + (2*NP + 8) LOP3 (pair-min
    # signed-delta ripple) + 1 ISETP; its 4 LDS.64 run on the LSU, its address arithmetic on the uniform datapath.
    # (three-plane sets keep the row-table instance: 8 SHF + (2*NP + 13) LOP3 + 1 ISETP per pair incl. the N plane)
    ops_per_rowstep = (2.0 * np_planes + 13.0) / 2.0 if not three_planes else (2.0 * np_planes + 26.0) / 2.0
    cells = cmps_per_step / 2.0 / c.world  # every cell serves the two K-mers of a pair
    k_ms = res["band_ms"]
    achieved = cells / 32.0 * ops_per_rowstep / (k_ms * 1e-3) / 1e9
    prof = _profile_note(workload + "_bands")
    return {"bound": "int_pipe(alu: lop3/shf)", "achieved": achieved, "peak": c.lop3_peak, "unit": "Gop/s",
            "frac": achieved / c.lop3_peak, "traffic": prof.get("dram_bytes_per_launch"),
            "alu_pipe_busy_ncu": prof.get("alu_pipe_busy"), "profile": prof.get("source"),
            "kernel": "%s, NP=%d|%d, %d planes: %d of %d slabs ran the narrow-counter instance (all band launches of one step "
                      "on this rank)" % ("diag_min2_kernel (two diagonal words per thread) where NP <= 6 and the launch fills the "
                                         "GPU, else diag_min_kernel (window-table instance)" if not three_planes else
                                         "diag_min_kernel (row-table instance)", info["np_small"], info["np_full"],
                                         3 if three_planes else 2, info["narrow_slabs"], info["slabs"]),
            "kernel_ms": k_ms, "kernel_share_of_step": k_ms * steps / res["ms"] if res["ms"] > 0 else None,
            "ops_model": "%.2f ALU-pipe thread-ops per 32-cell row step = the inner loop of the shipped kernel (per row PAIR: "
                         "4 SHF + 2*NP+8 LOP3 + 1 ISETP); cells = valid pairs, each serving 2 comparisons.  NOT counted as "
                         "useful: the prologue of every 32-row block, the K warm-up rows of each 8193-row segment, cells "
                         "outside the triangle in boundary CTAs, the flagged-cell slow path, the bootstrap.  This is the "
                         "builder's instruction model of its own algorithm, not SURVEY 8d's POPC word-compare contract (by "
                         "which this engine sits far above 1: it does O(1) work per pair instead of K/32 POPCs; see the "
                         "`popc` entry of configs for the kernel that implements 8d)" % ops_per_rowstep,
            "peak_source": "measured live: register-resident LOP3 microbenchmark (k4b_microbench_intpipe), nominal 64 "
                           "lanes/clk/SM x 148 SMs x 1.965 GHz = 18614",
            "hbm_note": "planes (%.1f MB) are L2 resident; HBM is not the bound" % (c.hamm.packed_image_bytes(L) / 1e6)}


def numpy_min_distance(concat, valid, K, q, both, self_pos):
    """Independent brute force (no library of this repo): minimum distance of the K-mer q to every
    valid K-mer of concat - forward strand excluding the K-mer's own position, reverse complement
    including it (hammings.cpp:2692, :3300-3489) - in K passes over the array."""
    M = len(concat) - K + 1
    best = K + 1
    for strand in range(2 if both else 1):
        kq = q if strand == 0 else (3 - q[::-1])
        acc = np.zeros(M, dtype=np.uint8 if K < 255 else np.uint16)
        for p in range(K):
            acc += concat[p:p + M] != kq[p]
        acc = acc.astype(np.int32)
        acc[~valid[:M]] = K + 1
        if strand == 0 and self_pos is not None:
            acc[self_pos] = K + 1
        best = min(best, int(acc.min()))
    return best


def parity_exhaustive(c, concat, chroms, K, both, packed, out, n_ranges=64, n_brute=8, seed=5):
    """rank 0: the minima in `out` (device int16, this run's result) against (a) the POPC all-pairs
    kernel on n_ranges x 256 sampled query positions and (b) a NumPy brute force on n_brute K-mers."""
    torch, hamm = c.torch, c.hamm
    if c.rank != 0:
        return None
    L = len(concat)
    rng = np.random.default_rng(seed)
    starts = sorted(set(int(v) for v in rng.integers(0, max(1, L - 256), size=n_ranges)) | {0, max(0, L - 256)})
    tmp = torch.empty(256, dtype=torch.int16, device=c.dev)
    bad = 0
    for b in starts:
        e = min(b + 256, L)
        hamm.allpairs_min_device(packed, packed, both, True, b, e, tmp.data_ptr(), 0, c.stream.cuda_stream)
        bad += int((tmp[:e - b] != out[b:e]).sum().item())
    host = out.cpu().numpy().view(np.uint16)
    valid = valid_mask(concat, K)
    cand = [p for p in (int(v) for v in rng.integers(0, L - K, size=4 * n_brute)) if valid[p] and (concat[p:p + K] < 4).all()][:n_brute]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        want = list(ex.map(lambda p: numpy_min_distance(concat, valid, K, concat[p:p + K], both, p), cand))
    bad_brute = sum(int(host[p] != w) for p, w in zip(cand, want))
    return {"ok": bad == 0 and bad_brute == 0, "popc_kernel_queries": int(256 * len(starts)), "popc_kernel_mismatches": bad,
            "numpy_brute_force_queries": len(cand), "numpy_brute_force_mismatches": bad_brute}


def bands_e2e(c, concat, K, both, steps, warmup):
    """the whole job through the host-buffer API, pinned host buffers, copies inside the timed region"""
    torch, hamm = c.torch, c.hamm
    h_concat = torch.from_numpy(concat).pin_memory().numpy()
    L = len(concat)

    def e2e_step():
        if c.world == 1:
            return hamm.exhaustive(h_concat, K, both)  # k4b_hamm_exhaustive, host buffers
        return c.kdist.exhaustive_distributed_bands(h_concat if c.rank == 0 else None, K, both, engine=c.engine)

    for _ in range(warmup):
        e2e_step()
    barrier(c)
    sampler = ClockSampler(c.local).start()
    t0 = time.perf_counter()
    r = None
    for _ in range(steps):
        r = e2e_step()
    barrier(c)
    dt = max_over_ranks(c, time.perf_counter() - t0)
    clocks = sampler.stop()
    checksum = int(r.astype(np.int64).sum()) if (c.rank == 0 and r is not None) else None  # outside the timed region
    return {"seconds": dt, "steps": steps, "checksum": checksum, "clocks": clocks, "h2d": int(L), "d2h": int(2 * L),
            "api": "k4b_hamm_exhaustive (host buffers)" if c.world == 1 else
                   "kit4b_b200.dist.exhaustive_distributed_bands (rank-0 host buffer, NCCL broadcast, all_reduce(MIN) after "
                   "the bootstrap and - overlapped with the next slab - after every slab)"}


def bands_leg(c, workload, steps, warmup, e2e_steps, e2e_warmup, n_brute=8):
    """one exhaustive config on the band engine: resident value, e2e, roofline, checksum, parity"""
    torch = c.torch
    concat, chroms, K, both = synth_genome(workload)
    L = len(concat)
    Nv = valid_count(chroms, K)
    cmps = float(Nv) * float(Nv) * (2 if both else 1)
    packed = broadcast_packed(c, concat, K)
    res = bands_resident(c, packed, L, K, both, steps, warmup)
    value = cmps * steps / (res["ms_max"] * 1e-3) / 1e9
    checksum = int(res["out"].to(torch.int64).sum().item()) if c.rank == 0 else 0
    roofline = bands_roofline(c, workload, res, cmps, steps, L, bool(packed.has_non_acgt))
    parity = parity_exhaustive(c, concat, chroms, K, both, packed, res["out"], n_brute=n_brute)
    barrier(c)
    e2e = None
    if e2e_steps:
        r = bands_e2e(c, concat, K, both, e2e_steps, e2e_warmup)
        e2e = {"value": cmps * r["steps"] / r["seconds"] / 1e9, "unit": "Gcmp/s", "h2d_bytes_per_step": r["h2d"],
               "d2h_bytes_per_step": r["d2h"], "steps": r["steps"], "warmup": e2e_warmup,
               "result_checksum_equals_resident_run": (r["checksum"] == checksum) if c.rank == 0 else None,
               "clocks": r["clocks"], "api": r["api"]}
    packed.free()
    return {"workload": workload, "descr": WORKLOAD_DESCR[workload], "K": K, "both": both, "L": L, "Nv": Nv, "cmps": cmps,
            "value": value, "ms_per_step": res["ms_max"] / max(1, steps), "steps": steps, "warmup": warmup,
            "launches": res["launches"], "clocks": res["clocks"], "checksum": checksum, "roofline": roofline,
            "parity": parity, "e2e": e2e, "concat": concat, "chroms": chroms}


def compact(leg):
    """entry of the `configs` array"""
    return {"workload": leg["descr"], "K": leg["K"], "value": leg["value"], "unit": "Gcmp/s", "ms_per_step": leg["ms_per_step"],
            "steps": leg["steps"], "warmup": leg["warmup"], "comparisons_per_step": leg["cmps"],
            "e2e": leg["e2e"], "roofline": {k: leg["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic",
                                                                              "kernel", "kernel_ms", "kernel_share_of_step")},
            "result_checksum": leg["checksum"], "parity": leg["parity"], "clocks": leg["clocks"]}


# ------------------------------------------------------------------------------------------
# targeted mode (cfg4) on the seed-and-verify engine
# ------------------------------------------------------------------------------------------
def targeted_leg(c, scale, steps, warmup, e2e_steps, with_cpu):
    torch, hamm = c.torch, c.hamm
    target, probes, K, R, both, Nt, Nq = synth_targeted(scale)
    core = K // (R + 1)
    clamp = K // core
    t_img = broadcast_packed(c, target, K)
    q_img = broadcast_packed(c, probes, K)
    L = len(probes)
    best = torch.empty(L, dtype=torch.int32, device=c.dev)
    out = torch.empty(L, dtype=torch.int16, device=c.dev)

    def step():  # this rank: its share of the index buckets, every probe K-mer
        hamm.best_init_device(best.data_ptr(), L, K, c.stream.cuda_stream)
        n = 1 + hamm.targeted_seed_part_device(q_img, t_img, both, clamp, core, c.rank, c.world, best.data_ptr(),
                                               c.stream.cuda_stream)
        if c.world > 1:
            c.dist.all_reduce(best, op=c.dist.ReduceOp.MIN)
        if c.rank == 0:
            hamm.targeted_finalize_device(q_img, best.data_ptr(), clamp, out.data_ptr(), c.stream.cuda_stream)
            n += 1
        return n

    for i in range(warmup):
        c.flush.fill_(i & 0xFF)
        step()
    barrier(c)
    sampler = ClockSampler(c.local).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, k_ms = 0, []
    ev0.record(c.stream)
    for i in range(steps):
        c.flush.fill_(i & 0xFF)
        launches += step() + 1
        k_ms.append(hamm.last_kernel_ms())
    ev1.record(c.stream)
    barrier(c)
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    ms_max = max_over_ranks(c, ms)
    cmps = float(Nq) * float(Nt) * 2.0
    value = cmps * steps / (ms_max * 1e-3) / 1e9
    checksum = int(out.to(torch.int64).sum().item()) if c.rank == 0 else 0
    info = hamm.last_seed_info()
    kms = float(np.mean(k_ms))
    bits = min(2 * core, 22)
    join = os.environ.get("K4B_SEED_JOIN", str(int((len(target) >> bits) >= 256))) != "0"
    hbm_peak = float(c.peaks.get("hbm_gbs", 6552.3))
    alg_bytes = 16.0 * info["occurrences"] + 0.75 * len(target) + 16.0 * info["indexed_cores"]
    prof = _profile_note("cfg4_seed_join" if join else "cfg4_seed")
    if join:
        ach = info["occurrences"] / (kms * 1e-3) / 1e9
        roofline = {"bound": "int_pipe(xu: popc)", "achieved": ach, "peak": c.popc_peak, "unit": "Gop/s", "frac": ach / c.popc_peak,
                    "traffic": prof.get("dram_bytes_per_launch"), "profile": prof.get("source"),
                    "kernel": "seed_join_kernel, one pass per core (+ index build, item keys, radix sorts of the items) of one step on this rank",
                    "kernel_ms": kms, "kernel_share_of_step": kms * steps / ms if ms > 0 else None,
                    "ops_model": "1 POPC per entry test (%d tests: every item the phase schedule leaves active against every entry "
                                 "of its core's bucket); index build, item keys and the radix sorts of the items count as "
                                 "overhead" % info["occurrences"],
                    "peak_source": "measured live: register-resident POPC microbenchmark (k4b_microbench_intpipe)"}
    else:
        roofline = {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "traffic": prof.get("dram_bytes_per_launch"),
                    "kernel": "seed_query_kernel (+ index build) of one step on this rank", "kernel_ms": kms,
                    "kernel_share_of_step": kms * steps / ms if ms > 0 else None,
                    "bytes_model": "16 B per bucket entry streamed by the query kernel (%d entries) + index build: planes read "
                                   "twice, 16 B written per indexed core (%d cores)" % (info["occurrences"], info["indexed_cores"]),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if c.peaks else "fallback 6552.3 GB/s (MEASURED_PEAKS.json absent)"}
    # the HBM-bound part of the job on its own: the index build, timed as a call that joins ONE probe K-mer
    index_build = None
    if c.world == 1:
        try:
            scratch = torch.empty(L, dtype=torch.int32, device=c.dev)
            i_ms = []
            for _ in range(3):
                hamm.best_init_device(scratch.data_ptr(), L, K, c.stream.cuda_stream)
                hamm.targeted_seed_device(q_img, t_img, both, clamp, core, 0, 1, scratch.data_ptr(), c.stream.cuda_stream)
                c.stream.synchronize()
                i_ms.append(hamm.last_kernel_ms())
            cores = hamm.last_seed_info()["indexed_cores"]
            partition = os.environ.get("K4B_SEED_INDEX", "1") != "0" and bits <= 16
            i_bytes = (48.0 if partition else 16.0) * cores + 0.75 * len(target)
            i_s = min(i_ms) * 1e-3
            index_build = {"bound": "hbm", "ms": min(i_ms), "achieved": i_bytes / i_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": i_bytes / i_s / 1e9 / hbm_peak, "indexed_cores": cores,
                           "kernel": ("seed_count_smem_kernel + seed_part_a_kernel + seed_part_b_persistent_kernel" if partition
                                      else "seed_scan_kernel<count> + seed_scan_kernel<fill>"),
                           "bytes_model": ("planes read twice (3/8 B per base each) + per indexed core 16 B written by the first "
                                           "partition pass, 16 B read + 16 B written by the second" if partition else
                                           "planes read twice + 16 B written per indexed core"),
                           "timed": "engine's own CUDA events around a call that builds the whole index and joins one probe "
                                    "K-mer (includes ~1 ms of fixed launches: probe reverse complement, item keys, sort, join); "
                                    "best of 3",
                           "profile": "profiles/r02_ncu_full_seed_index_partition_v29_summary.json" if partition else
                                      "profiles/r02_ncu_full_seed_scan_summary.json",
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if c.peaks else "fallback 6552.3 GB/s"}
        except Exception as ex:  # a diagnostic: never lose the leg over it
            index_build = {"error": repr(ex)}
    # parity: sampled probe K-mers by the POPC all-pairs kernel (targeted rules) and a NumPy brute force
    parity = None
    if c.rank == 0:
        rng = np.random.default_rng(6)
        pl = (L - 1) // 2
        starts = [0, pl - 300, pl + 1] + [int(v) for v in rng.integers(0, L - 300, size=13)]
        tmp = torch.empty(256, dtype=torch.int16, device=c.dev)
        bad = 0
        for b in starts:
            hamm.allpairs_min_device(q_img, t_img, both, False, b, b + 256, tmp.data_ptr(), clamp, c.stream.cuda_stream)
            bad += int((tmp != out[b:b + 256]).sum().item())
        host = out.cpu().numpy().view(np.uint16)
        tvalid = valid_mask(target, K)
        cand = [starts[0] + 5, starts[1] + 9, starts[2] + 17, starts[5] + 3]
        cand = [p for p in cand if (probes[p:p + K] < 4).all()]
        with concurrent.futures.ThreadPoolExecutor(max_workers=4) as ex:
            want = list(ex.map(lambda p: min(clamp, numpy_min_distance(target, tvalid, K, probes[p:p + K], both, None)), cand))
        bad_brute = sum(int(host[p] != w) for p, w in zip(cand, want))
        hist = {int(v): int((host == v).sum()) for v in range(0, clamp + 1)}
        parity = {"ok": bad == 0 and bad_brute == 0, "popc_kernel_probes": 256 * len(starts), "popc_kernel_mismatches": bad,
                  "numpy_brute_force_probes": len(cand), "numpy_brute_force_mismatches": bad_brute, "result_histogram": hist}
    barrier(c)
    e2e = None
    if e2e_steps and c.world == 1:
        h_t = torch.from_numpy(target).pin_memory().numpy()
        h_p = torch.from_numpy(probes).pin_memory().numpy()
        c.k4b.targeted(h_t, h_p, K, R, both)  # warm-up of the host path
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h = c.k4b.targeted(h_t, h_p, K, R, both)
        dt = time.perf_counter() - t0
        hv = h != 0xFF
        ok = int(h[hv].astype(np.int64).sum()) == int(out.cpu().numpy().view(np.uint16)[hv].astype(np.int64).sum())
        e2e = {"value": cmps * e2e_steps / dt / 1e9, "unit": "Gcmp/s", "h2d_bytes_per_step": int(len(target) + len(probes)),
               "d2h_bytes_per_step": int(2 * L), "steps": e2e_steps, "seconds_per_step": dt / e2e_steps,
               "result_checksum_equals_resident_run": ok, "api": "k4b_hamm_targeted (pinned host buffers)"}
    elif e2e_steps:
        h_t = torch.from_numpy(target).pin_memory().numpy() if c.rank == 0 else None
        h_p = torch.from_numpy(probes).pin_memory().numpy() if c.rank == 0 else None
        c.kdist.targeted_distributed(h_t, h_p, K, R, both, engine=c.engine)
        barrier(c)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h = c.kdist.targeted_distributed(h_t, h_p, K, R, both, engine=c.engine)
        barrier(c)
        dt = max_over_ranks(c, time.perf_counter() - t0)
        ok = None
        if c.rank == 0:
            hv = h != 0xFF
            ok = int(h[hv].astype(np.int64).sum()) == int(out.cpu().numpy().view(np.uint16)[hv].astype(np.int64).sum())
        e2e = {"value": cmps * e2e_steps / dt / 1e9, "unit": "Gcmp/s", "h2d_bytes_per_step": int(len(target) + len(probes)),
               "d2h_bytes_per_step": int(2 * L), "steps": e2e_steps, "seconds_per_step": dt / e2e_steps,
               "result_checksum_equals_resident_run": ok, "api": "kit4b_b200.dist.targeted_distributed (rank-0 pinned host buffers)"}
    cpu = None
    if with_cpu and c.rank == 0 and c.world == 1:
        try:
            cpu = cfg4_reference(scale)
        except Exception as exc:  # the reference arm must never take the GPU numbers down with it
            cpu = {"unavailable": str(exc)[:300]}
    q_img.free()
    t_img.free()
    return {"workload": "BASELINE configs[3]: hammings -m0 -K32 -r3 -c -I probes: %d probe K-mers (1 Mbp mutated copy + 1 Mbp random, "
                        "x%.2f) vs a %d-base synthetic assembly (20 entries)" % (Nq, scale, len(target)),
            "K": K, "R": R, "value": value, "unit": "Gcmp/s", "ms_per_step": ms_max / max(1, steps), "steps": steps, "warmup": warmup,
            "comparisons_per_step": cmps, "probe_kmers": int(Nq), "target_kmers": int(Nt),
            "engine": "seed-and-verify (pigeonhole cores of %d bases, bucket index with flank signatures, %s)"
                      % (core, "bucket-major join" if join else "warp per item"),
            "parallelism": "index-bucket shards x%d (each rank builds 1/%d of the index, answers all probes) + all_reduce(MIN)" % (c.world, c.world),
            "e2e": e2e, "roofline": roofline, "index_build": index_build, "result_checksum": checksum, "parity": parity,
            "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks}


# ------------------------------------------------------------------------------------------
# the POPC all-pairs kernel alone (executed word-compares vs the POPC peak)
# ------------------------------------------------------------------------------------------
def popc_leg(c, workload, steps, warmup, batch):
    """step = one query batch per GPU against all targets (per-query work is uniform); weak scaling"""
    torch, hamm = c.torch, c.hamm
    concat, chroms, K, both = synth_genome(workload)
    L = len(concat)
    S = 2 if both else 1
    W = (K + 31) // 32
    Nt = valid_count(chroms, K)
    B = min(batch, L)
    packed = broadcast_packed(c, concat, K)
    lo, hi = c.kdist.shard_bounds(0, L, c.world)[c.rank]

    def batch_range(i):
        span = max(1, hi - lo - B)
        b = lo + (i * B) % span
        return b, min(b + B, hi)

    out = torch.empty(B, dtype=torch.int16, device=c.dev)
    for i in range(warmup):
        c.flush.fill_(i & 0xFF)
        c.engine.compute(packed, both, *batch_range(i), out)
    barrier(c)
    sampler = ClockSampler(c.local).start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, nq_total, kernel_ms = 0, 0, []
    ev0.record(c.stream)
    for i in range(steps):
        c.flush.fill_(i & 0xFF)
        b, e = batch_range(warmup + i)
        launches += c.engine.compute(packed, both, b, e, out) + 1
        nq_total += valid_count(chroms, K, b, e)
        kernel_ms.append(hamm.last_kernel_ms())
    ev1.record(c.stream)
    barrier(c)
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    ms_max = max_over_ranks(c, ms)
    q = torch.tensor([float(nq_total)], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        c.dist.all_reduce(q, op=c.dist.ReduceOp.SUM)
    value = float(q.item()) * Nt * S / (ms_max * 1e-3) / 1e9
    k_ms = float(np.mean(kernel_ms))
    achieved = (nq_total / max(1, steps)) * Nt * S * W / (k_ms * 1e-3) / 1e9
    prof = _profile_note(workload + "_popc")
    roofline = {"bound": "int_pipe(xu: popc)", "achieved": achieved, "peak": c.popc_peak, "unit": "Gwc/s",
                "frac": achieved / c.popc_peak, "traffic": prof.get("dram_bytes_per_launch"), "profile": prof.get("source"),
                "kernel": "allpairs_min_kernel<W=%d>" % W, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * steps / ms if ms > 0 else None,
                "ops_model": "SURVEY 8d: executed 32-base word-compares (queries x targets x strands x ceil(K/32)), 1 POPC each",
                "peak_source": "measured live: register-resident POPC microbenchmark; nominal 16 lanes/clk/SM x 148 SMs x "
                               "1.965 GHz = 4654"}
    # parity: the last batch against the oracle-independent NumPy brute force on 4 queries
    parity = None
    if c.rank == 0:
        b, e = batch_range(warmup + steps - 1)
        host = out[:e - b].cpu().numpy().view(np.uint16)
        valid = valid_mask(concat, K)
        cand = [p for p in (b + 3, b + (e - b) // 2, e - 200) if valid[p]]
        bad = sum(int(host[p - b] != numpy_min_distance(concat, valid, K, concat[p:p + K], both, p)) for p in cand)
        parity = {"ok": bad == 0, "numpy_brute_force_queries": len(cand), "numpy_brute_force_mismatches": bad}
    packed.free()
    return {"workload": WORKLOAD_DESCR[workload] + " - POPC all-pairs kernel alone (the north-star XOR / fold / POPC formulation)",
            "K": K, "value": value, "unit": "Gcmp/s", "ms_per_step": ms_max / max(1, steps), "steps": steps, "warmup": warmup,
            "step": "query batch of %d K-mers per GPU vs all %d targets, both strands (a full pass is queries/batch steps)" % (B, Nt),
            "scaling": "weak", "roofline": roofline, "parity": parity, "gpu_launches": launches, "clocks": clocks}


# ------------------------------------------------------------------------------------------
# one process, N GPUs: the path `k4b_hammings --gpus N` uses (k4b_gpu_init(N) + host-buffer ABI)
# ------------------------------------------------------------------------------------------
def inproc_leg(c, n_gpus, concat, chroms, K, both, cmps, checksum, steps=3):
    from kit4b_b200 import hostlib
    hamm, k4b, torch = c.hamm, c.k4b, c.torch
    res = {"api": "k4b_gpu_init(%d) + k4b_hamm_exhaustive (host buffers; one process drives all GPUs: NCCL broadcast of the "
                  "packed set, pair-matrix partition, ncclAllReduce(min) per slab)" % n_gpus}
    try:
        k4b.gpu_shutdown()
        k4b.gpu_init(n_gpus)
        h_concat = torch.from_numpy(concat).pin_memory().numpy()
        hamm.exhaustive(h_concat, K, both)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = hamm.exhaustive(h_concat, K, both)
        dt = time.perf_counter() - t0
        res.update({"value": cmps * steps / dt / 1e9, "unit": "Gcmp/s", "steps": steps, "seconds_per_step": dt / steps,
                    "result_checksum_equals_resident_run": int(r.astype(np.int64).sum()) == checksum})
        k4b.gpu_shutdown()
        # the CLI, file to file (CUDA context creation, bioseq read, 200 MB CSV write included)
        with tempfile.TemporaryDirectory() as d:
            fa, seq, csv = os.path.join(d, "g.fa"), os.path.join(d, "g.seq"), os.path.join(d, "out.csv")
            write_fasta(fa, [(nm, concat[st:st + n]) for nm, st, n in chroms])
            hostlib.fasta_to_bioseq(fa, seq, "bench")
            cli = {}
            for g in sorted({1, n_gpus}):
                a = [hostlib.cli_path(), "hammings", "-m1", "-K%d" % K] + (["-c"] if both else []) + \
                    ["--gpus=%d" % g, "-i", seq, "-o", csv]
                t0 = time.perf_counter()
                p = subprocess.run(a, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
                cli["gpus_%d" % g] = {"wall_seconds": round(time.perf_counter() - t0, 2), "rc": p.returncode,
                                      "csv_bytes": os.path.getsize(csv) if os.path.exists(csv) else 0}
            res["cli_file_to_file"] = cli
        k4b.gpu_init(1, [c.local])
    except Exception as exc:
        res["error"] = str(exc)[:300]
    return res


# ------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------
def run_ours(args):
    c = make_ctx()
    want = [] if args.configs == "none" else (ALL_CONFIGS if args.configs == "all" else args.configs.split(","))
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else min(args.steps, 8)
    line = None
    if args.workload == "cfg4":
        leg = targeted_leg(c, args.scale, args.steps, args.warmup, e2e_steps, not args.no_cpu)
        if c.rank == 0:
            line = {"metric": "kmer_comparisons_per_sec", "value": leg["value"], "unit": "Gcmp/s", "n_gpus": c.world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                    "config": {"workload": leg["workload"], "K": leg["K"], "R": leg["R"], "engine": leg["engine"],
                               "parallelism": leg["parallelism"], "l2": "256 MB flush write between timed steps",
                               "result_checksum": leg["result_checksum"]},
                    "roofline": leg["roofline"], "index_build": leg["index_build"], "cpu_baseline": leg["cpu_baseline"],
                    "e2e": leg["e2e"], "parity": leg["parity"], "gpu_launches": leg["gpu_launches"], "clocks": leg["clocks"]}
    elif args.engine == "popc":
        leg = popc_leg(c, args.workload, args.steps, args.warmup, args.batch)
        if c.rank == 0:
            line = {"metric": "kmer_comparisons_per_sec", "value": leg["value"], "unit": "Gcmp/s", "n_gpus": c.world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                    "config": {"workload": leg["workload"], "K": leg["K"], "step": leg["step"],
                               "l2": "256 MB flush write between timed steps"},
                    "roofline": leg["roofline"], "cpu_baseline": None, "e2e": None, "parity": leg["parity"],
                    "gpu_launches": leg["gpu_launches"], "clocks": leg["clocks"]}
    else:
        head = bands_leg(c, args.workload, args.steps, args.warmup, e2e_steps, min(args.warmup, 3))
        cpu, file_parity = None, None
        if c.rank == 0 and c.world == 1 and not args.no_cpu:
            rs = ReferenceSampler(args.workload)
            try:
                r = rs.sample(args.cpu_seconds)
                cpu = {k: r[k] for k in ("value", "value_incl_fixed_costs", "unit", "cores", "kind", "sample")}
                file_parity = rs.file_parity(min(args.cpu_seconds, 8.0))
            except Exception as exc:
                cpu = {"unavailable": str(exc)[:300]}
            finally:
                rs.close()
        configs = []
        for name in want:
            if name == args.workload:
                continue
            t0 = time.perf_counter()
            try:
                if name == "cfg4":
                    ent = targeted_leg(c, args.scale, 5, 3, 3, not args.no_cpu)
                elif name == "popc":
                    ent = popc_leg(c, args.workload, 5, 3, args.batch)
                elif name == "cfg3":
                    # 5e15 comparisons: ~75 s per job on one B200 -> one warm-up and one timed job (two from 4 GPUs up)
                    ent = compact(bands_leg(c, name, 2 if c.world >= 4 else 1, 1, 1 if c.world >= 4 else 0, 0, n_brute=4))
                    ent["note"] = "1 warm-up job (the jobs take minutes at N=1); e2e only from 4 GPUs up"
                elif name in WORKLOADS:
                    ent = compact(bands_leg(c, name, 3, 3, 2, 1))
                else:
                    continue
                ent["leg_wall_seconds"] = round(time.perf_counter() - t0, 1)
            except Exception as exc:
                ent = {"workload": name, "error": str(exc)[:300]}
            configs.append(ent)
        if c.rank == 0:
            par = head["parity"] or {}
            if file_parity is not None:
                par["reference_file_diff"] = file_parity
                par["ok"] = bool(par.get("ok")) and (file_parity.get("identical", True) if file_parity.get("ran") else True)
            head["e2e"] = head["e2e"] or {"value": None, "unit": "Gcmp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
            line = {
                "metric": "kmer_comparisons_per_sec", "value": head["value"], "unit": "Gcmp/s", "n_gpus": c.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": head["descr"], "K": head["K"], "both_strands": head["both"],
                           "genome_bases": int(head["L"]), "kmers": int(head["Nv"]),
                           "step": "the whole all-vs-all job: every K-mer vs every K-mer, %s (%.3g comparisons)"
                                   % ("both strands" if head["both"] else "Watson only", head["cmps"]),
                           "engine": "diagonal bands (bit-sliced sliding counters) bootstrapped by the POPC all-pairs kernel",
                           "parallelism": "pair-matrix partition x%d + all_reduce(MIN) after the bootstrap and after every slab "
                                          "(overlapped with the next slab)" % c.world,
                           "l2": "256 MB flush write between timed steps", "result_checksum": head["checksum"]},
                "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": head["e2e"], "parity": par,
                "gpu_launches": head["launches"], "clocks": head["clocks"], "configs": configs,
            }
    # ---- N > 1: afterwards rank 0 alone drives all GPUs from one process ----
    inproc_args = None
    if c.world > 1 and args.engine != "popc" and args.workload != "cfg4" and not args.no_inproc:
        inproc_args = (c.world, head["concat"], head["chroms"], head["K"], head["both"], head["cmps"], head["checksum"])
    if c.world > 1:
        c.dist.barrier()
        c.dist.destroy_process_group()
    if c.rank != 0:
        c.k4b.gpu_shutdown()
        return
    if inproc_args is not None:
        time.sleep(2.0)  # the other ranks are exiting: let them release their contexts
        line["e2e_inproc"] = inproc_leg(c, *inproc_args)
    print(json.dumps(line), flush=True)
    c.k4b.gpu_shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg4"])
    ap.add_argument("--configs", default="all", help="extra BASELINE configs measured after the headline: all | none | "
                                                     "comma list of " + ",".join(ALL_CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="cfg4: size factor of assembly and probes")
    ap.add_argument("--engine", default="bands", choices=["bands", "popc"],
                    help="bands: diagonal-band engine, step = whole job (default); popc: all-pairs POPC kernel, step = query batch")
    ap.add_argument("--batch", type=int, default=131072, help="POPC leg: query K-mers per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the reference sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-inproc", action="store_true", help="skip the one-process multi-GPU leg (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args)


if __name__ == "__main__":
    main()
