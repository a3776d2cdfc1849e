"""BASELINE configs[3] (targeted mode) against the UNMODIFIED reference on this host, at a scale
the reference finishes in minutes: a 20-entry synthetic assembly (default 20 x 2.5 Mbp) indexed
with the reference's own `index`, probes = mutated copy + fresh random sequence through the
reference's `genbioseq`, then
    reference:  ngskit4b_ref hammings -m0 -K32 -r3 -c -T<cores> -i asm.sfx -I probes.seq -o ref.csv
    this repo:  k4b_hammings  hammings -m0 -K32 -r3 -c          -i asm.sfx -I probes.seq -o ours.csv
timed end to end (process start to exit) and compared byte for byte.  Test/measurement
infrastructure: needs oracle/_ref (built from /root/reference by oracle/build_ref.sh).
usage: python tools/cfg4_reference.py [scale] [--self]   (scale 1.0 = 50 Mbp assembly, 200 kbp probes;
       --self drops -I: the K-mers of the assembly against the assembly itself)"""
import json, os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ngskit4b_ref")
CLI = os.path.join(ROOT, "kit4b_b200", "bin", "k4b_hammings")
scale = float(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] != "--self" else 1.0
SELF = "--self" in sys.argv  # -m0 without -I: the K-mers of the assembly against the assembly itself
K, R = 32, 3
L = "ACGT"


def write_fasta(path, entries):
    with open(path, "w") as f:
        for name, codes in entries:
            f.write(">%s\n" % name)
            s = np.frombuffer(L.encode(), dtype=np.uint8)[codes].tobytes().decode()
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + "\n")


def run(cmd, cwd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return time.perf_counter() - t0, p


def main():
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(41)
    nchr, clen = 20, int(2_500_000 * scale)
    chroms = [rng.integers(0, 4, size=clen, dtype=np.uint8) for _ in range(nchr)]
    rng = np.random.default_rng(42)
    pl = int(100_000 * scale)
    copy = chroms[3][12345:12345 + pl].copy()
    idx = rng.choice(pl, size=int(0.03 * pl), replace=False)
    copy[idx] = (copy[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
    fresh = rng.integers(0, 4, size=pl, dtype=np.uint8)
    res = {"assembly_bases": nchr * clen, "probe_bases": 2 * pl, "K": K, "R": R, "host_cores": cores}
    with tempfile.TemporaryDirectory() as d:
        write_fasta(os.path.join(d, "asm.fa"), [("chr%d" % (i + 1), c) for i, c in enumerate(chroms)])
        write_fasta(os.path.join(d, "probes.fa"), [("mutated_copy", copy), ("unrelated", fresh)])
        t, p = run([REF, "index", "-i", "asm.fa", "-o", "asm.sfx", "-r", "asm", "-T%d" % min(cores, 64)], d)
        assert p.returncode == 0, p.stdout[-2000:]
        res["reference_index_s"] = round(t, 1)
        t, p = run([REF, "genbioseq", "-i", "probes.fa", "-o", "probes.seq", "-r", "probes"], d)
        assert p.returncode == 0, p.stdout[-2000:]
        args = ["hammings", "-m0", "-K%d" % K, "-r%d" % R, "-c", "-i", "asm.sfx"] + ([] if SELF else ["-I", "probes.seq"])
        res["mode"] = "-m0 without -I (assembly vs itself)" if SELF else "-m0 -I probes"
        t, p = run([REF] + args + ["-T%d" % min(cores, 64), "-o", "ref.csv"], d)
        assert p.returncode == 0, p.stdout[-2000:]
        res["reference_hammings_s"] = round(t, 2)
        res["reference_note"] = "wall time incl. the reference's fixed 12 s of sleeps (hammings.cpp:2433, SfxArray.cpp:1161)"
        for rep in range(2):  # second run: page cache and driver warm
            t, p = run([CLI] + args + ["-o", "ours.csv"], d)
            assert p.returncode == 0, p.stdout[-2000:]
            res["ours_hammings_s" if rep else "ours_hammings_first_run_s"] = round(t, 2)
            if os.environ.get("K4B_SHOW_LOG"):
                sys.stderr.write(p.stdout)
        ours, ref = open(os.path.join(d, "ours.csv"), "rb").read(), open(os.path.join(d, "ref.csv"), "rb").read()
        res["output_bytes"] = len(ref)
        res["outputs_identical"] = ours == ref
        nq = nchr * (clen - K + 1) if SELF else 2 * (pl - K + 1)
        nt = nchr * (clen - K + 1)
        res["logical_Gcmp"] = round(nq * nt * 2 / 1e9, 1)
        res["speedup_wall"] = round(res["reference_hammings_s"] / res["ours_hammings_s"], 1)
        res["speedup_excluding_reference_sleeps"] = round(max(0.0, res["reference_hammings_s"] - 12.0) / res["ours_hammings_s"], 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
