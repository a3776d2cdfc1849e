#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/ngskit4b_ref_nosleep, built from /root/reference by oracle/build_ref.sh) on small
seeded synthetic inputs.  Re-run only in the build container (the reference tree is not on the
GPU box).  Outputs: <case>.fa (input FASTA), <case>.seq (bioseq written by the reference's
genbioseq), <case>.<variant>.csv (hammings output) and manifest.json.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ngskit4b_ref_nosleep")


def fasta(entries, width=80):
    out = []
    for name, seq in entries:
        out.append(">%s synthetic\n" % name)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + "\n")
    return "".join(out)


def rnd(rng, n, alphabet="ACGT"):
    return "".join(rng.choices(alphabet, k=n))


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


def mutate(rng, s, n):
    s = list(s)
    for p in rng.sample(range(len(s)), n):
        s[p] = rng.choice([c for c in "ACGT" if c != s[p].upper()])
    return "".join(s)


def make_cases():
    cases = {}
    rng = random.Random(5)
    cases["tiny2"] = dict(entries=[("chrA", rnd(rng, 3000)), ("chrB", rnd(rng, 2000))],
                          runs=[dict(K=25, both=True), dict(K=25, both=False), dict(K=12, both=True)])
    # adversarial: palindromic K-mer, N runs, soft-masked (lower-case) bases, '-', a duplicate
    # between chromosomes, a reverse-complement copy, and a chromosome shorter than K
    rng = random.Random(7)
    a = rnd(rng, 400)
    pal = "ACGTTGCATGCAACGT"  # its own reverse complement (16-mer)
    dup = rnd(rng, 40)
    c1 = a[:100] + pal + a[100:200] + "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN" + a[200:300].lower() + dup + a[300:] + "N" + rnd(rng, 30)
    c2 = rnd(rng, 150) + dup + rnd(rng, 60) + "AC-GT" + rnd(rng, 80) + revcomp(a[20:70]) + rnd(rng, 40)
    c3 = "ACGTACG"  # shorter than K
    c4 = rnd(rng, 120) + "RYKM" + rnd(rng, 50) + mutate(rng, a[120:180], 2) + rnd(rng, 33)
    cases["adversarial"] = dict(entries=[("c1", c1), ("c2", c2), ("tinyc3", c3), ("c4", c4)],
                                runs=[dict(K=12, both=True), dict(K=12, both=False), dict(K=16, both=True),
                                      dict(K=33, both=True)])
    # multi-word K: planted near-copies so that small minima exist at every K
    rng = random.Random(11)
    base = rnd(rng, 1800)
    e1 = base + rnd(rng, 200)
    e2 = rnd(rng, 300) + mutate(rng, base[200:900], 25) + rnd(rng, 200) + revcomp(mutate(rng, base[1000:1500], 12))
    e3 = rnd(rng, 700)
    cases["multiword"] = dict(entries=[("m1", e1), ("m2", e2), ("m3", e3)],
                              runs=[dict(K=32, both=True), dict(K=50, both=True), dict(K=64, both=True),
                                    dict(K=65, both=False), dict(K=96, both=True), dict(K=100, both=True),
                                    dict(K=128, both=True), dict(K=150, both=True), dict(K=300, both=False)])
    # non-ACGT symbols at multi-word K (three-plane path)
    rng = random.Random(13)
    b2 = rnd(rng, 900)
    n1 = b2[:300] + "NNNN" + b2[300:600] + "-" + b2[600:] + rnd(rng, 100)
    n2 = rnd(rng, 100) + mutate(rng, b2[250:700], 9) + "N" + rnd(rng, 300)
    cases["nonacgt"] = dict(entries=[("n1", n1), ("n2", n2)],
                            runs=[dict(K=20, both=True), dict(K=40, both=True), dict(K=70, both=True),
                                  dict(K=100, both=False), dict(K=140, both=True)])
    return cases


def make_targeted_cases():
    """-m0 -I (targeted) cases: a small indexed assembly and probe sets built from mutated copies of
    it (forward and reverse complement) plus random sequence; one probe set carries N / '-' symbols."""
    rng = random.Random(41)
    t1, t2 = rnd(rng, 12000), rnd(rng, 8000)
    target = [("tA", t1), ("tB", t2)]
    rng = random.Random(42)

    def probe_set(with_n):
        parts = []
        for src, start, n, nmut, rc in [(t1, 500, 300, 0, False), (t1, 3000, 300, 4, False), (t2, 1000, 300, 9, True),
                                        (t1, 7000, 300, 14, False), (t2, 5000, 300, 20, True), (t1, 9000, 200, 30, False)]:
            seg = mutate(rng, src[start:start + n], nmut)
            parts.append(revcomp(seg) if rc else seg)
        parts.append(rnd(rng, 600))
        p1 = "".join(parts)
        p2 = mutate(rng, t2[6000:6400], 10) + rnd(rng, 150)
        if with_n:
            p1 = p1[:310] + "N" + p1[311:650] + "NN" + p1[652:900] + "NNNNN" + p1[905:1200] + "-" + p1[1201:]
            p2 = p2[:40] + "R" + p2[41:]
        return [("pX", p1), ("pshort", "ACGTACGTAC"), ("pY", p2)]

    return dict(target=target,
                probes={"tp": probe_set(False), "tpn": probe_set(True)},
                runs=[dict(probes="tp", K=32, R=3, both=True, fmt=0), dict(probes="tp", K=32, R=3, both=True, fmt=1),
                      dict(probes="tp", K=32, R=3, both=True, fmt=2), dict(probes="tp", K=25, R=3, both=False, fmt=0),
                      dict(probes="tp", K=20, R=1, both=True, fmt=0), dict(probes="tp", K=50, R=5, both=True, fmt=0),
                      dict(probes="tpn", K=32, R=3, both=True, fmt=0), dict(probes="tpn", K=25, R=2, both=True, fmt=2)])


def main_targeted(manifest):
    case = make_targeted_cases()
    tfa = os.path.join(HERE, "targ.fa")
    sfx = os.path.join(HERE, "targ.sfx")
    open(tfa, "w").write(fasta(case["target"]))
    subprocess.run([REF.replace("_nosleep", ""), "index", "-i", tfa, "-o", sfx, "-r", "targ", "-T2"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    probes = {}
    for pname, entries in case["probes"].items():
        fa = os.path.join(HERE, pname + ".fa")
        seq = os.path.join(HERE, pname + ".seq")
        open(fa, "w").write(fasta(entries))
        subprocess.run([REF, "genbioseq", "-i", fa, "-o", seq, "-r", pname], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        probes[pname] = dict(fasta=os.path.basename(fa), bioseq=os.path.basename(seq))
    runs = []
    for r in case["runs"]:
        ext = ("csv", "bed", "wig")[r["fmt"]]
        tag = "%s.K%dr%d%s%s" % (r["probes"], r["K"], r["R"], "c" if r["both"] else "w",
                                 ("k%d" % r["sample"]) if r.get("sample") else "")
        out = os.path.join(HERE, "targeted.%s.%s" % (tag, ext))
        # the Wiggle header embeds the -o path: run with a relative name from inside tests/golden
        args = [REF.replace("_nosleep", ""), "hammings", "-m0", "-K%d" % r["K"], "-r%d" % r["R"], "-S%d" % r["fmt"],
                "-T2", "-i", "targ.sfx", "-I", probes[r["probes"]]["bioseq"], "-o", os.path.basename(out)]
        if r["both"]:
            args.insert(3, "-c")
        if r.get("sample"):
            args.insert(3, "-k%d" % r["sample"])
        subprocess.run(args, check=True, cwd=HERE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        runs.append(dict(r, out=os.path.basename(out)))
    manifest["__targeted__"] = dict(target_fasta="targ.fa", sfx="targ.sfx", probes=probes, runs=runs)


def main_targeted_self(manifest, only_new=False):
    """-m0 without -I: K-mers of the indexed assembly against the assembly itself."""
    rng = random.Random(61)
    a = rnd(rng, 2500)
    pal = "ACGTTGCATGCAACGTACGTTGCATGCAACGT"  # 32-mer equal to its own reverse complement
    s1 = a[:800] + pal + a[800:1500] + "NNNNNNN" + a[1500:2000] + mutate(rng, a[100:400], 3) + a[2000:] + "N" + rnd(rng, 40)
    s2 = rnd(rng, 300) + a[1000:1200] + rnd(rng, 100) + revcomp(mutate(rng, a[300:700], 5)) + rnd(rng, 200) + "AC-GT" + rnd(rng, 120)
    s3 = mutate(rng, a[1700:1900], 1) + rnd(rng, 60)
    tfa = os.path.join(HERE, "targs.fa")
    sfx = os.path.join(HERE, "targs.sfx")
    if not (only_new and os.path.exists(sfx)):
        open(tfa, "w").write(fasta([("sA", s1), ("sB", s2), ("sC", s3)]))
        subprocess.run([REF.replace("_nosleep", ""), "index", "-i", tfa, "-o", sfx, "-r", "targs", "-T2"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    runs = []
    # z = -z intra/inter filter of exact sense hits (1 intra only, 2 inter only; SfxArray.cpp:4421-4426)
    for K, R, both, fmt, z in [(32, 3, True, 0, 0), (25, 2, False, 0, 0), (20, 1, True, 1, 0), (50, 5, True, 2, 0),
                               (25, 2, True, 0, 1), (25, 2, True, 0, 2), (20, 1, False, 0, 1), (32, 3, False, 0, 2),
                               (140, 9, True, 0, 2)]:
        ext = ("csv", "bed", "wig")[fmt]
        out = "targself.K%dr%d%s%s.%s" % (K, R, "c" if both else "w", ("z%d" % z) if z else "", ext)
        args = [REF.replace("_nosleep", ""), "hammings", "-m0", "-K%d" % K, "-r%d" % R, "-S%d" % fmt, "-T2",
                "-i", "targs.sfx", "-o", out]
        if both:
            args.insert(3, "-c")
        if z:
            args.insert(3, "-z%d" % z)
        if not (only_new and os.path.exists(os.path.join(HERE, out))):
            subprocess.run(args, check=True, cwd=HERE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        runs.append(dict(K=K, R=R, both=both, fmt=fmt, z=z, out=out))
    manifest["__targeted_self__"] = dict(target_fasta="targs.fa", sfx="targs.sfx", runs=runs)


def main_sweeps(manifest):
    """-m1 with -b/-B sub-ranges and -m2 node slices on the adversarial genome."""
    seq = "adversarial.seq"
    runs = []
    for K, both, b, B in [(12, True, 1, 50), (12, True, 37, 400), (12, False, 5, 5), (16, True, 300, 0),
                          (12, True, 1000, 1300), (12, True, 1, 1), (33, True, 2, 900)]:
        out = "sweep.K%d%s.b%dB%d.csv" % (K, "c" if both else "w", b, B)
        args = [REF, "hammings", "-m1", "-K%d" % K, "-T3", "-b%d" % b, "-B%d" % B, "-i", seq, "-o", out]
        if both:
            args.insert(3, "-c")
        subprocess.run(args, check=True, cwd=HERE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        runs.append(dict(mode=1, K=K, both=both, b=b, B=B, csv=out))
    for K, both, n, N in [(12, True, 3, 1), (12, True, 3, 2), (12, True, 3, 3), (12, False, 4, 2), (16, True, 2, 1)]:
        out = "slice.K%d%s.n%dN%d.csv" % (K, "c" if both else "w", n, N)
        args = [REF, "hammings", "-m2", "-K%d" % K, "-T3", "-n%d" % n, "-N%d" % N, "-i", seq, "-o", out]
        if both:
            args.insert(3, "-c")
        subprocess.run(args, check=True, cwd=HERE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        runs.append(dict(mode=2, K=K, both=both, n=n, N=N, csv=out))
    manifest["__sweeps__"] = dict(bioseq=seq, runs=runs)


def main():
    if not os.access(REF, os.X_OK):
        sys.exit("reference binary missing: run oracle/build_ref.sh first")
    if len(sys.argv) > 1 and sys.argv[1] == "--targeted-self-only":
        # add new -m0 (no -I) runs without regenerating the other fixtures
        mpath = os.path.join(HERE, "manifest.json")
        manifest = json.load(open(mpath))
        main_targeted_self(manifest, only_new=True)
        json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)
        return
    manifest = {}
    main_targeted(manifest)
    main_targeted_self(manifest)
    main_sweeps(manifest)
    for name, case in make_cases().items():
        fa = os.path.join(HERE, name + ".fa")
        seq = os.path.join(HERE, name + ".seq")
        open(fa, "w").write(fasta(case["entries"]))
        subprocess.run([REF, "genbioseq", "-i", fa, "-o", seq, "-r", name], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        runs = []
        for r in case["runs"]:
            tag = "K%d%s" % (r["K"], "c" if r["both"] else "w")
            out = os.path.join(HERE, "%s.%s.csv" % (name, tag))
            args = [REF, "hammings", "-m1", "-K%d" % r["K"], "-T3", "-i", seq, "-o", out]
            if r["both"]:
                args.insert(3, "-c")
            subprocess.run(args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            runs.append(dict(K=r["K"], both=r["both"], csv=os.path.basename(out)))
        manifest[name] = dict(fasta=os.path.basename(fa), bioseq=os.path.basename(seq), runs=runs)
    # -m3 merge: Watson-only file merged INTO a copy of ... both-strand file, and a copy-merge
    import shutil
    into = os.path.join(HERE, "merge.tiny2.into.csv")
    shutil.copyfile(os.path.join(HERE, "tiny2.K25w.csv"), into)
    subprocess.run([REF, "hammings", "-m3", "-i", os.path.join(HERE, "tiny2.K25c.csv"), "-o", into], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    fresh = os.path.join(HERE, "merge.tiny2.copy.csv")
    if os.path.exists(fresh):
        os.remove(fresh)
    subprocess.run([REF, "hammings", "-m3", "-i", os.path.join(HERE, "tiny2.K25c.csv"), "-o", fresh], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    manifest["__merge__"] = {"from": "tiny2.K25c.csv", "into_before": "tiny2.K25w.csv",
                             "into_after": "merge.tiny2.into.csv", "copy_after": "merge.tiny2.copy.csv"}
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(manifest), "cases")


if __name__ == "__main__":
    main()
