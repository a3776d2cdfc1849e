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


def main():
    if not os.access(REF, os.X_OK):
        sys.exit("reference binary missing: run oracle/build_ref.sh first")
    manifest = {}
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
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(manifest), "cases")


if __name__ == "__main__":
    main()
