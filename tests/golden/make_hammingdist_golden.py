#!/usr/bin/env python
"""Golden fixtures for the region mode of the HammingDist drop-in (tests/golden/hammingdist/).

Runs the UNMODIFIED reference tool (oracle/_ref/hammingdist_ref = HammingDist/HammingDist.cpp, and
`genbiobed` of oracle/_ref/ngskit4b_ref for the binary feature container; both built from
/root/reference by oracle/build_ref.sh) on small seeded inputs.  Re-run only in the build container.

    python tests/golden/make_hammingdist_golden.py            # writes the fixtures + manifest
    python tests/golden/make_hammingdist_golden.py --fuzz 200 # reference vs drop-in on random cases (no files kept)
    python tests/golden/make_hammingdist_golden.py --fuzz-gff 200 # same with GFF3 gene models as the feature file
"""
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_HD = os.path.join(ROOT, "oracle", "_ref", "hammingdist_ref")
REF_NGS = os.path.join(ROOT, "oracle", "_ref", "ngskit4b_ref")
OURS = os.path.join(ROOT, "kit4b_b200", "bin", "k4b_hammingdist")
OUT = os.path.join(HERE, "hammingdist")


def gene_line(rng, chrom, start, length, name, sep="\t", detail=True):
    """one BED12 (or BED6) line; exons are sorted, non-overlapping, first starts at 0, last ends at length"""
    strand = rng.choice("+-+-?.")
    fields = [chrom, str(start), str(start + length), name, str(rng.randint(0, 900)), strand]
    if detail:
        n = rng.randint(1, 6)
        cuts = sorted(rng.sample(range(1, length), min(2 * n - 2, length - 1)))
        if len(cuts) % 2:
            cuts = cuts[:-1]
        bounds = [0] + cuts + [length]
        exons = [(bounds[i], bounds[i + 1]) for i in range(0, len(bounds), 2)]
        t0 = rng.randint(0, length - 1)
        t1 = rng.randint(t0, length)
        fields += [str(start + t0), str(start + t1), "0", str(len(exons)),
                   ",".join(str(e - s) for s, e in exons) + ",", ",".join(str(s) for s, e in exons) + ","]
    return sep.join(fields) + "\n"


def make_case(rng, n_genes=(3, 40), detail=True, sep="\t", header=False, chroms=("chrA", "chrB", "ChrC"),
              chrom_len=6000, csv_names=None, descriptor=False, unknown_at=None, n_files=1):
    bed = []
    if header:
        bed.append("track name=genes description=\"synthetic\"\n")
        bed.append("# a comment\n")
    k = 0
    for c in chroms:
        for _ in range(rng.randint(*n_genes)):
            length = rng.randint(2, 900)
            start = rng.randint(0, chrom_len - length)
            k += 1
            bed.append(gene_line(rng, c, start, length, "g%d" % k, sep, detail))
    names = csv_names or chroms
    csvs = []
    for f in range(n_files):
        rows = []
        if descriptor:
            rows.append("%d,2,%d\n" % (chrom_len * len(chroms), chrom_len * len(chroms)))
        r = 0
        for c in names:
            for loci in range(0, chrom_len, rng.choice([1, 3, 7])):
                r += 1
                if unknown_at is not None and f == n_files - 1 and r == unknown_at:
                    rows.append('"nosuchchrom",%d,%d\n' % (loci, 3))
                rows.append('"%s",%d,%d\n' % (c, loci, min(200, int(rng.expovariate(0.12)))))
        csvs.append("".join(rows))
    return "".join(bed), csvs


def make_gff_case(rng, chroms=("chrA", "chrB"), chrom_len=3000, n_genes=(2, 12), messy=0.0, csv_step=None):
    """GFF3 gene models: gene [mRNA] exon/CDS/UTR lines, second isoforms (a later mRNA line opens a new gene in the
    reference), genes without Name= (dropped), intron / other feature types, comments; `messy` adds the odd lines:
    swapped coordinates, attribute keys inside other keys and values, white space in values, 8-column lines"""
    out = ["##gff-version 3\n"]
    k = 0
    for c in chroms:
        out.append("##sequence-region %s 1 %d\n" % (c, chrom_len))
        if rng.random() < 0.5:
            out.append("%s\tsrc\tregion\t1\t%d\t.\t+\t.\tID=%s;Name=%s\n" % (c, chrom_len, c, c))
        for _ in range(rng.randint(*n_genes)):
            k += 1
            length = rng.randint(2, 700)
            start = rng.randint(1, chrom_len - length)
            end = start + length - 1
            strand = rng.choice("+-+-.?")
            score = rng.choice([".", ".", "0", "12.7", "1500", "-3"])
            name = "" if rng.random() < 0.12 else ";Name=gene%d" % k
            if messy and rng.random() < messy:
                name = rng.choice([";Alias=x%d;geneName=inner%d" % (k, k), ";Note=see Name=note%d here;x=1" % k,
                                   ";Name=spaced name %d" % k, ";Dbxref=ID=5;Name=late%d" % k])
            kind = rng.choice(["gene", "gene", "gene", "mRNA", "Gene"])
            out.append("%s\tsrc\t%s\t%d\t%d\t%s\t%s\t.\tID=g%d%s\n" % (c, kind, start, end, score, strand, k, name))
            for iso in range(rng.choice([1, 1, 1, 2, 3])):
                if kind != "mRNA" or iso:
                    if rng.random() < 0.8:
                        out.append("%s\tsrc\tmRNA\t%d\t%d\t.\t%s\t.\tID=g%d.t%d;Parent=g%d;Name=tr%d_%d\n"
                                   % (c, start, end, strand, k, iso, k, k, iso))
                n_ex = rng.randint(0, 5)
                cuts = sorted(rng.sample(range(start, end + 1), min(2 * n_ex, end - start + 1)))
                if len(cuts) % 2:
                    cuts = cuts[:-1]
                for i in range(0, len(cuts), 2):
                    a, b = cuts[i], cuts[i + 1]
                    ftype = rng.choice(["exon", "exon", "exon", "CDS", "five_prime_UTR", "three_prime_UTR", "EXON"])
                    if messy and rng.random() < messy * 0.3:
                        a, b = b, a  # swapped: accepted only when it still lies inside the gene
                    out.append("%s\tsrc\t%s\t%d\t%d\t.\t%s\t%s\tParent=g%d.t%d\n" % (c, ftype, a, b, strand, rng.choice(".012"), k, iso))
                    if ftype.lower() == "exon" and rng.random() < 0.4 and b - a > 4:
                        out.append("%s\tsrc\tCDS\t%d\t%d\t.\t%s\t0\tParent=g%d.t%d\n" % (c, min(a, b) + 1, max(a, b) - 1, strand, k, iso))
                if rng.random() < 0.2:
                    out.append("%s\tsrc\tintron\t%d\t%d\t.\t%s\t.\tParent=g%d\n" % (c, start, end, strand, k))
                if rng.random() < 0.15:
                    out.append("# a comment\n")
                if messy and rng.random() < messy * 0.2:
                    out.append("%s\tsrc\texon\t%d\t%d\t.\t%s\t.\n" % (c, start, min(end, start + 3), strand))  # no 9th column
    rows = []
    step = csv_step or rng.choice([1, 2, 5])
    for c in chroms:
        for loci in range(0, chrom_len + 50, step):
            rows.append('"%s",%d,%d\n' % (c, loci, min(200, int(rng.expovariate(0.12)))))
    return "".join(out), ["".join(rows)]


def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def reference_output(work, bed_path, csv_paths, r, R, out_name="ref.csv"):
    out = os.path.join(work, out_name)
    cmd = [REF_HD]
    for p in csv_paths:
        cmd += ["-i", p]
    cmd += ["-I", bed_path, "-o", out]
    if r is not None:
        cmd += ["-r", str(r)]
    if R is not None:
        cmd += ["-R", str(R)]
    p = run(cmd)
    return p.returncode, (open(out, "rb").read() if os.path.exists(out) else None)


def our_output(work, bed_path, csv_paths, r, R):
    out = os.path.join(work, "ours.csv")
    cmd = [OURS]
    for p in csv_paths:
        cmd += ["-i", p]
    cmd += ["-I", bed_path, "-o", out]
    if r is not None:
        cmd += ["-r", str(r)]
    if R is not None:
        cmd += ["-R", str(R)]
    p = run(cmd)
    return p.returncode, (open(out, "rb").read() if os.path.exists(out) else None)


CASES = [
    # name, make_case kwargs, -r, -R
    ("genes_r0", dict(), None, None),
    ("genes_r500", dict(), 500, None),
    ("genes_r2000_Rm150", dict(n_genes=(25, 60)), 2000, -150),
    ("genes_r50_R200", dict(n_genes=(25, 60)), 50, 200),
    ("bed6_r300", dict(detail=False), 300, None),
    ("bed6_r0", dict(detail=False), None, None),
    ("commas_header_r100", dict(sep=",", header=True), 100, None),
    ("alias_chloroplast_r100", dict(csv_names=("chrA", "chloroplast", "chrB")), 100, None),
    ("descriptor_row_r100", dict(descriptor=True), 100, None),
    ("unknown_chrom_midfile_r100", dict(unknown_at=900, n_files=2), 100, None),
    ("two_files_r1000", dict(n_files=2, n_genes=(25, 60)), 1000, 7),
]


GFF_CASES = [
    ("gff3_genes_r0", dict(), None, None),
    ("gff3_genes_r300_R5", dict(n_genes=(8, 20)), 300, 5),
    ("gff3_messy_r100", dict(messy=0.5, n_genes=(8, 20)), 100, None),
]


def write_fixtures():
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for idx, (name, kw, r, R) in enumerate(CASES + GFF_CASES):
        gff = name.startswith("gff3")
        for attempt in range(50):  # GFF3: the first seed whose file the reference reads to the end (many random ones
            rng = random.Random(1000 + idx + 100 * attempt)  # hold a feature outside its gene, which is fatal)
            bed, csvs = make_gff_case(rng, chrom_len=1500, csv_step=1, **kw) if gff else make_case(rng, chrom_len=1200, **kw)
            if not gff:
                break
            with tempfile.TemporaryDirectory() as work:
                open(os.path.join(work, "f.gff3"), "w").write(bed)
                open(os.path.join(work, "in.csv"), "w").write(csvs[0])
                rc, ref = reference_output(work, os.path.join(work, "f.gff3"), [os.path.join(work, "in.csv")], r, R)
            if rc == 0 and ref:
                break
        bed_path = os.path.join(OUT, name + (".gff3" if gff else ".bed"))
        open(bed_path, "w").write(bed)
        csv_paths = []
        for k, text in enumerate(csvs):
            p = os.path.join(OUT, "%s.in%d.csv" % (name, k))
            open(p, "w").write(text)
            csv_paths.append(p)
        rc, ref = reference_output(OUT, bed_path, csv_paths, r, R, name + ".dist.csv")
        assert rc == 0 and ref is not None, name
        entry = dict(name=name, bed=os.path.basename(bed_path), csvs=[os.path.basename(p) for p in csv_paths], r=r, R=R,
                     out=name + ".dist.csv")
        if name == "genes_r500":  # the same features as the binary container of `genbiobed`
            bio = os.path.join(OUT, name + ".biobed")
            p = run([REF_NGS, "genbiobed", "-i", bed_path, "-o", bio, "-d", "golden", "-t", "golden", "-b", "1"])
            assert p.returncode == 0, p.stdout
            rc2, ref2 = reference_output(OUT, bio, csv_paths, r, R, "tmp.csv")
            os.remove(os.path.join(OUT, "tmp.csv"))
            assert ref2 == ref
            entry["biobed"] = name + ".biobed"
        manifest.append(entry)
        print(name, len(ref), "bytes")
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)


def fuzz_gff(n):
    bad = 0
    for i in range(n):
        rng = random.Random(4242 + i)
        kw = dict(n_genes=rng.choice([(1, 4), (2, 12), (10, 30)]), messy=rng.choice([0.0, 0.0, 0.3, 0.8]),
                  chrom_len=rng.choice([800, 3000]), chroms=rng.choice([("chrA",), ("chrA", "chrB"), ("chrA", "ChrC", "chrB")]))
        r = rng.choice([None, 0, 1, 10, 100, 2000])
        R = rng.choice([None, 0, -200, -3, 5, 200])
        with tempfile.TemporaryDirectory() as work:
            gff, csvs = make_gff_case(rng, **kw)
            gff_path = os.path.join(work, "f.gff3")
            open(gff_path, "w").write(gff)
            csv_path = os.path.join(work, "in0.csv")
            open(csv_path, "w").write(csvs[0])
            a = reference_output(work, gff_path, [csv_path], r, R)
            b = our_output(work, gff_path, [csv_path], r, R)
            if a != b:
                bad += 1
                print("MISMATCH gff case", i, kw, r, R, a[0], b[0])
                if os.environ.get("KEEP"):
                    open("/tmp/gff_mismatch_%d.gff3" % i, "w").write(gff)
                    open("/tmp/gff_mismatch_%d.csv" % i, "w").write(csvs[0])
    print("fuzz gff: %d cases, %d mismatches" % (n, bad))
    return bad


def fuzz(n):
    bad = 0
    for i in range(n):
        rng = random.Random(777 + i)
        kw = dict(n_genes=rng.choice([(1, 5), (3, 40), (25, 80)]), detail=rng.random() < 0.8,
                  sep=rng.choice(["\t", "\t", ","]), header=rng.random() < 0.3,
                  descriptor=rng.random() < 0.1, n_files=rng.choice([1, 1, 2]),
                  unknown_at=rng.choice([None, None, None, 50, 2000]), chrom_len=rng.choice([1500, 6000]))
        r = rng.choice([None, 0, 1, 10, 100, 2000, 1000000])
        R = rng.choice([None, 0, -200, -3, 5, 200])
        with tempfile.TemporaryDirectory() as work:
            bed, csvs = make_case(rng, **kw)
            bed_path = os.path.join(work, "f.bed")
            open(bed_path, "w").write(bed)
            csv_paths = []
            for k, text in enumerate(csvs):
                p = os.path.join(work, "in%d.csv" % k)
                open(p, "w").write(text)
                csv_paths.append(p)
            a = reference_output(work, bed_path, csv_paths, r, R)
            b = our_output(work, bed_path, csv_paths, r, R)
            if a != b:
                bad += 1
                print("MISMATCH case", i, kw, r, R, a[0], b[0])
    print("fuzz: %d cases, %d mismatches" % (n, bad))
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--fuzz":
        sys.exit(1 if fuzz(int(sys.argv[2])) else 0)
    if len(sys.argv) > 2 and sys.argv[1] == "--fuzz-gff":
        sys.exit(1 if fuzz_gff(int(sys.argv[2])) else 0)
    write_fixtures()
