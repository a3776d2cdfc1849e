"""Region mode of the HammingDist drop-in (k4b_hammingdist -I) against
golden files written by the UNMODIFIED reference tool (HammingDist/HammingDist.cpp through
oracle/_ref/hammingdist_ref; tests/golden/make_hammingdist_golden.py made them): BED12 gene models,
BED6, the binary biobed container of `genbiobed`, GFF3 gene models, -r / -R, chromosome aliases and the
reference's reading rules (descriptor row, unknown chromosome mid-file, a feature file it cannot parse)."""
import json
import os
import subprocess

import pytest

from kit4b_b200 import hostlib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hammingdist")
CASES = json.load(open(os.path.join(GOLDEN, "manifest.json")))
EXE = os.path.join(os.path.dirname(hostlib.cli_path()), "k4b_hammingdist")


def _args(case, feats):
    a = []
    for c in case["csvs"]:
        a += ["-i", os.path.join(GOLDEN, c)]
    a += ["-I", os.path.join(GOLDEN, feats)]
    if case["r"] is not None:
        a += ["-r", str(case["r"])]
    if case["R"] is not None:
        a += ["-R%d" % case["R"]]
    return a


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_region_distribution_equals_reference(case, tmp_path):
    want = open(os.path.join(GOLDEN, case["out"]), "rb").read()
    out = str(tmp_path / "dist.csv")
    p = subprocess.run([EXE] + _args(case, case["bed"]) + ["-o", out], capture_output=True)
    assert p.returncode == 0, p.stdout
    assert open(out, "rb").read() == want
    # same through the C ABI of the host library
    out2 = str(tmp_path / "dist2.csv")
    hostlib.hamming_dist_regions([os.path.join(GOLDEN, c) for c in case["csvs"]], os.path.join(GOLDEN, case["bed"]), out2,
                                 reg_len=case["r"] or 0, ofs_loci=case["R"] or 0)
    assert open(out2, "rb").read() == want
    if "biobed" in case:  # the preprocessed binary container holds the same features
        out3 = str(tmp_path / "dist3.csv")
        p = subprocess.run([EXE] + _args(case, case["biobed"]) + ["-o", out3], capture_output=True)
        assert p.returncode == 0, p.stdout
        assert open(out3, "rb").read() == want


def test_goldens_cover_the_reading_rules():
    """the fixtures really exercise what their names say"""
    by = {c["name"]: c for c in CASES}
    size = lambda n: os.path.getsize(os.path.join(GOLDEN, by[n]["out"]))
    assert size("descriptor_row_r100") == 0      # the G,b,B row names no chromosome: nothing is read
    assert size("commas_header_r100") == 0       # comma-separated BED is not parseable by the reference either
    assert size("genes_r0") > 0 and size("unknown_chrom_midfile_r100") > 0
    header = open(os.path.join(GOLDEN, by["genes_r0"]["out"])).readline()
    assert header.startswith(',"CDS","UTR5","UTR3","Intron","UP5","DN3","Intergenic","Proportion CDS"')


def test_feature_bits_by_hand(tmp_path):
    """overlap rules on a hand-made gene: exons 100-249, 400-599, 650-899 (+ strand), coding 200..800"""
    bed = tmp_path / "g.bed"
    bed.write_text("chrA\t100\t900\tgeneP\t0\t+\t200\t800\t0\t3\t150,200,250,\t0,300,550,\n"
                   "chrA\t1200\t2000\tgeneM\t0\t-\t1300\t1900\t0\t2\t300,400,\t0,400,\n")
    CDS, U5, U3, INTRON, UP, DN = 1, 2, 4, 8, 16, 32
    bits = hostlib.feature_bits(str(bed), "CHRa", [99, 100, 150, 249, 250, 399, 400, 899, 900, 1199, 1200, 1999, 2000], 0)
    exon = CDS | U5 | U3  # any exon overlap raises all three exon bits (BEDfile.h:43)
    assert bits == [0, exon, exon, exon, INTRON, INTRON, exon, exon, 0, 0, exon, exon, 0]
    # regulatory length 50: upstream of the '+' gene is 50..99, downstream 900..949; the '-' gene has its
    # downstream on the left, 1151..1199 (the reference's test there is exclusive, BEDfile.cpp:4062-4064), upstream 2000..2049
    loci = [0, 49, 50, 99, 900, 949, 950, 1000, 1149, 1150, 1151, 1199, 2000, 2049, 2050]
    assert hostlib.feature_bits(str(bed), "chrA", loci, 50) == [0, 0, UP, UP, DN, DN, 0, 0, 0, 0, DN, DN, UP, UP, 0]
    assert hostlib.feature_bits(str(bed), "chrZ", [5], 0) == [-1]


def test_gff3_gene_models_by_hand(tmp_path):
    """the reference's GFF3 reading rules (BEDfile.cpp:757-1131): 1-based coordinates kept as they are, exon blocks
    = painted runs, only the first two blocks of a gene survive its writer, a later mRNA line opens a new gene, an
    mRNA directly after its gene is ignored, a gene without Name= is dropped, one without exons is one block"""
    gff = tmp_path / "g.gff3"
    gff.write_text("##gff-version 3\n"
                   "chrA\tsrc\tgene\t101\t900\t.\t+\t.\tID=g1;Name=geneP\n"
                   "chrA\tsrc\tmRNA\t101\t900\t.\t+\t.\tID=m1;Parent=g1;Name=mP\n"
                   "chrA\tsrc\texon\t101\t250\t.\t+\t.\tParent=m1\n"
                   "chrA\tsrc\tCDS\t200\t250\t.\t+\t0\tParent=m1\n"
                   "chrA\tsrc\texon\t401\t600\t.\t+\t.\tParent=m1\n"
                   "chrA\tsrc\texon\t651\t900\t.\t+\t.\tParent=m1\n"
                   "chrA\tsrc\tmRNA\t1001\t1100\t.\t-\t.\tID=m2;Parent=g1;Name=isoform2\n"
                   "chrA\tsrc\texon\t1001\t1050\t.\t-\t.\tParent=m2\n"
                   "chrA\tsrc\tgene\t1301\t1400\t.\t+\t.\tID=g3\n"
                   "chrA\tsrc\texon\t1301\t1400\t.\t+\t.\tParent=g3\n"
                   "chrA\tsrc\tgene\t1501\t1600\t.\t.\t.\tID=g4;Name=lonely\n")
    exon, intron = 7, 8
    loci = [100, 101, 250, 251, 400, 401, 600, 601, 651, 900, 1000, 1001, 1050, 1051, 1100, 1101, 1301, 1400, 1500, 1501, 1599, 1600, 1601]
    want = [0, exon, exon, intron, intron, exon, exon, 0, 0, 0, 0, exon, exon, 0, 0, 0, 0, 0, 0, exon, exon, 0, 0]
    assert hostlib.feature_bits(str(gff), "chrA", loci, 0) == want


def test_region_mode_errors(tmp_path):
    csv = tmp_path / "h.csv"
    csv.write_text('"chrA",5,3\n"chrA",6,250\n')
    bed = tmp_path / "g.bed"
    bed.write_text("chrA\t100\t900\n")
    out = str(tmp_path / "o.csv")
    with pytest.raises(RuntimeError, match="does not fit"):       # distances above 200 overflow the reference's table
        hostlib.hamming_dist_regions([str(csv)], str(bed), out)
    bad = tmp_path / "bad.bed"
    bad.write_text("chrA\t100\t900\tg\t0\t+\t100\t900\t255,0,0\t1\t800,\t0,\n")  # itemRgb triple: reference cannot read it
    with pytest.raises(RuntimeError, match="malformed"):
        hostlib.hamming_dist_regions([str(csv)], str(bad), out)
    gff = tmp_path / "g.gff3"  # a feature outside its gene is fatal in the reference's GFF3 reader too
    gff.write_text("##gff-version 3\nchrA\tsrc\tgene\t100\t900\t.\t+\t.\tID=g1;Name=g\nchrA\tsrc\texon\t90\t200\t.\t+\t.\tParent=g1\n")
    with pytest.raises(RuntimeError, match="outside range of gene"):
        hostlib.hamming_dist_regions([str(csv)], str(gff), out)
    gff.write_text("##gff-version 3\nchrA\tsrc\tgene\t100\t900\t.\t+\t.\tID=g1\n")  # no Name=: the gene is dropped
    with pytest.raises(RuntimeError, match="Unable to load any features"):
        hostlib.hamming_dist_regions([str(csv)], str(gff), out)
    for flag in (["-r", "1000001"], ["-R", "201"], ["-m", "1"], ["-s", "3"]):
        p = subprocess.run([EXE, "-i", str(csv), "-I", str(bed), "-o", out] + flag, capture_output=True)
        assert p.returncode == 1, flag
