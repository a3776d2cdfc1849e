"""CPU tests of the C++ host front end (containers, LoadGenome layout, byte-exact writers, CLI
parsing) through libk4bhost.so.  Arrays come from the pinned oracle; expected bytes are the
files written by the unmodified reference (tests/golden)."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, case_id, golden_cases, golden_manifest, load_case

from kit4b_b200 import hostlib


def targeted_runs():
    m = golden_manifest()["__targeted__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("name", [n for n in golden_manifest() if not n.startswith("__")])
def test_bioseq_reader_and_genome_layout(oracle, name):
    case = golden_manifest()[name]
    want, chroms, glen = load_case(oracle, name)
    got, gch, gl = hostlib.concat_from_bioseq(os.path.join(GOLDEN, case["bioseq"]), 25)
    assert gl == glen and np.array_equal(got, want)
    assert gch == [(n, s, l) for (n, s, l) in chroms]


@pytest.mark.parametrize("case", golden_cases(), ids=case_id)
def test_exhaustive_csv_writer_is_byte_exact(oracle, case, tmp_path):
    name, r = case
    concat, _, _ = load_case(oracle, name)
    hd = oracle.exhaustive_sliding(concat, r["K"], r["both"], threads=3)
    out = str(tmp_path / "o.csv")
    hostlib.write_exhaustive_csv(os.path.join(GOLDEN, golden_manifest()[name]["bioseq"]), r["K"], hd, out)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, r["csv"]), "rb").read()


def test_exhaustive_csv_writer_large_multi_threaded(oracle, tmp_path):
    """The writer formats runs of more than 2^20 lines in several tasks and threads: a 2.3 M-base
    chromosome, one shorter than K (the reference's mis-step), values above K (no line), against
    the literal restatement of the reference loop (hammings.cpp:2899-2929)."""
    rng = np.random.default_rng(5)
    K = 25
    entries = [("big", rng.integers(0, 4, size=2_300_000, dtype=np.uint8)), ("tiny", rng.integers(0, 4, size=11, dtype=np.uint8)),
               ("mid", rng.integers(0, 4, size=70_000, dtype=np.uint8)), ("last", rng.integers(0, 4, size=1_200_000, dtype=np.uint8))]
    seq = str(tmp_path / "g.seq")
    oracle.write_bioseq(seq, entries, title="big")
    concat, chroms, glen = hostlib.concat_from_bioseq(seq, K)
    hd = rng.integers(0, K + 3, size=len(concat)).astype(np.uint16)  # K+1, K+2: nothing to report there
    out = str(tmp_path / "o.csv")
    hostlib.write_exhaustive_csv(seq, K, hd, out)
    assert open(out, "rb").read() == oracle.exhaustive_csv(glen, chroms, K, hd)


@pytest.mark.parametrize("mr", targeted_runs(), ids=lambda mr: mr[1]["out"])
def test_restricted_writers_are_byte_exact(oracle, mr, tmp_path, monkeypatch):
    m, r = mr
    ents, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    pb = os.path.join(GOLDEN, m["probes"][r["probes"]]["bioseq"])
    concat, chroms, _ = oracle.concat_entries(oracle.read_bioseq(pb))
    h = oracle.restricted_per_loci(chroms, oracle.targeted_brute(tseq, concat, r["K"], r["R"], r["both"]))
    monkeypatch.chdir(tmp_path)  # the Wiggle header embeds the -o path as given
    hostlib.write_restricted(pb, r["K"], r["R"], r["fmt"], h, r["out"])
    assert open(r["out"], "rb").read() == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def test_sfx_reader(oracle):
    m = golden_manifest()["__targeted__"]
    ents, seq = hostlib.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    wents, wseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    assert ents == wents and np.array_equal(seq, wseq)


def test_fasta_front_end_equals_reference_genbioseq(oracle, tmp_path):
    for name, case in golden_manifest().items():
        if name.startswith("__"):
            continue
        out = str(tmp_path / (name + ".seq"))
        hostlib.fasta_to_bioseq(os.path.join(GOLDEN, case["fasta"]), out, name)
        mine = oracle.read_bioseq(out)
        ref = oracle.read_bioseq(os.path.join(GOLDEN, case["bioseq"]))
        assert [n for n, _ in mine] == [n for n, _ in ref]
        for (_, a), (_, b) in zip(mine, ref):
            assert np.array_equal(a, b)
        # and the C++ reader reads its own writer
        got, _, _ = hostlib.concat_from_bioseq(out, 25)
        want, _, _ = oracle.concat_entries(ref)
        assert np.array_equal(got, want)


def test_gzip_fasta_reads_like_plain_fasta(oracle, tmp_path):
    """read_fasta goes through zlib: .fa.gz, CRLF line ends and a last line without newline give the
    same entries as the plain file."""
    import gzip
    case = golden_manifest()["nonacgt"]
    plain = open(os.path.join(GOLDEN, case["fasta"]), "rb").read()
    variants = {"g.fa.gz": gzip.compress(plain), "crlf.fa": plain.replace(b"\n", b"\r\n"),
                "nonl.fa.gz": gzip.compress(plain.rstrip(b"\n"))}
    ref = oracle.read_bioseq(os.path.join(GOLDEN, case["bioseq"]))
    for fn, data in variants.items():
        path, out = str(tmp_path / fn), str(tmp_path / (fn + ".seq"))
        open(path, "wb").write(data)
        hostlib.fasta_to_bioseq(path, out, "x")
        mine = oracle.read_bioseq(out)
        assert [n for n, _ in mine] == [n for n, _ in ref], fn
        assert all(np.array_equal(a, b) for (_, a), (_, b) in zip(mine, ref)), fn
    bad = str(tmp_path / "bad.fa.gz")
    open(bad, "wb").write(gzip.compress(plain)[:-20])  # truncated stream: an error, not silent data
    with pytest.raises(RuntimeError):
        hostlib.fasta_to_bioseq(bad, str(tmp_path / "bad.seq"), "x")


def test_merge_matches_reference(tmp_path):
    m = golden_manifest()["__merge__"]
    into = str(tmp_path / "into.csv")
    shutil.copyfile(os.path.join(GOLDEN, m["into_before"]), into)
    hostlib.merge_csv(os.path.join(GOLDEN, m["from"]), into)
    assert open(into, "rb").read() == open(os.path.join(GOLDEN, m["into_after"]), "rb").read()
    fresh = str(tmp_path / "fresh.csv")
    hostlib.merge_csv(os.path.join(GOLDEN, m["from"]), fresh)
    assert open(fresh, "rb").read() == open(os.path.join(GOLDEN, m["copy_after"]), "rb").read()


def test_cli_syntax_variants(tmp_path):
    d = hostlib.parse_cli(["-m1", "-K25", "-c", "-i", "g.seq", "-o", "out.csv"])
    assert (d["mode"], d["K"], d["crick"], d["in_file"], d["out_file"]) == (1, 25, 1, "g.seq", "out.csv")
    d = hostlib.parse_cli(["--mode=1", "--seqlen", "50", "--strandcrick", "--in=g.seq"])
    assert (d["mode"], d["K"], d["crick"]) == (1, 50, 1)
    d = hostlib.parse_cli(["-m", "0", "--seq", "p.seq", "-r2", "-S2", "-i", "a.sfx", "--gpus=4"])
    assert (d["mode"], d["rhamm"], d["resformat"], d["in_seq_file"], d["gpus"]) == (0, 2, 2, "p.seq", 4)
    # defaults: restricted mode, K=100, Watson only, R=3 (hammings.cpp:312-332)
    d = hostlib.parse_cli(["-i", "x"])
    assert (d["mode"], d["K"], d["crick"], d["rhamm"], d["sample"]) == (0, 100, 0, 3, 1)
    # unique long-option prefixes, argtable-style integers
    d = hostlib.parse_cli(["--seql=0x20", "--sweeps=2", "-B", "1KB", "-i", "x", "-m1"])
    assert (d["K"], d["sweep_start"], d["sweep_end"]) == (32, 2, 1024)
    # grouped literals
    d = hostlib.parse_cli(["-cm1", "-i", "x"])
    assert d["crick"] == 1 and d["mode"] == 1
    # @parameter file: several options per line, comments, blank lines
    pf = tmp_path / "params.txt"
    pf.write_text("# comment\n-m1 -K31\n\n; other comment\n// c++ comment\n  -c\n-i genome.seq\n")
    d = hostlib.parse_cli(["@" + str(pf), "-o", "o.csv"])
    assert (d["mode"], d["K"], d["crick"], d["in_file"], d["out_file"]) == (1, 31, 1, "genome.seq", "o.csv")
    for bad in (["-m1"], ["-i", "x", "-K", "abc"], ["-i", "x", "--nosuch"], ["-i"], ["-i", "x", "stray"]):
        with pytest.raises(RuntimeError):
            hostlib.parse_cli(bad)


def test_cli_binary_exit_codes():
    import subprocess
    exe = hostlib.cli_path()
    assert os.path.exists(exe)
    assert subprocess.run([exe, "-h"], capture_output=True).returncode == 1      # hammings.cpp:255-265
    assert subprocess.run([exe, "-v"], capture_output=True).returncode == 1
    assert subprocess.run([exe], capture_output=True).returncode == 1            # -i is required
    assert subprocess.run([exe, "hammings", "-m1", "-K5", "-i", "x"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "-m0", "-K32", "-r9", "-i", "x"], capture_output=True).returncode == 1  # K/(R+1) < 4


def test_bham_round_trip(tmp_path):
    """-m4 then -m5: every "chrom",loci,dist row survives; the CSV carries the reference's -m5 header."""
    src = os.path.join(GOLDEN, "tiny2.K25c.csv")
    b, back = str(tmp_path / "h.bham"), str(tmp_path / "back.csv")
    hostlib.csv_to_bham(src, b)
    raw = open(b, "rb").read()
    assert raw[:4] == b"bham" and int.from_bytes(raw[8:12], "little") == len(raw)
    hostlib.bham_to_csv(b, back)
    got = open(back).read().split("\n")
    want = open(src).read().split("\n")
    assert got[0] == '"Chrom","Loci","Hamming"'
    assert got[1:] == want[1:]  # the "G,2,G" descriptor row is dropped, as in the reference
    import subprocess
    exe = hostlib.cli_path()
    b2, back2 = str(tmp_path / "h2.bham"), str(tmp_path / "back2.csv")
    assert subprocess.run([exe, "-m4", "-i", src, "-o", b2], capture_output=True).returncode == 0
    assert subprocess.run([exe, "-m5", "-i", b2, "-o", back2], capture_output=True).returncode == 0
    assert open(b2, "rb").read() == raw and open(back2).read() == open(back).read()


def _expected_distribution(values):
    """HammingDist's file layout (HammingDist/HammingDist.cpp:616-700) for a list of distances, field-3
    semantics: rows 0 .. max-1 (the reference's loops stop before the largest value), %f proportions."""
    counts = np.bincount(np.asarray(values, dtype=np.int64))
    maxh = len(counts) - 1
    out = ',"All","Proportion All","Cumulative All"'
    total = int(counts[:maxh].sum())
    cum = 0.0
    for d in range(maxh):
        prop = counts[d] / total if total else 0.0
        cum += prop
        out += "\n%d,%d,%f,%f" % (d, counts[d], prop, cum)
    return out.encode()


def test_hammingdist_distribution_of_exhaustive_csv(tmp_path):
    """region-less HammingDist: rows of a hammings -m1 CSV (descriptor row skipped) and of its -m5 form
    (header row skipped) give the same distribution file; two inputs add up; the binary takes the
    reference's flags."""
    import subprocess
    src = os.path.join(GOLDEN, "multiword.K50c.csv")
    values = [int(line.rsplit(",", 1)[1]) for line in open(src).read().split("\n")[1:] if line]
    want = _expected_distribution(values)
    out = str(tmp_path / "dist.csv")
    hostlib.hamming_dist([src], out)
    assert open(out, "rb").read() == want
    b, back = str(tmp_path / "h.bham"), str(tmp_path / "back.csv")
    hostlib.csv_to_bham(src, b)
    hostlib.bham_to_csv(b, back)  # carries the "Chrom","Loci","Hamming" header line
    hostlib.hamming_dist([back], out)
    assert open(out, "rb").read() == want
    hostlib.hamming_dist([src, back], out)
    assert open(out, "rb").read() == _expected_distribution(values + values)
    exe = os.path.join(os.path.dirname(hostlib.cli_path()), "k4b_hammingdist")
    out2 = str(tmp_path / "dist2.csv")
    p = subprocess.run([exe, "-m0", "-s0", "-r2000", "-i", str(tmp_path / "*.csv").replace("*.csv", "back.csv"), "--incsv=" + src,
                        "-o", out2], capture_output=True)
    assert p.returncode == 0, p.stdout
    assert open(out2, "rb").read() == _expected_distribution(values + values)
    assert subprocess.run([exe, "-h"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "-i", src], capture_output=True).returncode == 1           # -o is required
    assert subprocess.run([exe, "-i", src, "-o", out2, "-I", str(tmp_path / "nosuch.bed")], capture_output=True).returncode == 1
    bad = str(tmp_path / "bad.csv")
    open(bad, "w").write('"chr1",5\n')
    with pytest.raises(RuntimeError):
        hostlib.hamming_dist([bad], out)
