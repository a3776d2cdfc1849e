"""CPU tests: the oracle against the golden vectors written by the unmodified reference
(tests/golden/make_golden.py), i.e. the pin that every GPU parity test leans on."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_id, golden_cases, golden_manifest, load_case, random_genome


@pytest.mark.parametrize("case", golden_cases(), ids=case_id)
def test_sliding_oracle_matches_reference_output(oracle, case):
    name, r = case
    concat, chroms, glen = load_case(oracle, name)
    hd = oracle.exhaustive_sliding(concat, r["K"], r["both"], threads=3)
    gold = open(os.path.join(GOLDEN, r["csv"]), "rb").read()
    assert oracle.exhaustive_csv(glen, chroms, r["K"], hd) == gold


@pytest.mark.parametrize("case", [c for c in golden_cases() if c[0] in ("adversarial", "nonacgt")], ids=case_id)
def test_brute_and_numpy_oracles_match_reference_output(oracle, case):
    name, r = case
    concat, chroms, glen = load_case(oracle, name)
    gold = open(os.path.join(GOLDEN, r["csv"]), "rb").read()
    assert oracle.exhaustive_csv(glen, chroms, r["K"], oracle.exhaustive_brute(concat, r["K"], r["both"])) == gold
    assert oracle.exhaustive_csv(glen, chroms, r["K"], oracle.numpy_brute(concat, r["K"], r["both"])) == gold


def test_fasta_encoding_equals_reference_bioseq(oracle):
    for name, case in golden_manifest().items():
        if name.startswith("__"):
            continue
        fa = oracle.encode_fasta(open(os.path.join(GOLDEN, case["fasta"])).read())
        bs = oracle.read_bioseq(os.path.join(GOLDEN, case["bioseq"]))
        assert [n for n, _ in fa] == [n for n, _ in bs]
        for (_, a), (_, b) in zip(fa, bs):
            assert np.array_equal(a, b)


def test_bioseq_writer_round_trip(oracle, tmp_path):
    entries = oracle.read_bioseq(os.path.join(GOLDEN, "adversarial.seq"))
    p = str(tmp_path / "rt.seq")
    oracle.write_bioseq(p, entries)
    back = oracle.read_bioseq(p)
    assert [n for n, _ in back] == [n for n, _ in entries]
    for (_, a), (_, b) in zip(entries, back):
        assert np.array_equal(a, b)


def test_oracle_formulations_agree_on_random_inputs(oracle):
    for seed, lens, K, both in [(1, [700, 300], 10, True), (2, [900], 31, True), (3, [400, 20, 500], 32, False),
                                (4, [1200], 70, True), (5, [50, 60, 70], 64, True)]:
        c = random_genome(seed, lens)
        a = oracle.exhaustive_sliding(c, K, both, threads=2)
        b = oracle.exhaustive_brute(c, K, both)
        assert np.array_equal(a, b), (seed, K)


def test_oracle_edge_cases(oracle):
    # every chromosome shorter than K: nothing is lowered
    c = random_genome(9, [8, 9, 5])
    assert (oracle.exhaustive_sliding(c, 12, True, threads=1) == 13).all()
    # a single K-mer: Watson-only has no partner, Crick compares it with its own reverse complement
    c = np.array([0, 1, 2, 3] * 3, dtype=np.uint8)  # ACGTACGTACGT is its own reverse complement
    assert oracle.exhaustive_sliding(c, 12, False, threads=1)[0] == 13
    assert oracle.exhaustive_sliding(c, 12, True, threads=1)[0] == 0


def test_sampled_sliding_is_an_upper_bound(oracle):
    c = random_genome(21, [1500])
    full = oracle.exhaustive_sliding(c, 25, True, threads=2)
    part, cells = oracle.exhaustive_sliding_sample(c, 25, True, 2, 1, 4)
    assert cells > 0 and (part >= full).all()


def _targeted_runs():
    m = golden_manifest()["__targeted__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_runs(), ids=lambda mr: mr[1]["out"])
def test_targeted_oracle_matches_reference_output(oracle, mr):
    """-m0 -I: oracle array + restated writers == the reference's CSV / BED / Wiggle files."""
    m, r = mr
    _, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    concat, chroms, _ = oracle.concat_entries(oracle.read_bioseq(os.path.join(GOLDEN, m["probes"][r["probes"]]["bioseq"])))
    h = oracle.restricted_per_loci(chroms, oracle.targeted_brute(tseq, concat, r["K"], r["R"], r["both"]))
    rep = oracle.restricted_report(chroms, r["K"], r["R"], h, r["fmt"], out_name=r["out"])
    assert rep == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def _targeted_self_runs():
    m = golden_manifest()["__targeted_self__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_self_runs(), ids=lambda mr: mr[1]["out"])
def test_targeted_self_oracle_matches_reference_output(oracle, mr):
    """-m0 without -I (K-mers of the indexed assembly against itself)."""
    m, r = mr
    ents, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    h = oracle.restricted_per_loci(ents, oracle.targeted_self_brute(tseq, r["K"], r["R"], r["both"], r.get("z", 0)))
    rep = oracle.restricted_report(ents, r["K"], r["R"], h, r["fmt"], out_name=r["out"])
    assert rep == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def _sweep_runs():
    m = golden_manifest()["__sweeps__"]
    return [(m, r) for r in m["runs"]]


def _sweep_range(r, glen, nchroms):
    from kit4b_b200 import hostlib
    if r["mode"] == 2:
        return hostlib.node_sweep_range(glen, nchroms, not r["both"], r["n"], r["N"])
    return min(r["b"], glen), (glen if r["B"] == 0 else min(r["B"], glen))


@pytest.mark.parametrize("mr", _sweep_runs(), ids=lambda mr: mr[1]["csv"])
def test_sweep_subranges_and_node_slices_match_reference_output(oracle, mr):
    """-b/-B sub-ranges and -m2 slices: oracle pair selection + host range arithmetic."""
    m, r = mr
    concat, chroms, glen = oracle.concat_entries(oracle.read_bioseq(os.path.join(GOLDEN, m["bioseq"])))
    ss, se = _sweep_range(r, glen, len(chroms))
    hd = oracle.exhaustive_sliding_sweep(concat, r["K"], r["both"], ss, se)
    assert oracle.exhaustive_csv(glen, chroms, r["K"], hd, ss, se) == open(os.path.join(GOLDEN, r["csv"]), "rb").read()


def test_merged_node_slices_equal_the_single_node_run(tmp_path):
    """-m2 slices folded together with -m3 reproduce the -m1 values (the reference's multi-node flow)."""
    import shutil
    from kit4b_b200 import hostlib
    into = str(tmp_path / "merged.csv")
    shutil.copyfile(os.path.join(GOLDEN, "slice.K12c.n3N1.csv"), into)
    for n in (2, 3):
        hostlib.merge_csv(os.path.join(GOLDEN, "slice.K12c.n3N%d.csv" % n), into)
    merged = open(into).read().split("\n")[1:]
    full = open(os.path.join(GOLDEN, "adversarial.K12c.csv")).read().strip("\n").split("\n")[1:]
    assert merged == full


def test_reference_depth_cut_case(oracle):
    """tests/golden/depth_case.py: at the default sensitivity (and at -s3) the reference gives a core with
    42000 copies up and reports "not found" (4) for a K-mer whose true minimum is 3; at -s1 / -s2 it finds
    it.  The oracle is exact: it reproduces the -s1 / -s2 files and differs from -s0 / -s3 in that one K-mer."""
    import sys
    sys.path.insert(0, GOLDEN)
    import depth_case
    asm, probes, upos = depth_case.build()
    target = np.concatenate([np.concatenate([c, [7]]) for _, c in asm]).astype(np.uint8)
    concat, chroms, _ = oracle.concat_entries(probes)
    h = oracle.targeted_brute(target, concat, depth_case.K, depth_case.R, True)
    assert h[upos] == 3
    rep = oracle.restricted_report(chroms, depth_case.K, depth_case.R, oracle.restricted_per_loci(chroms, h), 0)
    gold = {s: open(os.path.join(GOLDEN, "depth.K32r3c.s%d.csv" % s), "rb").read() for s in range(4)}
    assert rep == gold[1] == gold[2]
    assert gold[0] == gold[3] != rep
    h4 = h.copy()
    h4[upos] = 4  # what the truncated search reports
    assert oracle.restricted_report(chroms, depth_case.K, depth_case.R, oracle.restricted_per_loci(chroms, h4), 0) == gold[0]
