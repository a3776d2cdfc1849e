"""CPU tests of the checker code inside bench.py: its NumPy brute force (the third, library-free
derivation the `parity` fields rest on), the valid-start mask, the FASTA writer and the workload
synthesis must agree with the oracle, otherwise a green `parity` would mean nothing."""
import os

import numpy as np

import bench
from conftest import random_genome


def test_numpy_brute_force_equals_oracle(oracle):
    c = random_genome(11, [700, 40, 500])
    c[100:140] = c[600:640]                      # a duplicate
    c[300:330] = oracle.CPL[c[50:80][::-1]]      # a reverse-complement copy
    c = np.ascontiguousarray(c)
    for K, both in ((12, True), (25, True), (31, False)):
        want = oracle.exhaustive_brute(c, K, both)
        valid = bench.valid_mask(c, K)
        assert np.array_equal(valid, oracle.valid_starts(c, K))
        for p in np.flatnonzero(valid)[::37]:
            assert bench.numpy_min_distance(c, valid, K, c[p:p + K], both, int(p)) == want[p], (K, both, p)


def test_numpy_brute_force_targeted_rules(oracle):
    target = random_genome(12, [3000, 1500])
    probes = np.ascontiguousarray(np.concatenate([target[200:400], [7], random_genome(13, [150])]), dtype=np.uint8)
    probes[30] = (probes[30] + 1) % 4
    K, R = 32, 3
    clamp = K // (K // (R + 1))
    want = oracle.targeted_brute(target, probes, K, R, True)
    tvalid = bench.valid_mask(target, K)
    for p in np.flatnonzero(bench.valid_mask(probes, K))[::17]:
        got = min(clamp, bench.numpy_min_distance(target, tvalid, K, probes[p:p + K], True, None))
        assert got == want[p], p


def test_fasta_writer_round_trip(oracle, tmp_path):
    from kit4b_b200 import hostlib
    ents = [("chr1", random_genome(14, [161])[:161]), ("chr2", random_genome(15, [80])[:80]), ("chr3", random_genome(16, [7])[:7])]
    fa, seq = str(tmp_path / "x.fa"), str(tmp_path / "x.seq")
    bench.write_fasta(fa, ents)
    text = open(fa).read().split("\n")
    assert text[0] == ">chr1" and len(text[1]) == 80 and len(text[3]) == 1  # 161 = 80 + 80 + 1
    hostlib.fasta_to_bioseq(fa, seq, "x")
    back = oracle.read_bioseq(seq)
    assert [n for n, _ in back] == ["chr1", "chr2", "chr3"]
    assert all(np.array_equal(a, b) for (_, a), (_, b) in zip(back, ents))


def test_workload_synthesis_is_pinned():
    """the seeded genomes behind the published checksums must not drift"""
    concat, chroms, K, both = bench.synth_genome("cfg1")
    assert (len(concat), K, both, len(chroms)) == (1_000_000, 25, True, 1)
    concat, chroms, K, both = bench.synth_genome("cfg2")
    assert len(concat) == 10_000_004 and bench.valid_count(chroms, K) == 9_999_755
    assert int(concat[:1000].astype(np.int64).sum()) == int(bench.synth_genome("cfg2")[0][:1000].astype(np.int64).sum())
    assert set(bench.ALL_CONFIGS) >= {"cfg1", "cfg3", "cfg4", "popc"}
