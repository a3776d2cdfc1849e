"""GPU parity tests: the CUDA path, called through the C ABI, against (1) the golden files
written by the unmodified reference, (2) the pinned oracle on seeded inputs, and (3)
size-independent properties at BASELINE.json's config-1 size.  Bit-exact everywhere (integer
work)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, case_id, golden_cases, load_case, random_genome

import kit4b_b200 as k4b

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _engine():
    k4b.gpu_init(1)
    yield
    k4b.gpu_shutdown()


@pytest.mark.parametrize("case", golden_cases(), ids=case_id)
def test_cuda_matches_reference_golden_files(oracle, case):
    name, r = case
    concat, chroms, glen = load_case(oracle, name)
    hd = k4b.exhaustive(concat, r["K"], r["both"])
    gold = open(os.path.join(GOLDEN, r["csv"]), "rb").read()
    assert oracle.exhaustive_csv(glen, chroms, r["K"], hd) == gold


RANDOM_CASES = [
    # seed, chromosome lengths, K, both strands, alphabet size
    (101, [5000], 10, True, 4),
    (102, [3000, 2500], 25, True, 4),
    (103, [4000, 31, 32, 33, 1000], 32, True, 4),
    (104, [6000], 33, False, 4),
    (105, [2000, 2000, 2000], 50, True, 4),
    (106, [5000], 64, True, 4),
    (107, [3000, 3000], 96, True, 4),
    (108, [4000], 100, True, 4),
    (109, [4100], 128, False, 4),
    (110, [2500, 500], 129, True, 4),   # generic path
    (111, [3000], 500, True, 4),        # generic path, many words
    (112, [4000, 2000], 25, True, 7),   # N / Undef / InDel symbols: three-plane path
    (113, [3000], 100, True, 5),
    (114, [9000, 9000], 16, True, 2),   # low-complexity: many zero distances, early exit
    (115, [20000], 25, True, 4),        # spans several tiles and query blocks
]


@pytest.mark.parametrize("seed,lens,K,both,alpha", RANDOM_CASES, ids=lambda v: str(v).replace(" ", ""))
def test_cuda_matches_oracle_on_seeded_inputs(oracle, seed, lens, K, both, alpha):
    c = random_genome(seed, lens, alpha)
    want = oracle.exhaustive_sliding(c, K, both)
    got = k4b.exhaustive(c, K, both)
    assert np.array_equal(got, want)


def test_shards_concatenate_to_the_full_result(oracle):
    c = random_genome(201, [7000, 5000])
    K = 25
    want = oracle.exhaustive_sliding(c, K, True)
    out = np.full(len(c), K + 1, dtype=np.uint16)
    cuts = [0, 1, 2049, 5000, 7003, 9999, len(c)]
    for b, e in zip(cuts[:-1], cuts[1:]):
        k4b.exhaustive_shard(c, K, True, b, e, out)
    assert np.array_equal(out, want)


def test_edge_cases(oracle):
    # all chromosomes shorter than K -> nothing lowered
    c = random_genome(301, [8, 9, 5])
    assert (k4b.exhaustive(c, 12, True) == 13).all()
    # sequence shorter than K
    assert (k4b.exhaustive(np.zeros(7, np.uint8), 10, True) == 11).all()
    # a single K-mer that is its own reverse complement
    c = np.array([0, 1, 2, 3] * 3, dtype=np.uint8)
    assert k4b.exhaustive(c, 12, False)[0] == 13
    assert k4b.exhaustive(c, 12, True)[0] == 0
    # K bounds (hammings.cpp:36-38)
    for bad in (9, 5001):
        with pytest.raises(k4b.K4BError):
            k4b.exhaustive(random_genome(1, [100]), bad, True)
    # length exactly K and K+1
    for n in (25, 26):
        c = random_genome(302 + n, [n])
        assert np.array_equal(k4b.exhaustive(c, 25, True), oracle.exhaustive_brute(c, 25, True))


def test_targeted_matches_oracle(oracle):
    rng = np.random.default_rng(401)
    target = random_genome(402, [30000, 20000])
    # probes: mutated copies of target regions (forward and reverse complement) + random
    parts = []
    for start, nmut, rc in [(1000, 0, False), (5000, 1, False), (9000, 2, True), (12000, 3, False),
                            (31000, 4, True), (35000, 6, False)]:
        seg = target[start:start + 200].copy()
        for p in rng.choice(200, size=nmut * 6, replace=False):
            seg[p] = (seg[p] + 1 + rng.integers(0, 3)) % 4
        if rc:
            seg = oracle.CPL[seg[::-1]]
        parts.append(seg)
    parts.append(rng.integers(0, 4, size=500, dtype=np.uint8))
    probes = []
    for i, p in enumerate(parts):
        probes.append(p)
        if i + 1 < len(parts):
            probes.append(np.array([7], dtype=np.uint8))
    probes = np.ascontiguousarray(np.concatenate(probes))
    for K, R in [(32, 3), (25, 3), (20, 1), (50, 5)]:
        want = oracle.targeted_brute(target, probes, K, R, True)
        got = k4b.targeted(target, probes, K, R, True)
        assert np.array_equal(got, want), (K, R)


@pytest.mark.parametrize("K", [25])
def test_config1_size_properties(oracle, K):
    """1 Mbp, K=25, both strands (BASELINE.json configs[0]): properties that hold at any size."""
    G = 1_000_000
    rng = np.random.default_rng(12)
    c = rng.integers(0, 4, size=G, dtype=np.uint8)
    # plant an exact duplicate, a reverse-complement copy and a 2-mismatch copy
    c[700000:700060] = c[1000:1060]
    c[800000:800060] = oracle.CPL[c[2000:2060][::-1]]
    c[900000:900040] = c[3000:3040]
    c[900010] = (c[900010] + 1) % 4
    c[900030] = (c[900030] + 2) % 4
    c = np.ascontiguousarray(c)
    hd = k4b.exhaustive(c, K, True)
    n = G - K + 1
    assert (hd[n:] == K + 1).all() and (hd[:n] <= K).all()
    assert (hd[1000:1000 + 36] == 0).all() and (hd[700000:700000 + 36] == 0).all()
    assert (hd[2000:2000 + 36] == 0).all() and (hd[800000:800000 + 36] == 0).all()
    assert hd[3000] <= 2 and hd[900000] <= 2
    # Watson-only result can never be below the both-strand result
    hw = k4b.exhaustive(c, K, False)
    assert (hw >= hd).all()
    assert (hw[2000:2000 + 36] > 0).any()
    # spot check 48 random K-mers with an O(G*K) brute force on the host
    X = np.lib.stride_tricks.sliding_window_view(c, K)
    RCX = oracle.CPL[X[:, ::-1]]
    for i in rng.choice(n, size=48, replace=False):
        d = (X != X[i]).sum(1)
        d[i] = K + 1
        best = min(d.min(), (RCX != X[i]).sum(1).min())
        assert hd[i] == best, i
    # idempotence: a second run gives the same array
    assert np.array_equal(hd, k4b.exhaustive(c, K, True))


# ---- the drop-in binary end to end: same flags in, same bytes out ---------------------------
def _run_cli(args, cwd=None):
    import subprocess
    from kit4b_b200 import hostlib
    p = subprocess.run([hostlib.cli_path()] + args, cwd=cwd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


@pytest.mark.parametrize("case", golden_cases(), ids=case_id)
def test_cli_exhaustive_output_file_equals_reference(case, tmp_path):
    from conftest import golden_manifest
    name, r = case
    out = str(tmp_path / "out.csv")
    args = ["hammings", "-m1", "-K%d" % r["K"], "-T3", "-i", os.path.join(GOLDEN, golden_manifest()[name]["bioseq"]),
            "-o", out]
    if r["both"]:
        args.insert(2, "-c")
    _run_cli(args)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, r["csv"]), "rb").read()


def _targeted_runs():
    from conftest import golden_manifest
    m = golden_manifest()["__targeted__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_runs(), ids=lambda mr: mr[1]["out"])
def test_cli_targeted_output_file_equals_reference(mr, tmp_path):
    m, r = mr
    args = ["hammings", "-m0", "-K%d" % r["K"], "-r%d" % r["R"], "-S%d" % r["fmt"], "-i", os.path.join(GOLDEN, m["sfx"]),
            "-I", os.path.join(GOLDEN, m["probes"][r["probes"]]["bioseq"]), "-o", r["out"]]
    if r["both"]:
        args.insert(2, "-c")
    _run_cli(args, cwd=str(tmp_path))  # relative -o: the Wiggle header embeds the path as given
    assert open(os.path.join(str(tmp_path), r["out"]), "rb").read() == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def _gz(src, dst):
    import gzip
    import shutil
    with open(src, "rb") as fi, gzip.open(dst, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return dst


@pytest.mark.parametrize("case", golden_cases()[::2], ids=case_id)
def test_cli_reads_fasta_and_gzip_directly(case, tmp_path):
    """-i genome.fa / genome.fa.gz (no genbioseq step): same bytes as the reference produced from the
    bioseq the reference's own genbioseq made of that FASTA (Fasta.cpp:967-1209, genbioseq.cpp:356-446)."""
    from conftest import golden_manifest
    name, r = case
    fa = os.path.join(GOLDEN, golden_manifest()[name]["fasta"])
    want = open(os.path.join(GOLDEN, r["csv"]), "rb").read()
    for src in (fa, _gz(fa, str(tmp_path / "g.fa.gz"))):
        out = str(tmp_path / "out.csv")
        _run_cli(["hammings", "-m1"] + (["-c"] if r["both"] else []) + ["-K%d" % r["K"], "-i", src, "-o", out])
        assert open(out, "rb").read() == want, src


@pytest.mark.parametrize("mr", _targeted_runs()[::2], ids=lambda mr: mr[1]["out"])
def test_cli_targeted_from_fasta_without_suffix_array_file(mr, tmp_path):
    """-m0 -i assembly.fa(.gz) -I probes.fa(.gz): the engines need only the sequence area of the
    suffix-array file, which the entries determine - same bytes as the reference run on .sfx + .seq."""
    m, r = mr
    tfa = os.path.join(GOLDEN, m["target_fasta"])
    pfa = os.path.join(GOLDEN, m["probes"][r["probes"]]["fasta"])
    for ti, pi in ((tfa, _gz(pfa, str(tmp_path / "p.fa.gz"))), (_gz(tfa, str(tmp_path / "t.fa.gz")), pfa)):
        args = ["hammings", "-m0"] + (["-c"] if r["both"] else []) + ["-K%d" % r["K"], "-r%d" % r["R"], "-S%d" % r["fmt"],
                                                                      "-i", ti, "-I", pi, "-o", r["out"]]
        _run_cli(args, cwd=str(tmp_path))
        assert open(os.path.join(str(tmp_path), r["out"]), "rb").read() == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def test_targeted_wildcards_and_n_rule(oracle):
    """probe symbols >= N are wildcards, > 4 of them report 0, target N never matches."""
    rng = np.random.default_rng(501)
    target = random_genome(502, [6000, 4000]).copy()
    target[1000:1003] = 4            # N run inside the target
    target[2500] = 6
    probes = target[900:1500].copy()
    for p in (50, 130, 131, 200, 260, 261, 262, 263, 264, 400):
        probes[p] = 4
    probes[300] = 5
    probes = np.ascontiguousarray(np.concatenate([probes, [7], rng.integers(0, 5, size=400, dtype=np.uint8)]),
                                  dtype=np.uint8)
    for K, R, both in [(32, 3, True), (25, 2, False), (64, 5, True), (140, 9, True)]:
        want = oracle.targeted_brute(target, probes, K, R, both)
        got = k4b.targeted(target, probes, K, R, both)
        assert np.array_equal(got, want), (K, R, both)


def test_in_process_multi_gpu_shards(oracle):
    """k4b_gpu_init(n>1): one process drives several GPUs - NCCL broadcast of the packed set,
    query shards per device, minima concatenated on the host.  Needs >= 2 visible GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    k4b.gpu_shutdown()
    try:
        k4b.gpu_init(min(n, 8))
        assert k4b.gpu_count() == min(n, 8)
        c = random_genome(601, [30000, 12000, 9000])
        for K, both in [(25, True), (100, False)]:
            assert np.array_equal(k4b.exhaustive(c, K, both), oracle.exhaustive_sliding(c, K, both))
        # diagonal-band engine: pair matrix partitioned over the devices + ncclAllReduce(min)
        from kit4b_b200 import hamm
        k4b.set_engine(hamm.ENGINE_DIAG)
        try:
            for K, both in [(25, True), (50, True), (100, False)]:
                assert np.array_equal(k4b.exhaustive(c, K, both), oracle.exhaustive_sliding(c, K, both))
            big = random_genome(603, [400000, 150000])
            got = k4b.exhaustive(big, 32, True)
            k4b.set_engine(hamm.ENGINE_POPC)
            assert np.array_equal(got, k4b.exhaustive(big, 32, True))
        finally:
            k4b.set_engine(hamm.ENGINE_AUTO)
        target = random_genome(602, [40000])
        probes = np.ascontiguousarray(np.concatenate([target[100:900], [7], target[5000:5600]]), dtype=np.uint8)
        probes[50] = (probes[50] + 1) % 4
        assert np.array_equal(k4b.targeted(target, probes, 32, 3, True), oracle.targeted_brute(target, probes, 32, 3, True))
        # targeted on every engine: probes split over the devices (seed), diagonals split (bands);
        # wildcard probes ride along on the POPC engine
        probes_n = probes.copy()
        probes_n[[200, 201, 1000]] = 4
        want = oracle.targeted_brute(target, probes, 32, 3, True)
        want_n = oracle.targeted_brute(target, probes_n, 32, 3, True)
        for eng in (hamm.ENGINE_SEED, hamm.ENGINE_DIAG, hamm.ENGINE_POPC):
            k4b.set_engine(eng)
            try:
                assert np.array_equal(k4b.targeted(target, probes, 32, 3, True), want), eng
                assert np.array_equal(k4b.targeted(target, probes_n, 32, 3, True), want_n), eng
            finally:
                k4b.set_engine(hamm.ENGINE_AUTO)
    finally:
        k4b.gpu_shutdown()
        k4b.gpu_init(1)


def _targeted_self_runs():
    from conftest import golden_manifest
    m = golden_manifest()["__targeted_self__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_self_runs(), ids=lambda mr: mr[1]["out"])
def test_cli_targeted_without_probe_file_equals_reference(oracle, mr, tmp_path):
    """-m0 without -I: exact self hits skipped on the sense strand, wildcards, cap."""
    m, r = mr
    args = ["hammings", "-m0", "-K%d" % r["K"], "-r%d" % r["R"], "-S%d" % r["fmt"], "-i", os.path.join(GOLDEN, m["sfx"]),
            "-o", r["out"]]
    if r["both"]:
        args.insert(2, "-c")
    z = r.get("z", 0)
    if z:
        args.insert(2, "-z%d" % z)  # intra / inter filter of exact sense hits
    _run_cli(args, cwd=str(tmp_path))
    assert open(os.path.join(str(tmp_path), r["out"]), "rb").read() == open(os.path.join(GOLDEN, r["out"]), "rb").read()
    # and the array-level API against the oracle
    _, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    assert np.array_equal(k4b.targeted(tseq, None, r["K"], r["R"], r["both"], intra_inter_both=z),
                          oracle.targeted_self_brute(tseq, r["K"], r["R"], r["both"], z))


def _sweep_runs():
    from conftest import golden_manifest
    m = golden_manifest()["__sweeps__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _sweep_runs(), ids=lambda mr: mr[1]["csv"])
def test_cli_sweep_subranges_and_node_slices_equal_reference(mr, tmp_path):
    m, r = mr
    out = str(tmp_path / "o.csv")
    args = ["hammings", "-m%d" % r["mode"], "-K%d" % r["K"], "-i", os.path.join(GOLDEN, m["bioseq"]), "-o", out]
    args += ["-b%d" % r["b"], "-B%d" % r["B"]] if r["mode"] == 1 else ["-n%d" % r["n"], "-N%d" % r["N"]]
    if r["both"]:
        args.insert(2, "-c")
    _run_cli(args)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, r["csv"]), "rb").read()


def test_sweep_subranges_match_oracle_on_seeded_inputs(oracle):
    for seed, lens, K, both, ss, se in [(701, [6000, 3000], 25, True, 1, 100), (702, [9000], 50, True, 4000, 6000),
                                        (703, [5000, 40, 2500], 100, True, 2, 0), (704, [7000], 32, False, 10, 10),
                                        (705, [3000, 3000], 150, True, 100, 2500), (706, [4000], 25, True, 3990, 0)]:
        c = random_genome(seed, lens)
        glen = len(c) + 2
        end = glen if se == 0 else se
        want = oracle.exhaustive_sliding_sweep(c, K, both, ss, end)
        got = k4b.exhaustive(c, K, both, ss, se)
        assert np.array_equal(got, want), (seed, K, ss, se)


def test_config1_full_output_md5_equals_reference_run(tmp_path):
    """BASELINE.json configs[0] / BASELINE.md: 1 Mbp genome (Python random seed 12), K=25, both
    strands.  The unmodified reference needed 36 min on 8 cores for this file; its output md5 and
    the md5 of the input FASTA are recorded in BASELINE.md.  Same flags in, same bytes out."""
    import hashlib
    import random
    from kit4b_b200 import hostlib
    random.seed(12)
    s = "".join(random.choices("ACGT", k=1_000_000))
    fa = ">chr1 synthetic\n" + "".join(s[i:i + 80] + "\n" for i in range(0, len(s), 80))
    assert hashlib.md5(fa.encode()).hexdigest() == "e02b7eebf84f54c2c205567402029947"
    fa_path, seq, out = str(tmp_path / "g.fa"), str(tmp_path / "g.seq"), str(tmp_path / "out.csv")
    open(fa_path, "w").write(fa)
    hostlib.fasta_to_bioseq(fa_path, seq, "cfg1")
    _run_cli(["hammings", "-m1", "-K25", "-c", "-T8", "-i", seq, "-o", out])
    data = open(out, "rb").read()
    assert len(data) == 15_888_524 and data.count(b"\n") == 999_977
    assert hashlib.md5(data).hexdigest() == "f1b2e85a8f64dc68dcc60cc2403cfb1a"


def test_gpu_histogram_and_cli_distribution_file(tmp_path):
    """k4b_hamm_histogram (shared-memory privatised GPU histogram of the minima) equals np.bincount, and
    `k4b_hammings --dist` writes what the HammingDist drop-in makes of the CSV of the same run."""
    from conftest import golden_manifest
    from kit4b_b200 import hamm, hostlib
    rng = np.random.default_rng(77)
    for K, n in ((25, 1_000_003), (5000, 70_001), (12, 5)):
        v = rng.integers(0, K + 2, size=n).astype(np.uint16)
        assert np.array_equal(hamm.histogram(v, K), np.bincount(v, minlength=K + 2).astype(np.uint64))
    seq = os.path.join(GOLDEN, golden_manifest()["multiword"]["bioseq"])
    csv, dist, want = str(tmp_path / "o.csv"), str(tmp_path / "d.csv"), str(tmp_path / "w.csv")
    _run_cli(["hammings", "-m1", "-c", "-K50", "-i", seq, "-o", csv, "--dist=" + dist])
    assert open(csv, "rb").read() == open(os.path.join(GOLDEN, "multiword.K50c.csv"), "rb").read()
    hostlib.hamming_dist([csv], want)
    assert open(dist, "rb").read() == open(want, "rb").read()
