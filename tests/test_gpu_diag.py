"""GPU parity tests of the diagonal-band engine (k4b_diag.cu): same goldens, same oracle, and a
cross-check against the POPC all-pairs engine at sizes the CPU oracle cannot reach."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, case_id, golden_cases, golden_manifest, load_case, random_genome

import kit4b_b200 as k4b
from kit4b_b200 import hamm, hostlib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _engine():
    k4b.gpu_init(1)
    k4b.set_engine(hamm.ENGINE_DIAG)
    yield
    k4b.set_engine(hamm.ENGINE_AUTO)
    k4b.gpu_shutdown()


@pytest.mark.parametrize("case", golden_cases(), ids=case_id)
def test_diag_engine_matches_reference_golden_files(oracle, case):
    name, r = case
    concat, chroms, glen = load_case(oracle, name)
    hd = k4b.exhaustive(concat, r["K"], r["both"])
    assert oracle.exhaustive_csv(glen, chroms, r["K"], hd) == open(os.path.join(GOLDEN, r["csv"]), "rb").read()


CASES = [
    (101, [5000], 10, True, 4), (102, [3000, 2500], 25, True, 4), (103, [4000, 31, 32, 33, 1000], 32, True, 4),
    (104, [6000], 33, False, 4), (105, [2000, 2000, 2000], 50, True, 4), (106, [5000], 64, True, 4),
    (108, [4000], 100, True, 4), (110, [2500, 500], 129, True, 4), (111, [3000], 500, True, 4),
    (112, [4000, 2000], 25, True, 7), (113, [3000], 100, True, 5), (114, [9000, 9000], 16, True, 2),
    (115, [20000], 25, True, 4), (116, [40000, 9000], 31, True, 4), (117, [33000], 50, False, 4),
]


@pytest.mark.parametrize("seed,lens,K,both,alpha", CASES, ids=lambda v: str(v).replace(" ", ""))
def test_diag_engine_matches_oracle_on_seeded_inputs(oracle, seed, lens, K, both, alpha):
    c = random_genome(seed, lens, alpha)
    assert np.array_equal(k4b.exhaustive(c, K, both), oracle.exhaustive_sliding(c, K, both))


def test_diag_engine_edge_cases(oracle):
    c = random_genome(301, [8, 9, 5])
    assert (k4b.exhaustive(c, 12, True) == 13).all()
    assert (k4b.exhaustive(np.zeros(7, np.uint8), 10, True) == 11).all()
    c = np.array([0, 1, 2, 3] * 3, dtype=np.uint8)
    assert k4b.exhaustive(c, 12, False)[0] == 13
    assert k4b.exhaustive(c, 12, True)[0] == 0
    for n in (25, 26, 57):
        c = random_genome(302 + n, [n])
        assert np.array_equal(k4b.exhaustive(c, 25, True), oracle.exhaustive_brute(c, 25, True))


@pytest.mark.parametrize("lens,K,both,alpha", [([300000, 50000, 700], 25, True, 4), ([400000], 50, True, 4),
                                               ([250000, 250000], 100, False, 4), ([200000], 32, True, 5),
                                               ([150000, 100000], 20, True, 2)])
def test_diag_engine_equals_popc_engine_at_scale(lens, K, both, alpha):
    c = random_genome(900 + K, lens, alpha)
    # plant repeats so that small minima (and long runs of flagged cells) exist
    c[5000:9000] = c[100000:104000]
    c[20000:20300] = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)[c[150000:150300][::-1]]
    c = np.ascontiguousarray(c)
    got = k4b.exhaustive(c, K, both)
    k4b.set_engine(hamm.ENGINE_POPC)
    try:
        want = k4b.exhaustive(c, K, both)
    finally:
        k4b.set_engine(hamm.ENGINE_DIAG)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nparts", [1, 3])
def test_slab_ranges_and_parts_combine_to_the_full_result(nparts):
    """k4b_diag_slabs_device: slab by slab, part by part (separate minima arrays combined by an
    element-wise minimum after every slab, as the multi-GPU drivers do) == one all-slab call."""
    import torch
    c = random_genome(950, [700000, 90000])
    c[3000:5000] = c[400000:402000]
    c = np.ascontiguousarray(c)
    K, both, L = 32, True, len(c)
    want = k4b.exhaustive(c, K, both)
    g = hamm.Packed.from_host(c, K)
    try:
        n_slabs = hamm.diag_slab_count(g, both, nparts)
        assert n_slabs >= 4  # large enough for a real schedule
        bests = [torch.empty(L, dtype=torch.int32, device="cuda") for _ in range(nparts)]
        for i, b in enumerate(bests):
            hamm.best_init_device(b.data_ptr(), L, K)
            lo, hi = L * i // nparts, L * (i + 1) // nparts
            hamm.diag_bootstrap_device(g, both, lo, hi, b.data_ptr())
        def combine():
            m = bests[0]
            for b in bests[1:]:
                m = torch.minimum(m, b)
            for b in bests:
                b.copy_(m)
        torch.cuda.synchronize()
        combine()
        launches = 0
        for slab in range(n_slabs):
            for i, b in enumerate(bests):
                launches += hamm.diag_slabs_device(g, both, i, nparts, slab, slab + 1, b.data_ptr())
            torch.cuda.synchronize()
            combine()
        assert launches > 0 and hamm.last_kernel_ms() > 0
        out = torch.empty(L, dtype=torch.int16, device="cuda")
        hamm.best_finalize_device(g, bests[0].data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint16), want)
    finally:
        g.free()


def test_diag_engine_config1_output_md5(tmp_path):
    """BASELINE configs[0] through the CLI with the diagonal engine forced."""
    import random
    random.seed(12)
    s = "".join(random.choices("ACGT", k=1_000_000))
    fa = ">chr1 synthetic\n" + "".join(s[i:i + 80] + "\n" for i in range(0, len(s), 80))
    fa_path, seq, out = str(tmp_path / "g.fa"), str(tmp_path / "g.seq"), str(tmp_path / "out.csv")
    open(fa_path, "w").write(fa)
    hostlib.fasta_to_bioseq(fa_path, seq, "cfg1")
    env = dict(os.environ, K4B_ENGINE="diag")
    p = subprocess.run([hostlib.cli_path(), "hammings", "-m1", "-K25", "-c", "-i", seq, "-o", out], env=env,
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:]
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == "f1b2e85a8f64dc68dcc60cc2403cfb1a"


def _targeted_runs():
    m = golden_manifest()["__targeted__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_runs(), ids=lambda mr: mr[1]["out"])
def test_band_engine_targeted_equals_reference_files(oracle, mr):
    """-m0 -I on the band engine (rectangular mode, fixed threshold, wildcard rules)."""
    m, r = mr
    _, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    concat, chroms, _ = oracle.concat_entries(oracle.read_bioseq(os.path.join(GOLDEN, m["probes"][r["probes"]]["bioseq"])))
    got = k4b.targeted(tseq, concat, r["K"], r["R"], r["both"])
    rep = oracle.restricted_report(chroms, r["K"], r["R"], oracle.restricted_per_loci(chroms, got), r["fmt"], out_name=r["out"])
    assert rep == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def test_band_engine_targeted_matches_oracle_with_wildcards(oracle):
    rng = np.random.default_rng(501)
    target = random_genome(502, [6000, 4000]).copy()
    target[1000:1003] = 4
    target[2500] = 6
    probes = target[900:1500].copy()
    for p in (50, 130, 131, 200, 260, 261, 262, 263, 264, 400):
        probes[p] = 4
    probes[300] = 5
    probes = np.ascontiguousarray(np.concatenate([probes, [7], rng.integers(0, 5, size=400, dtype=np.uint8),
                                                  [7], oracle.CPL[target[7000:7300][::-1]]]), dtype=np.uint8)
    for K, R, both in [(32, 3, True), (25, 2, False), (64, 5, True), (140, 9, True), (20, 1, True)]:
        assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both)), (K, R)


def test_band_engine_targeted_equals_popc_engine_at_scale():
    rng = np.random.default_rng(77)
    target = random_genome(78, [1500000, 500000])
    parts = []
    for start, nmut in [(1000, 0), (200000, 3), (900000, 10), (1600000, 25)]:
        seg = target[start:start + 3000].copy()
        idx = rng.choice(3000, size=nmut * 10, replace=False)
        seg[idx] = (seg[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
        parts += [seg, np.array([7], np.uint8)]
    parts.append(rng.integers(0, 4, size=20000, dtype=np.uint8))
    probes = np.ascontiguousarray(np.concatenate(parts), dtype=np.uint8)
    got = k4b.targeted(target, probes, 32, 3, True)
    k4b.set_engine(hamm.ENGINE_POPC)
    try:
        want = k4b.targeted(target, probes, 32, 3, True)
    finally:
        k4b.set_engine(hamm.ENGINE_DIAG)
    assert np.array_equal(got, want)
    assert (got[:2900] == 0).all() and got[got != 0xFF].max() <= 4 and (got == 0xFF).sum() < 5 * 32


def test_window_table_path_equals_row_table_path(monkeypatch):
    """The two two-plane instances of diag_min_kernel (K4B_DIAG_EWIN=1: per-thread mismatch-window
    table in shared memory, the default; =0: per-warp row table + IMAD broadcast XOR) and both row
    segment lengths give identical minima, exhaustive and rectangular (targeted) mode."""
    c = random_genome(960, [350000, 60000, 900])
    c[7000:9500] = c[200000:202500]
    c[30000:30400] = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)[c[250000:250400][::-1]]
    c = np.ascontiguousarray(c)
    probes = np.ascontiguousarray(np.concatenate([c[1000:9000], [7], random_genome(961, [5000])]), dtype=np.uint8)
    probes[100] = (probes[100] + 1) % 4
    res = {}
    for ewin, dw in (("1", "1"), ("1", "2"), ("0", "1")):  # K4B_DIAG_DW=2: two diagonal words per thread
        for rows in ("8192", "4096"):
            monkeypatch.setenv("K4B_DIAG_EWIN", ewin)
            monkeypatch.setenv("K4B_DIAG_DW", dw)
            monkeypatch.setenv("K4B_DIAG_ROWS", rows)
            res[(ewin, dw, rows)] = (k4b.exhaustive(c, 40, True), k4b.targeted(c, probes, 32, 3, True))
    ref = res[("0", "1", "4096")]
    for key, (ex, tg) in res.items():
        assert np.array_equal(ex, ref[0]), key
        assert np.array_equal(tg, ref[1]), key


def _numpy_min_distance(concat, valid, K, q, both, self_pos):
    """Independent check: minimum distance of the K-mer q (codes 0..3) to every valid K-mer of
    concat (forward, excluding self_pos; reverse complement incl. self), K passes over the array."""
    M = len(concat) - K + 1
    best = K + 1
    for strand in range(2 if both else 1):
        kq = q if strand == 0 else (3 - q[::-1])
        acc = np.zeros(M, dtype=np.uint8)
        for p in range(K):
            acc += concat[p:p + M] != kq[p]
        acc = acc.astype(np.int32)
        acc[~valid[:M]] = K + 1
        if strand == 0:
            acc[self_pos] = K + 1
        best = min(best, int(acc.min()))
    return best


def test_band_engine_at_config2_size_against_popc_samples_and_brute_force():
    """BASELINE configs[1] size (10 Mbp multifasta, K=50, both strands): the band result is compared
    with the POPC engine on 64 x 256 sampled query K-mers (k4b_hamm_exhaustive_shard, a different
    kernel and formulation) and with a NumPy brute force on 6 of them."""
    import bench
    concat, chroms, K, both = bench.synth_genome("cfg2")
    got = k4b.exhaustive(concat, K, both)
    L = len(concat)
    rng = np.random.default_rng(5)
    starts = sorted(int(v) for v in rng.integers(0, L - 300, size=64))
    starts[0] = 0
    starts[-1] = L - 256                      # the tail: positions without a K-mer report K+1
    starts[10] = chroms[1][1] - 128           # across a chromosome boundary
    k4b.set_engine(hamm.ENGINE_POPC)
    try:
        for b in starts:
            want = np.full(L, K + 1, dtype=np.uint16)
            hamm.exhaustive_shard(concat, K, both, b, b + 256, want)
            assert np.array_equal(got[b:b + 256], want[b:b + 256]), b
    finally:
        k4b.set_engine(hamm.ENGINE_DIAG)
    valid = np.zeros(L, dtype=bool)
    for _, st, n in chroms:
        valid[st:st + max(0, n - K + 1)] = True
    for pos in (starts[3], starts[20] + 7, chroms[2][1] + 11, starts[40], chroms[4][1] + 5, starts[55] + 100):
        assert valid[pos]
        assert got[pos] == _numpy_min_distance(concat, valid, K, concat[pos:pos + K], both, pos), pos
