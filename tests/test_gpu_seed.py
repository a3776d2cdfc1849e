"""GPU parity tests of the seed-and-verify engine for the targeted mode (k4b_seed.cu): reference
golden files, the oracle on seeded inputs (short and hashed cores, K > 128, targets with N,
entry boundaries) and the band engine at a size the CPU oracle cannot reach."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_manifest, random_genome

import kit4b_b200 as k4b
from kit4b_b200 import hamm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _engine():
    k4b.gpu_init(1)
    k4b.set_engine(hamm.ENGINE_SEED)
    yield
    k4b.set_engine(hamm.ENGINE_AUTO)
    k4b.gpu_shutdown()


def _targeted_runs():
    m = golden_manifest()["__targeted__"]
    return [(m, r) for r in m["runs"]]


@pytest.mark.parametrize("mr", _targeted_runs(), ids=lambda mr: mr[1]["out"])
def test_seed_engine_targeted_equals_reference_files(oracle, mr):
    """-m0 -I: reference outputs (probes with N fall back to the brute-force engines)."""
    m, r = mr
    _, tseq = oracle.read_sfx(os.path.join(GOLDEN, m["sfx"]))
    concat, chroms, _ = oracle.concat_entries(oracle.read_bioseq(os.path.join(GOLDEN, m["probes"][r["probes"]]["bioseq"])))
    got = k4b.targeted(tseq, concat, r["K"], r["R"], r["both"])
    rep = oracle.restricted_report(chroms, r["K"], r["R"], oracle.restricted_per_loci(chroms, got), r["fmt"], out_name=r["out"])
    assert rep == open(os.path.join(GOLDEN, r["out"]), "rb").read()


def _planted(seed, tlens, alpha_t=4):
    """targets + pure-ACGT probes made of mutated copies (forward and reverse complement), a
    piece that spans an entry boundary, and unrelated sequence"""
    rng = np.random.default_rng(seed)
    cpl = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)
    target = random_genome(seed + 1, tlens, alpha_t).copy()
    acgt = np.flatnonzero(target < 4)
    parts = []
    for start, nmut, rc in [(300, 0, False), (900, 2, False), (1500, 5, True), (2100, 9, False), (2700, 1, True)]:
        seg = target[start:start + 400].copy()
        seg = seg[seg < 4]
        idx = rng.choice(len(seg), size=min(nmut * 3, len(seg)), replace=False)
        seg[idx] = (seg[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
        parts += [cpl[seg[::-1]] if rc else seg, np.array([7], np.uint8)]
    b = tlens[0]  # first entry boundary (an EOS sits at index tlens[0])
    span = target[b - 150:b + 150]
    parts += [span[span < 4], np.array([7], np.uint8), rng.integers(0, 4, size=500, dtype=np.uint8)]
    assert len(acgt)
    return np.ascontiguousarray(target), np.ascontiguousarray(np.concatenate(parts), dtype=np.uint8)


@pytest.mark.parametrize("K,R,both", [(32, 3, True), (25, 2, False), (48, 5, True), (64, 1, True), (100, 4, True),
                                      (140, 9, True), (300, 2, False), (20, 1, True), (24, 3, True), (500, 10, True)])
def test_seed_engine_matches_oracle(oracle, K, R, both):
    """cores of 6..250 bases: direct codes (<= 11), hashed buckets, multi-word cores"""
    target, probes = _planted(4000 + K, [6000, 3000, 2500])
    assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both))


@pytest.mark.parametrize("K,R,both", [(32, 3, True), (25, 2, False), (100, 4, True), (300, 2, False)])
def test_seed_engine_bucket_major_join_matches_oracle(oracle, monkeypatch, K, R, both):
    """the join schedule (items sorted by bucket, buckets staged in shared memory) is chosen for
    long buckets only; K4B_SEED_JOIN=1 forces it on these small inputs (the library reads the
    variable at every launch)"""
    monkeypatch.setenv("K4B_SEED_JOIN", "1")
    target, probes = _planted(4600 + K, [6000, 3000, 2500])
    assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both))
    assert np.array_equal(k4b.targeted(target, None, K, R, both, intra_inter_both=2),
                          oracle.targeted_self_brute(target, K, R, both, 2))


@pytest.mark.parametrize("phases", ["0", "1"])
@pytest.mark.parametrize("variant", ["0", "1", "2"])
def test_seed_join_shapes_and_phase_schedule(oracle, monkeypatch, variant, phases):
    """every shape of the join kernel (items per CTA, tile size: K4B_SEED_JOIN_VARIANT) with and without the
    core-by-core phase schedule
    (K4B_SEED_PHASES) equals the oracle - probes with hits at every distance, both strands, and the
    assembly against itself with the -z filter"""
    monkeypatch.setenv("K4B_SEED_JOIN", "1")
    monkeypatch.setenv("K4B_SEED_JOIN_VARIANT", variant)
    monkeypatch.setenv("K4B_SEED_PHASES", phases)
    for K, R, both in ((32, 3, True), (25, 2, False), (100, 4, True)):
        target, probes = _planted(5600 + K, [6000, 3000, 2500])
        assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both)), (K, R)
    target, _ = _planted(5700, [5000, 2500])
    assert np.array_equal(k4b.targeted(target, None, 32, 3, True, intra_inter_both=1),
                          oracle.targeted_self_brute(target, 32, 3, True, 1))


def test_seed_engine_targets_with_non_acgt(oracle):
    target, probes = _planted(4100, [5000, 4000], alpha_t=5)  # N in the targets, probes stay ACGT
    target[1000:1010] = 4
    target[2200] = 6
    for K, R, both in [(32, 3, True), (40, 2, True), (70, 4, False)]:
        assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both)), (K, R)


def test_seed_engine_probe_sets_with_wildcards(oracle):
    """probe K-mers holding N / InDel are skipped by the seed kernel and answered by the POPC
    engine on their position intervals (wildcard rule, > 4 wildcards -> 0)"""
    target, probes = _planted(4300, [6000, 3000])
    probes = probes.copy()
    for p in (50, 130, 131, 700, 701, 702, 703, 704, 1200, len(probes) - 5):
        if probes[p] != 7:
            probes[p] = 4
    probes[900] = 6
    probes = np.ascontiguousarray(probes)
    for K, R, both in [(32, 3, True), (25, 2, False), (64, 5, True)]:
        assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both)), (K, R)


@pytest.mark.parametrize("K,R,both,z", [(32, 3, True, 0), (25, 2, False, 0), (32, 3, True, 1), (32, 3, True, 2),
                                         (25, 2, True, 1), (140, 9, True, 2), (20, 1, False, 2)])
def test_seed_engine_self_mode_matches_oracle(oracle, K, R, both, z):
    """-m0 without -I: probes are the assembly's own K-mers - exact sense self hit skipped, -z
    filter of exact sense hits, cap at 20, wildcard K-mers through the POPC engine"""
    rng = np.random.default_rng(4400 + K)
    cpl = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)
    a = rng.integers(0, 4, size=3000, dtype=np.uint8)
    e1 = np.concatenate([a[:1500], a[200:500], a[1500:], rng.integers(0, 4, size=300, dtype=np.uint8)])  # intra duplicate
    e2 = np.concatenate([rng.integers(0, 4, size=400, dtype=np.uint8), a[1000:1400], cpl[a[2000:2300][::-1]],
                         rng.integers(0, 4, size=200, dtype=np.uint8)])                                    # inter + antisense
    e2[50:53] = 4                                                                                          # wildcards
    e3 = a[2500:2900].copy()
    e3[[10, 200]] = (e3[[10, 200]] + 1) % 4
    target = np.ascontiguousarray(np.concatenate([e1, [7], e2, [7], e3, [7]]), dtype=np.uint8)
    want = oracle.targeted_self_brute(target, K, R, both, z)
    assert np.array_equal(k4b.targeted(target, None, K, R, both, intra_inter_both=z), want)
    k4b.set_engine(hamm.ENGINE_POPC)
    try:
        assert np.array_equal(k4b.targeted(target, None, K, R, both, intra_inter_both=z), want)
    finally:
        k4b.set_engine(hamm.ENGINE_SEED)


def test_seed_engine_self_mode_at_scale_equals_popc_engine():
    rng = np.random.default_rng(97)
    t = random_genome(98, [300000, 120000]).copy()
    t[5000:8000] = t[200000:203000]
    t[310000:311000] = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)[t[40000:41000][::-1]]
    t[100000:100005] = 4
    t = np.ascontiguousarray(t)
    for z in (0, 2):
        got = k4b.targeted(t, None, 32, 3, True, intra_inter_both=z)
        k4b.set_engine(hamm.ENGINE_POPC)
        try:
            want = k4b.targeted(t, None, 32, 3, True, intra_inter_both=z)
        finally:
            k4b.set_engine(hamm.ENGINE_SEED)
        assert np.array_equal(got, want), z
        # the planted copy lies in the same entry: an exact hit, unless -z2 (inter only) filters it
        assert (got[5000:7900] == 0).all() == (z == 0)


def test_seed_engine_device_api_ranges_combine(oracle):
    """k4b_targeted_seed_device on probe sub-ranges into one minima array == the host entry point"""
    import torch
    target, probes = _planted(4200, [7000, 2000])
    K, R, both = 32, 3, True
    core = K // (R + 1)
    clamp = K // core
    t = hamm.Packed.from_host(target, K)
    q = hamm.Packed.from_host(probes, K)
    try:
        L = len(probes)
        best = torch.empty(L, dtype=torch.int32, device="cuda")
        hamm.best_init_device(best.data_ptr(), L, K)
        for b, e in [(0, 700), (700, 701), (701, L)]:
            assert hamm.targeted_seed_device(q, t, both, clamp, core, b, e, best.data_ptr()) > 0
        out = torch.empty(L, dtype=torch.int16, device="cuda")
        hamm.targeted_finalize_device(q, best.data_ptr(), clamp, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint16)
        want = oracle.targeted_brute(target, probes, K, R, both)
        valid = want != 0xFF
        assert np.array_equal(got[valid].astype(np.uint8), want[valid])
        with pytest.raises(Exception):  # the pigeonhole bound needs clamp <= K / core_len
            hamm.targeted_seed_device(q, t, both, clamp + 1, core, 0, L, best.data_ptr())
    finally:
        t.free()
        q.free()


def test_seed_engine_equals_band_engine_at_scale():
    rng = np.random.default_rng(87)
    target = random_genome(88, [3000000, 1000000])
    parts = []
    for start, nmut in [(1000, 0), (200000, 3), (900000, 10), (3100000, 25), (2500000, 40)]:
        seg = target[start:start + 3000].copy()
        idx = rng.choice(3000, size=nmut * 10, replace=False)
        seg[idx] = (seg[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
        parts += [seg, np.array([7], np.uint8)]
    parts.append(rng.integers(0, 4, size=30000, dtype=np.uint8))
    probes = np.ascontiguousarray(np.concatenate(parts), dtype=np.uint8)
    for K, R in [(32, 3), (50, 5), (96, 2)]:
        got = k4b.targeted(target, probes, K, R, True)
        k4b.set_engine(hamm.ENGINE_DIAG)
        try:
            want = k4b.targeted(target, probes, K, R, True)
        finally:
            k4b.set_engine(hamm.ENGINE_SEED)
        assert np.array_equal(got, want), (K, R)
        assert (got[:2900] == 0).all()


def test_targeted_distributed_single_rank_on_cuda(oracle):
    """kit4b_b200.dist.targeted_distributed with the real CudaEngine (one rank; the two-rank
    collective flow is covered on CPU in tests/test_dist_gloo.py)."""
    import socket
    import torch.distributed as dist
    from kit4b_b200.dist import targeted_distributed
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        target, probes = _planted(4500, [6000, 3000])
        for K, R, both in [(32, 3, True), (50, 5, False)]:
            assert np.array_equal(targeted_distributed(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("join", ["0", "1"])
@pytest.mark.parametrize("K,R,both,nparts", [(32, 3, True, 3), (100, 4, True, 2), (25, 2, False, 8)])
def test_seed_engine_bucket_parts_combine_to_the_full_result(oracle, monkeypatch, K, R, both, nparts, join):
    """k4b_targeted_seed_part_device: every part indexes only its share of the core buckets and
    answers all probes for it; the element-wise minimum over the parts (what the multi-GPU drivers
    all_reduce) equals the oracle, in both schedules."""
    import torch
    monkeypatch.setenv("K4B_SEED_JOIN", join)
    target, probes = _planted(5200 + K, [7000, 2500, 3000])
    want = oracle.targeted_brute(target, probes, K, R, both)
    core = K // (R + 1)
    clamp = K // core
    t, q = hamm.Packed.from_host(target, K), hamm.Packed.from_host(probes, K)
    try:
        L = len(probes)
        total = torch.empty(L, dtype=torch.int32, device="cuda")
        hamm.best_init_device(total.data_ptr(), L, K)
        cores_seen = 0
        for part in range(nparts):
            best = torch.empty(L, dtype=torch.int32, device="cuda")
            hamm.best_init_device(best.data_ptr(), L, K)
            assert hamm.targeted_seed_part_device(q, t, both, clamp, core, part, nparts, best.data_ptr()) > 0
            cores_seen += hamm.last_seed_info()["indexed_cores"]
            total = torch.minimum(total, best)
        full = torch.empty(L, dtype=torch.int32, device="cuda")
        hamm.best_init_device(full.data_ptr(), L, K)
        hamm.targeted_seed_device(q, t, both, clamp, core, 0, L, full.data_ptr())
        assert cores_seen == hamm.last_seed_info()["indexed_cores"]  # the parts partition the index
        out = torch.empty(L, dtype=torch.int16, device="cuda")
        hamm.targeted_finalize_device(q, total.data_ptr(), clamp, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint16)
        res = np.full(L, 0xFF, dtype=np.uint8)
        res[got <= K] = got[got <= K].astype(np.uint8)
        assert np.array_equal(res, want)
    finally:
        t.free()
        q.free()


@pytest.mark.parametrize("join", ["0", "1"])
@pytest.mark.parametrize("build", ["1", "2", "14", "29"])
def test_seed_index_by_partition_passes_matches_oracle(oracle, monkeypatch, join, build):
    """K4B_SEED_INDEX != 0: the index built by two partition passes (flags, k4b_kernels.cuh: 1 plain, 2 tiles of 4096
    entries, 4 count pass with shared-memory counters, 8 three-word field extraction, 16 persistent second pass;
    cores of 6, 7 and 8 bases: buckets of 12, 14 and 16 bits; longer cores keep the one-by-one build) answers like
    the oracle in both query schedules: planted probes, targets with N, the assembly against itself with -z, and
    bucket shards."""
    import torch
    monkeypatch.setenv("K4B_SEED_INDEX", build)
    monkeypatch.setenv("K4B_SEED_JOIN", join)
    for K, R, both in ((32, 3, True), (25, 3, False), (28, 3, True), (48, 5, True), (20, 1, True), (64, 7, True)):
        target, probes = _planted(6100 + K, [6000, 3000, 2500])
        assert np.array_equal(k4b.targeted(target, probes, K, R, both), oracle.targeted_brute(target, probes, K, R, both)), (K, R)
    target, probes = _planted(6200, [5000, 4000], alpha_t=5)
    target[1000:1010] = 4
    assert np.array_equal(k4b.targeted(target, probes, 32, 3, True), oracle.targeted_brute(target, probes, 32, 3, True))
    target, _ = _planted(6300, [5000, 2500])
    assert np.array_equal(k4b.targeted(target, None, 32, 3, True, intra_inter_both=1),
                          oracle.targeted_self_brute(target, 32, 3, True, 1))
    # bucket shards: three parts of the index, minima combined
    K, R, both, nparts = 32, 3, True, 3
    target, probes = _planted(6400, [7000, 2500, 3000])
    want = oracle.targeted_brute(target, probes, K, R, both)
    t, q = hamm.Packed.from_host(target, K), hamm.Packed.from_host(probes, K)
    try:
        L = len(probes)
        total = torch.empty(L, dtype=torch.int32, device="cuda")
        hamm.best_init_device(total.data_ptr(), L, K)
        for part in range(nparts):
            best = torch.empty(L, dtype=torch.int32, device="cuda")
            hamm.best_init_device(best.data_ptr(), L, K)
            assert hamm.targeted_seed_part_device(q, t, both, 4, 8, part, nparts, best.data_ptr()) > 0
            total = torch.minimum(total, best)
        out = torch.empty(L, dtype=torch.int16, device="cuda")
        hamm.targeted_finalize_device(q, total.data_ptr(), 4, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint16)
        res = np.full(L, 0xFF, dtype=np.uint8)
        res[got <= K] = got[got <= K].astype(np.uint8)
        assert np.array_equal(res, want)
    finally:
        t.free()
        q.free()


def test_seed_index_partition_equals_one_by_one_at_3Mbp(monkeypatch):
    """all index builds give the same minima and index the same number of cores on a 3 Mbp assembly whose coarse
    partitions span several tiles, with a 150 kbp poly-A run and a 60 kbp (AC)n run (one bucket of 150 k entries:
    a digit that fills whole tiles) and an entry shorter than a tile"""
    rng = np.random.default_rng(6500)
    ents = [rng.integers(0, 4, size=n, dtype=np.uint8) for n in (1_700_000, 900, 1_300_000)]
    ents[0][400_000:550_000] = 0
    ents[2][100_000:160_000] = np.tile(np.array([0, 1], np.uint8), 30_000)
    ents[2][700_000:700_020] = 4
    target = np.ascontiguousarray(np.concatenate([np.concatenate([e, [7]]) for e in ents]).astype(np.uint8))
    cpl = np.array([3, 2, 1, 0, 4, 5, 6, 7], np.uint8)
    parts = []
    for start, n, rc in ((399_900, 400, False), (1_000_000, 3000, False), (2_000_000, 3000, True), (1_800_950, 200, False)):
        seg = target[start:start + n].copy()
        seg = seg[seg < 4]
        idx = rng.choice(len(seg), size=len(seg) // 30, replace=False)
        seg[idx] = (seg[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
        parts += [cpl[seg[::-1]] if rc else seg, np.array([7], np.uint8)]
    parts.append(rng.integers(0, 4, size=2000, dtype=np.uint8))
    probes = np.ascontiguousarray(np.concatenate(parts), dtype=np.uint8)
    res = {}
    for K, R in ((32, 3), (24, 3), (28, 3)):
        for join in ("0", "1"):
            for build in ("0", "1", "2", "5", "9", "17", "14", "29"):
                monkeypatch.setenv("K4B_SEED_JOIN", join)
                monkeypatch.setenv("K4B_SEED_INDEX", build)
                res[build] = (k4b.targeted(target, probes, K, R, True), hamm.last_seed_info()["indexed_cores"])
                assert res[build][1] == res["0"][1] > 2_900_000, (K, R, join, build)
                assert np.array_equal(res["0"][0], res[build][0]), (K, R, join, build)
            assert (res["1"][0] < K // (K // (R + 1))).sum() > 4000  # the planted copies are found


def test_seed_index_partition_with_one_huge_bucket(monkeypatch):
    """8 Mbp of poly-A beside 1 Mbp of random sequence: one bucket with 8 M entries - a shared-memory counter that
    hands on 2^15 several times, thousands of tiles with a single digit in both passes - gives the same index size
    and the same minima in every build"""
    rng = np.random.default_rng(6600)
    rnd = rng.integers(0, 4, size=1_000_000, dtype=np.uint8)
    target = np.ascontiguousarray(np.concatenate([rnd[:400_000], [7], np.zeros(8_000_000, np.uint8), [7], rnd[400_000:], [7]]), dtype=np.uint8)
    seg = rnd[700_000:703_000].copy()
    idx = rng.choice(len(seg), size=100, replace=False)
    seg[idx] = (seg[idx] + 1 + rng.integers(0, 3, size=len(idx))) % 4
    tail = np.concatenate([rnd[399_990:400_000], np.zeros(40, np.uint8)])  # no K-mer of the assembly: crosses an entry end
    probes = np.ascontiguousarray(np.concatenate([seg, [7], tail, [7], rng.integers(0, 4, size=1000, dtype=np.uint8)]), dtype=np.uint8)
    res = {}
    monkeypatch.setenv("K4B_SEED_JOIN", "1")
    for build in ("0", "1", "5", "29", "14"):
        monkeypatch.setenv("K4B_SEED_INDEX", build)
        res[build] = (k4b.targeted(target, probes, 32, 3, True), hamm.last_seed_info()["indexed_cores"])
        assert res[build][1] == res["0"][1] > 8_900_000, build
        assert np.array_equal(res["0"][0], res[build][0]), build
    assert (res["0"][0][:2900] < 4).sum() > 2500


def test_exact_where_the_reference_depth_cut_fires(oracle, tmp_path):
    """Repeat-rich assembly (tests/golden/depth_case.py): the seed engine reports the true minimum where the
    reference at its default sensitivity truncates its search and reports "not found"; the depth signal
    (k4b_last_depth_cut, logged by the CLI) counts exactly that K-mer for -s0 / -s3 and none for -s1 / -s2."""
    import subprocess
    import sys
    sys.path.insert(0, GOLDEN)
    import bench
    import depth_case
    from kit4b_b200 import hostlib
    asm, probes, upos = depth_case.build()
    target = np.ascontiguousarray(np.concatenate([np.concatenate([c, [7]]) for _, c in asm]).astype(np.uint8))
    concat, chroms, _ = oracle.concat_entries(probes)
    K, R = depth_case.K, depth_case.R
    want = oracle.targeted_brute(target, concat, K, R, True)
    gold = {s: open(os.path.join(GOLDEN, "depth.K32r3c.s%d.csv" % s), "rb").read() for s in range(4)}
    caps = {0: 40000, 1: 80000, 2: 200000, 3: 20000}  # 4 x MaxCoreDepth (hammings.cpp:2366-2386, SfxArray.cpp:4303)
    for join in ("0", "1"):
        os.environ["K4B_SEED_JOIN"] = join
        try:
            for s in range(4):
                hamm.set_reference_sensitivity(s)
                got = k4b.targeted(target, concat, K, R, True)
                assert np.array_equal(got, want) and got[upos] == 3
                cut = hamm.last_depth_cut()
                assert cut["max_copies"] == caps[s]
                assert cut["probe_kmers"] == (1 if s in (0, 3) else 0), (join, s, cut)
                rep = oracle.restricted_report(chroms, K, R, oracle.restricted_per_loci(chroms, got), 0)
                assert (rep == gold[s]) == (s in (1, 2))  # differs from the reference exactly where its cut fires
        finally:
            os.environ.pop("K4B_SEED_JOIN", None)
            hamm.set_reference_sensitivity(0)
    # the CLI: FASTA in, the -s1 golden out, and the log names the K-mer for -s0
    fa, pfa = str(tmp_path / "asm.fa"), str(tmp_path / "probes.fa")
    bench.write_fasta(fa, asm)
    bench.write_fasta(pfa, probes)
    for s, expect in ((0, "1 probe K-mers answered below the not-found value"), (1, "depth cut (-s1) cannot fire")):
        out = str(tmp_path / "o.csv")
        p = subprocess.run([hostlib.cli_path(), "hammings", "-m0", "-K32", "-r3", "-c", "-s%d" % s, "-i", fa, "-I", pfa, "-o", out],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stdout[-1500:]
        assert open(out, "rb").read() == gold[1]
        assert expect in p.stdout, p.stdout[-1500:]
