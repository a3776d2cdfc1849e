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
