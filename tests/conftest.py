import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_cases():
    """[(case name, run dict)] for parametrisation."""
    out = []
    for name, case in sorted(golden_manifest().items()):
        if name.startswith("__"):  # targeted / merge sections have their own layout
            continue
        for r in case["runs"]:
            out.append((name, r))
    return out


def case_id(p):
    name, r = p
    return "%s-K%d%s" % (name, r["K"], "c" if r["both"] else "w")


@pytest.fixture(scope="session")
def oracle():
    from oracle import hamm_oracle as ho
    ho.build()
    return ho


def load_case(ho, name):
    case = golden_manifest()[name]
    entries = ho.read_bioseq(os.path.join(GOLDEN, case["bioseq"]))
    return ho.concat_entries(entries)  # concat, chroms, genome_len


def random_genome(seed, lens, alphabet=4):
    """Seeded synthetic concat (codes 0..alphabet-1) with EOS between chromosomes."""
    rng = np.random.default_rng(seed)
    parts = []
    for i, n in enumerate(lens):
        parts.append(rng.integers(0, alphabet, size=n, dtype=np.uint8))
        if i + 1 < len(lens):
            parts.append(np.array([7], dtype=np.uint8))
    return np.ascontiguousarray(np.concatenate(parts))
