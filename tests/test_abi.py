"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/k4b_hamm.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import kit4b_b200 as k4b
from kit4b_b200 import hamm


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "k4b_hamm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(k4b_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("k4b_gpu_init", "k4b_hamm_exhaustive", "k4b_hamm_exhaustive_shard", "k4b_hamm_targeted",
                 "k4b_gpu_shutdown", "k4b_last_error", "k4b_allpairs_min_device", "k4b_pack_device"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(k4b.lib_path()), "build the library first (__graft_entry__.build())"
    L = ctypes.CDLL(k4b.lib_path())
    for name in declared_symbols():
        assert hasattr(L, name), name
    # and the python mirror binds exactly the declared set
    assert sorted(hamm.SIGNATURES) == declared_symbols()


def test_no_torch_or_oracle_in_the_boundary():
    text = open(os.path.join(ROOT, "include", "k4b_hamm.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "at::" not in code and "Tensor" not in code
    for root, _, files in os.walk(os.path.join(ROOT, "kit4b_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "hamm_oracle" not in src and "oracle/" not in src, os.path.join(root, f)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu():
    with pytest.raises(k4b.K4BError) as e:
        k4b.exhaustive(np.zeros(64, np.uint8), 25, True)
    assert e.value.code == -1000
    assert "no CPU fallback" in str(e.value)


def test_parameter_validation_precedes_device_use():
    L = k4b.load_lib()
    out = np.zeros(8, np.uint16)
    c = np.zeros(8, np.uint8)
    rc = L.k4b_hamm_exhaustive(c.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), 8, 9, 1, 1, 0,
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    assert rc == -100  # K below 10: eBSFerrParams
    rc = L.k4b_hamm_exhaustive(None, 8, 25, 1, 1, 0, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    assert rc == -100
