"""ctypes mirror of include/k4b_hamm.h.

Argument meaning and error behaviour follow the reference's exhaustive / restricted engines at
the seam described in include/k4b_hamm.h (ngskit4b/hammings.cpp:2740-2867 and :1691-1694):
results come back in arrays laid out exactly like the reference's m_pHamDist / m_pRHammings.
No CPU path exists here: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libk4bhamm.so"


class K4BError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("k4b error %d: %s" % (code, msg))
        self.code = code


def lib_path() -> str:
    return os.path.join(_HERE, _LIB_NAME)


_lib = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol declared in include/k4b_hamm.h
SIGNATURES = {
    "k4b_gpu_init": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "k4b_gpu_shutdown": (None, []),
    "k4b_gpu_count": (ctypes.c_int, []),
    "k4b_last_error": (ctypes.c_char_p, []),
    "k4b_hamm_exhaustive": (ctypes.c_int, [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                           ctypes.c_uint32, _u16p]),
    "k4b_hamm_exhaustive_shard": (ctypes.c_int, [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                                 ctypes.c_uint32, ctypes.c_uint32, _u16p]),
    "k4b_hamm_targeted": (ctypes.c_int, [_u8p, ctypes.c_uint64, _u8p, ctypes.c_uint32, ctypes.c_uint32,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _u8p]),
    "k4b_hamm_targeted_z": (ctypes.c_int, [_u8p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _u8p]),
    "k4b_packed_image_bytes": (ctypes.c_size_t, [ctypes.c_uint32]),
    "k4b_pack_host": (ctypes.c_int, [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(_vp)]),
    "k4b_pack_device": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp, ctypes.POINTER(_vp)]),
    "k4b_pack_device_into": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp, ctypes.c_size_t, _vp,
                                            ctypes.POINTER(_vp)]),
    "k4b_packed_image_ptr": (_vp, [_vp]),
    "k4b_packed_image_size": (ctypes.c_size_t, [_vp]),
    "k4b_packed_from_image": (ctypes.c_int, [_vp, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                             ctypes.POINTER(_vp)]),
    "k4b_packed_has_non_acgt": (ctypes.c_int, [_vp]),
    "k4b_packed_num_kmers": (ctypes.c_uint64, [_vp]),
    "k4b_packed_free": (None, [_vp]),
    "k4b_allpairs_min_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                               ctypes.c_uint32, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "k4b_last_kernel_ms": (ctypes.c_float, []),
    "k4b_set_engine": (ctypes.c_int, [ctypes.c_int]),
    "k4b_get_engine": (ctypes.c_int, []),
    "k4b_best_init_device": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    "k4b_exhaustive_diag_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp,
                                                  ctypes.POINTER(ctypes.c_int)]),
    "k4b_best_finalize_device": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "k4b_last_diag_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint32)] * 4),
    "k4b_diag_slab_count": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]),
    "k4b_diag_slabs_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                             ctypes.c_uint32, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "k4b_targeted_diag_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_uint32, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "k4b_targeted_seed_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                                ctypes.c_uint32, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "k4b_targeted_seed_part_device": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                                                     ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, ctypes.POINTER(ctypes.c_int)]),
    "k4b_last_seed_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64)] * 2),
    "k4b_set_reference_sensitivity": (ctypes.c_int, [ctypes.c_int]),
    "k4b_last_depth_cut": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]),
    "k4b_seed_watch_depth": (ctypes.c_int, [ctypes.c_uint32, _vp]),
    "k4b_hamm_histogram": (ctypes.c_int, [_u16p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64)]),
    "k4b_histogram_device": (ctypes.c_int, [_vp, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    "k4b_targeted_finalize_device": (ctypes.c_int, [_vp, _vp, ctypes.c_uint32, _vp, _vp]),
    "k4b_diag_bootstrap_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp]),
    "k4b_diag_bands_device": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _vp, _vp,
                                             ctypes.POINTER(ctypes.c_int)]),
    "k4b_microbench_intpipe": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
}


def load_lib():
    """Loads the CUDA engine.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise K4BError(-1000, "%s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "or `make -C kit4b_b200/csrc`; kit4b_b200 has no CPU fallback" % p)
        L = ctypes.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise K4BError(rc, (load_lib().k4b_last_error() or b"").decode("utf-8", "replace"))


def _as_u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def gpu_init(n_gpus: int = 1, device_ids=None) -> None:
    ids = None
    if device_ids is not None:
        ids = (ctypes.c_int * len(device_ids))(*device_ids)
    _check(load_lib().k4b_gpu_init(n_gpus, ids))


def gpu_shutdown() -> None:
    load_lib().k4b_gpu_shutdown()


def gpu_count() -> int:
    return load_lib().k4b_gpu_count()


def exhaustive(concat, K: int, both_strands: bool, sweep_start: int = 1, sweep_end: int = 0) -> np.ndarray:
    """All-vs-all minimum Hamming distances (-m1).  Returns uint16[len(concat)] laid out like the
    reference's m_pHamDist: K+1 wherever no K-mer starts (hammings.cpp:3120-3122)."""
    c = _as_u8(concat)
    out = np.full(len(c), K + 1, dtype=np.uint16)
    _check(load_lib().k4b_hamm_exhaustive(c.ctypes.data_as(_u8p), len(c), K, int(both_strands), sweep_start,
                                          sweep_end, out.ctypes.data_as(_u16p)))
    return out


def exhaustive_shard(concat, K: int, both_strands: bool, q_begin: int, q_end: int,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
    c = _as_u8(concat)
    if out is None:
        out = np.full(len(c), K + 1, dtype=np.uint16)
    _check(load_lib().k4b_hamm_exhaustive_shard(c.ctypes.data_as(_u8p), len(c), K, int(both_strands), q_begin, q_end,
                                                out.ctypes.data_as(_u16p)))
    return out


def targeted(target_concat, probe_concat, K: int, R: int, both_strands: bool, q_begin: int = 0,
             q_end: int = 0, intra_inter_both: int = 0) -> np.ndarray:
    """Probe K-mers vs assembly (-m0 -I).  Returns uint8[len(probe_concat)], 0xFF where no K-mer starts.
    probe_concat None = the probes are the assembly's own K-mers; only then does intra_inter_both
    (-z: 1 intra only, 2 inter only, applied to exact sense hits as in the reference) take effect."""
    t = _as_u8(target_concat)
    if probe_concat is None:  # probes are the target's own K-mers (no -I)
        out = np.full(len(t), 0xFF, dtype=np.uint8)
        _check(load_lib().k4b_hamm_targeted_z(t.ctypes.data_as(_u8p), len(t), K, R, int(both_strands),
                                              int(intra_inter_both), q_begin, q_end, out.ctypes.data_as(_u8p)))
        return out
    p = _as_u8(probe_concat)
    out = np.full(len(p), 0xFF, dtype=np.uint8)
    _check(load_lib().k4b_hamm_targeted(t.ctypes.data_as(_u8p), len(t), p.ctypes.data_as(_u8p), len(p), K, R,
                                        int(both_strands), q_begin, q_end, out.ctypes.data_as(_u8p)))
    return out


class Packed:
    """Device-resident bit-plane image of a concatenated sequence set."""

    def __init__(self, handle, length: int, K: int):
        self.handle = handle
        self.length = length
        self.K = K

    @classmethod
    def from_host(cls, concat, K: int) -> "Packed":
        c = _as_u8(concat)
        h = _vp()
        _check(load_lib().k4b_pack_host(c.ctypes.data_as(_u8p), len(c), K, ctypes.byref(h)))
        return cls(h, len(c), K)

    @classmethod
    def from_device(cls, d_ptr: int, length: int, K: int, stream: int = 0) -> "Packed":
        h = _vp()
        _check(load_lib().k4b_pack_device(_vp(d_ptr), length, K, _vp(stream), ctypes.byref(h)))
        return cls(h, length, K)

    @classmethod
    def from_device_into(cls, d_ptr: int, length: int, K: int, d_image_ptr: int, image_bytes: int,
                         stream: int = 0) -> "Packed":
        h = _vp()
        _check(load_lib().k4b_pack_device_into(_vp(d_ptr), length, K, _vp(d_image_ptr), image_bytes, _vp(stream),
                                               ctypes.byref(h)))
        return cls(h, length, K)

    @classmethod
    def from_image(cls, d_image_ptr: int, image_bytes: int, length: int, K: int, has_non_acgt: bool) -> "Packed":
        h = _vp()
        _check(load_lib().k4b_packed_from_image(_vp(d_image_ptr), image_bytes, length, K, int(has_non_acgt),
                                                ctypes.byref(h)))
        return cls(h, length, K)

    @property
    def image_ptr(self) -> int:
        return load_lib().k4b_packed_image_ptr(self.handle)

    @property
    def image_size(self) -> int:
        return load_lib().k4b_packed_image_size(self.handle)

    @property
    def has_non_acgt(self) -> bool:
        return bool(load_lib().k4b_packed_has_non_acgt(self.handle))

    @property
    def num_kmers(self) -> int:
        return load_lib().k4b_packed_num_kmers(self.handle)

    def free(self):
        if self.handle:
            load_lib().k4b_packed_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def allpairs_min_device(queries: Packed, targets: Packed, both_strands: bool, self_exclude: bool, q_begin: int,
                        q_end: int, d_out_ptr: int, clamp: int = 0, stream: int = 0) -> int:
    """Enqueues the all-pairs kernel on device-resident data; returns the number of kernel launches."""
    n = ctypes.c_int(0)
    _check(load_lib().k4b_allpairs_min_device(queries.handle, targets.handle, int(both_strands), int(self_exclude),
                                              q_begin, q_end, clamp, _vp(d_out_ptr), _vp(stream), ctypes.byref(n)))
    return n.value


ENGINE_AUTO, ENGINE_POPC, ENGINE_DIAG, ENGINE_SEED = 0, 1, 2, 3


def set_engine(engine: int) -> None:
    """0 automatic, 1 POPC all-pairs, 2 diagonal bands (exhaustive full sweeps only)."""
    _check(load_lib().k4b_set_engine(engine))


def get_engine() -> int:
    return load_lib().k4b_get_engine()


def best_init_device(d_best_ptr: int, n: int, K: int, stream: int = 0) -> None:
    _check(load_lib().k4b_best_init_device(_vp(d_best_ptr), n, K, _vp(stream)))


def exhaustive_diag_device(g: Packed, both_strands: bool, part: int, nparts: int, d_best_ptr: int,
                           stream: int = 0) -> int:
    n = ctypes.c_int(0)
    _check(load_lib().k4b_exhaustive_diag_device(g.handle, int(both_strands), part, nparts, _vp(d_best_ptr),
                                                 _vp(stream), ctypes.byref(n)))
    return n.value


def diag_bootstrap_device(g: Packed, both_strands: bool, q_begin: int, q_end: int, d_best_ptr: int,
                          stream: int = 0) -> None:
    _check(load_lib().k4b_diag_bootstrap_device(g.handle, int(both_strands), q_begin, q_end, _vp(d_best_ptr),
                                                _vp(stream)))


def diag_bands_device(g: Packed, both_strands: bool, part: int, nparts: int, d_best_ptr: int, stream: int = 0) -> int:
    n = ctypes.c_int(0)
    _check(load_lib().k4b_diag_bands_device(g.handle, int(both_strands), part, nparts, _vp(d_best_ptr), _vp(stream),
                                            ctypes.byref(n)))
    return n.value


def diag_slab_count(g: Packed, both_strands: bool, nparts: int) -> int:
    n = ctypes.c_uint32(0)
    _check(load_lib().k4b_diag_slab_count(g.handle, int(both_strands), nparts, ctypes.byref(n)))
    return n.value


def diag_slabs_device(g: Packed, both_strands: bool, part: int, nparts: int, slab_begin: int, slab_end: int,
                      d_best_ptr: int, stream: int = 0) -> int:
    n = ctypes.c_int(0)
    _check(load_lib().k4b_diag_slabs_device(g.handle, int(both_strands), part, nparts, slab_begin, slab_end,
                                            _vp(d_best_ptr), _vp(stream), ctypes.byref(n)))
    return n.value


def targeted_diag_device(probes: Packed, targets: Packed, both_strands: bool, clamp: int, part: int, nparts: int,
                         d_best_ptr: int, stream: int = 0) -> int:
    n = ctypes.c_int(0)
    _check(load_lib().k4b_targeted_diag_device(probes.handle, targets.handle, int(both_strands), clamp, part, nparts,
                                               _vp(d_best_ptr), _vp(stream), ctypes.byref(n)))
    return n.value


def targeted_seed_device(probes: Packed, targets: Packed, both_strands: bool, clamp: int, core_len: int, q_begin: int,
                         q_end: int, d_best_ptr: int, stream: int = 0) -> int:
    n = ctypes.c_int(0)
    _check(load_lib().k4b_targeted_seed_device(probes.handle, targets.handle, int(both_strands), clamp, core_len,
                                               q_begin, q_end, _vp(d_best_ptr), _vp(stream), ctypes.byref(n)))
    return n.value


def targeted_seed_part_device(probes: Packed, targets: Packed, both_strands: bool, clamp: int, core_len: int, part: int,
                              nparts: int, d_best_ptr: int, stream: int = 0) -> int:
    """Seed engine on part `part` of `nparts` of the index BUCKETS against every probe K-mer."""
    n = ctypes.c_int(0)
    _check(load_lib().k4b_targeted_seed_part_device(probes.handle, targets.handle, int(both_strands), clamp, core_len,
                                                    part, nparts, _vp(d_best_ptr), _vp(stream), ctypes.byref(n)))
    return n.value


def last_seed_info() -> dict:
    a, b = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _check(load_lib().k4b_last_seed_info(ctypes.byref(a), ctypes.byref(b)))
    return {"occurrences": a.value, "indexed_cores": b.value}


def targeted_finalize_device(probes: Packed, d_best_ptr: int, clamp: int, d_out_ptr: int, stream: int = 0) -> None:
    _check(load_lib().k4b_targeted_finalize_device(probes.handle, _vp(d_best_ptr), clamp, _vp(d_out_ptr), _vp(stream)))


def last_diag_info() -> dict:
    v = [ctypes.c_uint32(0) for _ in range(4)]
    _check(load_lib().k4b_last_diag_info(*[ctypes.byref(x) for x in v]))
    return dict(zip(("np_full", "np_small", "slabs", "narrow_slabs"), (x.value for x in v)))


def best_finalize_device(g: Packed, d_best_ptr: int, d_out_ptr: int, stream: int = 0) -> None:
    _check(load_lib().k4b_best_finalize_device(g.handle, _vp(d_best_ptr), _vp(d_out_ptr), _vp(stream)))


def set_reference_sensitivity(sensitivity: int) -> None:
    """-s of the reference run the results will be compared with (0 default, 1 more, 2 ultra, 3 less)."""
    _check(load_lib().k4b_set_reference_sensitivity(sensitivity))


def last_depth_cut() -> dict:
    """Probe K-mers of the last seed-engine targeted() call where the reference's depth cut may fire."""
    n, cap = ctypes.c_uint64(0), ctypes.c_uint32(0)
    _check(load_lib().k4b_last_depth_cut(ctypes.byref(n), ctypes.byref(cap)))
    return {"probe_kmers": int(n.value), "max_copies": int(cap.value)}


def histogram(minima, K: int) -> np.ndarray:
    """Distribution of an exhaustive result: uint64[K+2], bin d = positions whose minimum is d, bin K+1 =
    positions where no K-mer starts (GPU histogram, k4b_hamm_histogram)."""
    m = np.ascontiguousarray(minima, dtype=np.uint16)
    hist = np.zeros(K + 2, dtype=np.uint64)
    _check(load_lib().k4b_hamm_histogram(m.ctypes.data_as(_u16p), len(m), K, hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))))
    return hist


def packed_image_bytes(length: int) -> int:
    return int(load_lib().k4b_packed_image_bytes(length))


def last_kernel_ms() -> float:
    return float(load_lib().k4b_last_kernel_ms())


def microbench_intpipe(which: int, iters: int = 2000) -> float:
    g = ctypes.c_double(0)
    _check(load_lib().k4b_microbench_intpipe(which, iters, ctypes.byref(g)))
    return g.value
