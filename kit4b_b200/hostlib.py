"""ctypes mirror of the host-side C hooks (kit4b_b200/csrc/host/k4b_host_capi.cpp): bioseq /
suffix-file readers, the concatenated genome layout, the byte-exact report writers and the
CLI parser of the `hammings` drop-in.  No numeric work; usable without a GPU."""
from __future__ import annotations

import ctypes
import os
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib_path() -> str:
    return os.path.join(_HERE, "libk4bhost.so")


def cli_path() -> str:
    return os.path.join(_HERE, "bin", "k4b_hammings")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise RuntimeError("%s is missing - run __graft_entry__.build()" % lib_path())
        L = ctypes.CDLL(lib_path())
        L.k4bh_last_error.restype = ctypes.c_char_p
        L.k4bh_concat_from_bioseq.restype = ctypes.c_long
        L.k4bh_concat_from_bioseq.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]
        L.k4bh_write_exhaustive_csv.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.c_uint32, ctypes.c_uint32, ctypes.c_char_p]
        L.k4bh_write_restricted.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p]
        L.k4bh_read_sfx.restype = ctypes.c_long
        L.k4bh_read_sfx.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.k4bh_fasta_to_bioseq.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.k4bh_sweep_range.restype = None
        L.k4bh_sweep_range.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.k4bh_merge_csv.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.k4bh_csv_to_bham.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.k4bh_bham_to_csv.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.k4bh_parse_cli.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int),
                                     ctypes.c_char_p]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise RuntimeError("k4b host error %d: %s" % (rc, load().k4bh_last_error().decode("utf-8", "replace")))
    return rc


def _table(buf) -> List[Tuple[str, int, int]]:
    out = []
    for line in buf.value.decode("latin-1").splitlines():
        name, start, ln = line.split("\t")
        out.append((name, int(start), int(ln)))
    return out


def concat_from_bioseq(path: str, K: int):
    """(concat uint8, chroms [(name, start, len)], genome_len) as built by LoadGenome."""
    L = load()
    glen = ctypes.c_uint32(0)
    n = _check(L.k4bh_concat_from_bioseq(path.encode(), K, None, 0, None, 0, ctypes.byref(glen)))
    out = np.empty(n, dtype=np.uint8)
    names = ctypes.create_string_buffer(1 << 20)
    _check(L.k4bh_concat_from_bioseq(path.encode(), K, out.ctypes.data, n, names, len(names), ctypes.byref(glen)))
    return out, _table(names), glen.value


def write_exhaustive_csv(bioseq: str, K: int, hd: np.ndarray, out_path: str, sweep_start: int = 1, sweep_end: int = 0):
    hd = np.ascontiguousarray(hd, dtype=np.uint16)
    _check(load().k4bh_write_exhaustive_csv(bioseq.encode(), K, hd.ctypes.data, len(hd), sweep_start, sweep_end,
                                            out_path.encode()))


def write_restricted(probe_bioseq: str, K: int, R: int, fmt: int, h: np.ndarray, out_path: str, sensitivity: int = 0,
                     prefix: str = ""):
    h = np.ascontiguousarray(h, dtype=np.uint8)
    _check(load().k4bh_write_restricted(probe_bioseq.encode(), K, R, sensitivity, fmt, prefix.encode(), h.ctypes.data,
                                        len(h), out_path.encode()))


def read_sfx(path: str):
    L = load()
    n = _check(L.k4bh_read_sfx(path.encode(), None, 0, None, 0))
    out = np.empty(n, dtype=np.uint8)
    ents = ctypes.create_string_buffer(1 << 20)
    _check(L.k4bh_read_sfx(path.encode(), out.ctypes.data, n, ents, len(ents)))
    return _table(ents), out


def fasta_to_bioseq(fasta: str, bioseq: str, title: str = "k4b"):
    _check(load().k4bh_fasta_to_bioseq(fasta.encode(), bioseq.encode(), title.encode()))


def node_sweep_range(genome_len: int, num_chroms: int, watson_only: bool, num_nodes: int, node: int):
    """(SSeqStart, SSeqEnd) of a -m2 node slice."""
    a, b = ctypes.c_uint32(0), ctypes.c_uint32(0)
    load().k4bh_sweep_range(genome_len, num_chroms, int(watson_only), num_nodes, node, 0, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def merge_csv(src: str, into: str):
    _check(load().k4bh_merge_csv(src.encode(), into.encode()))


def csv_to_bham(csv: str, bham: str):
    _check(load().k4bh_csv_to_bham(csv.encode(), bham.encode()))


def hamming_dist(csvs: List[str], out: str):
    """HammingDist, region-less mode: distribution of field 3 of the `"chrom",loci,hamming` rows."""
    arr = (ctypes.c_char_p * len(csvs))(*[c.encode() for c in csvs])
    _check(load().k4bh_hamming_dist(len(csvs), arr, out.encode()))


def hamming_dist_regions(csvs: List[str], feats: str, out: str, reg_len: int = 0, ofs_loci: int = 0):
    """HammingDist, region mode (-I feats [-r reg_len] [-R ofs_loci]): per-region distributions, bit-exact
    with the reference; feats = BED text or the biobed container of `genbiobed`."""
    arr = (ctypes.c_char_p * len(csvs))(*[c.encode() for c in csvs])
    _check(load().k4bh_hamming_dist_regions(len(csvs), arr, feats.encode(), reg_len, ofs_loci, out.encode()))


def feature_bits(feats: str, chrom: str, loci: List[int], reg_len: int = 0) -> List[int]:
    """Region feature bits (CDS 1, 5'UTR 2, 3'UTR 4, intron 8, upstream 16, downstream 32) of single loci;
    -1 where the chromosome is not in the feature file."""
    n = len(loci)
    a = (ctypes.c_int * n)(*loci)
    b = (ctypes.c_int * n)()
    _check(load().k4bh_feature_bits(feats.encode(), chrom.encode(), n, a, reg_len, b))
    return list(b)


def bham_to_csv(bham: str, csv: str):
    _check(load().k4bh_bham_to_csv(bham.encode(), csv.encode()))


CLI_INT_FIELDS = ("mode", "sensitivity", "resformat", "crick", "intrainterboth", "rhamm", "numnodes", "node",
                  "sweep_start", "sweep_end", "K", "sample", "threads", "gpus", "help", "version")


def parse_cli(args: List[str]) -> dict:
    argv = [b"k4b_hammings"] + [a.encode() for a in args]
    arr = (ctypes.c_char_p * len(argv))(*argv)
    ints = (ctypes.c_int * 16)()
    strs = ctypes.create_string_buffer(4 * 512)
    _check(load().k4bh_parse_cli(len(argv), arr, ints, strs))
    d = dict(zip(CLI_INT_FIELDS, list(ints)))
    raw = strs.raw
    for i, nm in enumerate(("in_file", "in_seq_file", "out_file", "prefix")):
        d[nm] = raw[512 * i: 512 * (i + 1)].split(b"\0")[0].decode()
    return d
