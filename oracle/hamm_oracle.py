"""TEST INFRASTRUCTURE ONLY - never imported by the product package (kit4b_b200/).

Python side of the oracle: loads the C restatement (oracle/hamm_oracle.c), restates the input
encoding and the output writer of the reference so that oracle arrays can be compared
byte-for-byte with files written by the unmodified reference binary (oracle/_ref), and offers
a third, independent NumPy brute force for tiny cases.

Reference loci restated here:
  * FASTA symbol -> base code: libkit4b/Fasta.cpp:1658-1705 (CFasta::Ascii2Sense), non-alpha
    except '-' dropped at Fasta.cpp:1167; entry name = first whitespace token of the
    descriptor (ngskit4b/genbioseq.cpp:402-404)
  * concatenated genome layout: ngskit4b/hammings.cpp:2981-3134 (LoadGenome)
  * exhaustive CSV writer incl. its short-chromosome mis-step: hammings.cpp:2899-2929
  * bioseq container: libkit4b/BioSeqFile.h:28-58, BioSeqFile.cpp:1123-1220
"""
from __future__ import annotations

import ctypes
import os
import struct
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libk4oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

EOS = 7
CPL = np.array([3, 2, 1, 0, 4, 5, 6, 7], dtype=np.uint8)  # MapCpl, hammings.cpp:3173-3180


# --------------------------------------------------------------------------------------------
# building / loading
# --------------------------------------------------------------------------------------------
def build(force: bool = False) -> str:
    src = os.path.join(HERE, "hamm_oracle.c")
    if (not force and os.path.exists(LIB_PATH)
            and (not os.path.exists(src) or os.path.getmtime(LIB_PATH) >= os.path.getmtime(src))):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-pthread", "-o", LIB_PATH, src])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        u8p, u16p = ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint16)
        L.k4o_exhaustive_sliding.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, u16p, ctypes.c_int]
        L.k4o_exhaustive_sliding.restype = ctypes.c_uint64
        L.k4o_exhaustive_sliding_frac.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, u16p,
                                                  ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
        L.k4o_exhaustive_sliding_frac.restype = ctypes.c_uint64
        L.k4o_exhaustive_brute.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                           ctypes.c_uint32, u16p]
        L.k4o_exhaustive_brute.restype = None
        L.k4o_targeted_brute.argtypes = [u8p, ctypes.c_uint32, u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                         ctypes.c_int, u8p]
        L.k4o_targeted_brute.restype = None
        L.k4o_targeted_self_brute.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, u8p]
        L.k4o_targeted_self_brute.restype = None
        L.k4o_targeted_self_brute_z.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, u8p]
        L.k4o_targeted_self_brute_z.restype = None
        L.k4o_exhaustive_sliding_sweep.argtypes = [u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, u16p,
                                                   ctypes.c_uint32, ctypes.c_uint32]
        L.k4o_exhaustive_sliding_sweep.restype = None
        _lib = L
    return _lib


def ref_binary(nosleep: bool = True):
    """Path of the reference binary built by oracle/build_ref.sh, or None."""
    p = os.path.join(REF_DIR, "ngskit4b_ref_nosleep" if nosleep else "ngskit4b_ref")
    return p if os.access(p, os.X_OK) else None


def _u8(a):
    if a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("oracle inputs must be contiguous uint8 arrays")
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _u16(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))


# --------------------------------------------------------------------------------------------
# input encoding
# --------------------------------------------------------------------------------------------
_ASCII2CODE = np.full(256, 255, dtype=np.uint8)
for _ch, _code in (("a", 0), ("c", 1), ("g", 2), ("t", 3), ("u", 3)):
    _ASCII2CODE[ord(_ch)] = _code | 0x08        # lower case: soft-mask flag set
    _ASCII2CODE[ord(_ch.upper())] = _code
for _c in range(ord("a"), ord("z") + 1):
    for _cc in (_c, _c - 32):
        if _ASCII2CODE[_cc] == 255:
            _ASCII2CODE[_cc] = 4                     # any other letter -> N, no mask
_ASCII2CODE[ord("-")] = 6                            # InDel


def encode_fasta(text: str) -> List[Tuple[str, np.ndarray]]:
    """FASTA text -> [(entry name, codes incl. soft-mask bit 0x08)]."""
    out = []
    name, chunks = None, []
    for line in text.splitlines():
        if line.startswith(">"):
            if name is not None:
                out.append((name, np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)))
            toks = line[1:].split()
            name = toks[0] if toks else ""
            chunks = []
        elif name is not None:
            raw = np.frombuffer(line.encode("latin-1"), dtype=np.uint8)
            codes = _ASCII2CODE[raw]
            chunks.append(codes[codes != 255])
    if name is not None:
        out.append((name, np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)))
    return out


def concat_entries(entries: Sequence[Tuple[str, np.ndarray]]):
    """LoadGenome layout without the EOG sentinels: chr1 EOS chr2 EOS ... chrN.

    Returns (concat uint8, chroms [(name, start, len)], genome_len) where genome_len is the
    reference's m_GenomeLen = len(concat) + 2."""
    parts, chroms, pos = [], [], 0
    for i, (name, codes) in enumerate(entries):
        c = (codes & 0x07).astype(np.uint8)  # strip the soft-mask flag, hammings.cpp:3073-3074
        chroms.append((name[:80], pos, len(c)))
        parts.append(c)
        pos += len(c)
        if i + 1 < len(entries):
            parts.append(np.array([EOS], dtype=np.uint8))
            pos += 1
    concat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    return np.ascontiguousarray(concat), chroms, len(concat) + 2


# --------------------------------------------------------------------------------------------
# engines
# --------------------------------------------------------------------------------------------
def exhaustive_sliding(concat: np.ndarray, K: int, both: bool, threads: int = 0) -> np.ndarray:
    hd = np.full(len(concat), K + 1, dtype=np.uint16)
    if len(concat):
        lib().k4o_exhaustive_sliding(_u8(concat), len(concat), K, int(both), _u16(hd), threads or os.cpu_count() or 1)
    return hd


def exhaustive_sliding_sweep(concat: np.ndarray, K: int, both: bool, sweep_start: int, sweep_end: int) -> np.ndarray:
    """Only the pairs visited by sweep instances sweep_start..sweep_end (-b/-B, -m2)."""
    hd = np.full(len(concat), K + 1, dtype=np.uint16)
    if len(concat):
        lib().k4o_exhaustive_sliding_sweep(_u8(concat), len(concat), K, int(both), _u16(hd), sweep_start, sweep_end)
    return hd


def exhaustive_sliding_sample(concat: np.ndarray, K: int, both: bool, threads: int, num: int, den: int):
    """Runs the leading num/den of the diagonals; returns (hd, cells visited)."""
    hd = np.full(len(concat), K + 1, dtype=np.uint16)
    cells = lib().k4o_exhaustive_sliding_frac(_u8(concat), len(concat), K, int(both), _u16(hd), threads, num, den)
    return hd, int(cells)


def exhaustive_brute(concat: np.ndarray, K: int, both: bool, q_begin: int = 0, q_end: int | None = None) -> np.ndarray:
    hd = np.full(len(concat), K + 1, dtype=np.uint16)
    if len(concat):
        lib().k4o_exhaustive_brute(_u8(concat), len(concat), K, int(both), q_begin,
                                   len(concat) if q_end is None else q_end, _u16(hd))
    return hd


def targeted_brute(target: np.ndarray, probes: np.ndarray, K: int, R: int, both: bool) -> np.ndarray:
    out = np.full(len(probes), 0xFF, dtype=np.uint8)
    lib().k4o_targeted_brute(_u8(target), len(target), _u8(probes), len(probes), K, R, int(both), _u8(out))
    return out


def targeted_self_brute(target: np.ndarray, K: int, R: int, both: bool, zmode: int = 0) -> np.ndarray:
    """no -I: the probes are the K-mers of the indexed assembly itself; zmode = -z (0 both,
    1 intra-entry only, 2 inter-entry only - applies to exact sense-strand hits only)."""
    out = np.full(len(target), 0xFF, dtype=np.uint8)
    lib().k4o_targeted_self_brute_z(_u8(target), len(target), K, R, int(both), int(zmode), _u8(out))
    return out


def valid_starts(concat: np.ndarray, K: int) -> np.ndarray:
    """bool[len]: K-mer window starting here lies inside one chromosome."""
    n = len(concat)
    ok = np.zeros(n, dtype=bool)
    if n < K:
        return ok
    eos = (concat == EOS).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(eos)])
    ok[: n - K + 1] = (cs[K:] - cs[: n - K + 1]) == 0
    return ok


def numpy_brute(concat: np.ndarray, K: int, both: bool) -> np.ndarray:
    """Independent restatement - tiny inputs only (O(N^2 K) memory-light)."""
    n = len(concat)
    hd = np.full(n, K + 1, dtype=np.uint16)
    ok = valid_starts(concat, K)
    idx = np.nonzero(ok)[0]
    if len(idx) == 0:
        return hd
    X = np.stack([concat[i:i + K] for i in idx])
    RC = CPL[X[:, ::-1]]
    for r, i in enumerate(idx):
        d = (X != X[r]).sum(1)
        d[r] = K + 1
        best = d.min() if len(d) else K + 1
        if both:
            best = min(best, (RC != X[r]).sum(1).min())
        hd[i] = best
    return hd


# --------------------------------------------------------------------------------------------
# output writer (literal restatement of hammings.cpp:2899-2929, quirks included)
# --------------------------------------------------------------------------------------------
def exhaustive_csv(genome_len: int, chroms, K: int, hd: np.ndarray, sweep_start: int = 1,
                   sweep_end: int | None = None) -> bytes:
    if sweep_end is None:
        sweep_end = genome_len
    lines = ["%u,%d,%d\n" % (genome_len, sweep_start + 1, sweep_end)]
    nsub = [max(0, ln - K + 1) for (_, _, ln) in chroms]
    ci, cur, seq_idx = 0, 0, 0
    n = genome_len - 2
    vals = hd.tolist()
    while seq_idx < n:
        if cur >= nsub[ci]:
            if ci == len(chroms) - 1:
                break
            ci += 1
            cur = 0
            seq_idx += K
        v = vals[seq_idx] if seq_idx < len(vals) else K + 1  # the reference reads past the end here
        if v <= K:
            lines.append('"%s",%d,%d\n' % (chroms[ci][0], cur, v))
        seq_idx += 1
        cur += 1
    return "".join(lines).encode("latin-1")


def restricted_per_loci(chroms, concat_values: np.ndarray) -> np.ndarray:
    """concat-layout results (with separator slots) -> the reference's per-loci array
    H[sum of entry lengths] (hammings.cpp:2336-2354, :1615-1674)."""
    return np.concatenate([concat_values[st:st + ln] for (_, st, ln) in chroms]) if chroms else np.zeros(0, np.uint8)


def restricted_report(chroms, K: int, R: int, h: np.ndarray, fmt: int, out_name: str = "", sensitivity: int = 0,
                      prefix: str = "") -> bytes:
    """Literal restatement of the three restricted-mode writers incl. their off-by-ones
    (hammings.cpp:1711-1849 CSV, :1853-1984 BED, :1987-2118 Wiggle)."""
    tag = (prefix + "|#") if prefix else ""
    out = []
    if fmt == 0:
        out.append('"Chrom","StartLoci","Len","Hamming"\n')
    elif fmt == 1:
        out.append('track type=bedGraph name=ResHamming%d_%d description="Restricted Hammings for K-mer length %d '
                   'and Hamming limit %d"\n' % (K, R, K, R))
    else:
        out.append('track type=wiggle_0 color=50,150,255 autoScale=off maxHeightPixels=128:32:8 name="Hammings '
                   '(%d,%d,%d) - %s" description="Hammings Sensitivity: %d KMerLen: %d RHamm: %d  for %s"\n'
                   % (sensitivity, K, R, out_name, sensitivity, K, R, out_name))
    ofs = 0
    vals = h.tolist()
    for (name, _st, ln) in chroms:
        base = ofs
        ofs += ln
        if ln < K or (tag and name[:len(tag)].lower() != tag.lower()):
            continue
        nm = name[len(tag):]
        if fmt in (0, 1):
            run_len = 0 if fmt == 0 else 1
            cur, run_loci, p = vals[base], 0, base
            for loci in range(1, ln - K + 1):
                if vals[p] == cur:
                    run_len += 1
                else:
                    out.append(('"%s",%d,%d,%d\n' % (nm, run_loci, run_len, cur)) if fmt == 0 else
                               ("%s\t%d\t%d\t%d\n" % (nm, run_loci, run_loci + run_len - 1, cur)))
                    run_len, cur, run_loci = 1, vals[p], loci
                p += 1
            out.append(('"%s",%d,%d,%d\n' % (nm, run_loci, run_len, cur)) if fmt == 0 else
                       ("%s\t%d\t%d\t%d\n" % (nm, run_loci, run_loci + run_len - 1, cur)))
        else:
            span_len, span_start, cur, loci = 0, 0, vals[base], 0
            for loci in range(0, ln - K + 1):
                v = vals[base + loci]
                if v != cur:
                    out.append("variableStep chrom=%s span=%d\n%d %d\n" % (nm, span_len, span_start + 1, cur))
                    cur, span_len, span_start = v, 0, loci + 1
                span_len += 1
            loci = ln - K + 1
            if span_start != loci:
                out.append("variableStep chrom=%s span=%d\n%d %d\n" % (nm, span_len, span_start + 1, cur))
    return "".join(out).encode("latin-1")


def read_sfx(path: str):
    """(entries [(name, start, len)], concatenated sequence) of an 'sfx5' file (SfxArray.h:98-123, :194-207)."""
    b = open(path, "rb").read()
    if b[:4] != b"sfx5":
        raise ValueError("not an sfx5 file")
    entries_ofs, entries_size = struct.unpack_from("<QI", b, 20)
    block_ofs = struct.unpack_from("<Q", b, 44)[0]
    nent = struct.unpack_from("<I", b, entries_ofs)[0]
    entries = []
    for i in range(nent):
        off = entries_ofs + 8 + i * 111
        name = b[off + 8: b.index(b"\0", off + 8)].decode("latin-1")
        seqlen = struct.unpack_from("<I", b, off + 8 + 81 + 2)[0]
        start = struct.unpack_from("<Q", b, off + 8 + 81 + 2 + 4)[0]
        entries.append((name, start, seqlen))
    concat_len = struct.unpack_from("<Q", b, block_ofs + 8)[0]
    seq = np.frombuffer(b, dtype=np.uint8, count=concat_len, offset=block_ofs + 20) & 0x07
    return entries, np.ascontiguousarray(seq)


# --------------------------------------------------------------------------------------------
# bioseq container (reader restated for tests; writer so inputs can be made without the reference)
# --------------------------------------------------------------------------------------------
def read_bioseq(path: str) -> List[Tuple[str, np.ndarray]]:
    b = open(path, "rb").read()
    if b[:4] != b"bios":
        raise ValueError("not a bioseq file")
    dir_ofs, id_idx_ofs = struct.unpack_from("<q", b, 32)[0], struct.unpack_from("<q", b, 40)[0]
    ftype, version, _maxe, nent, _dsz = struct.unpack_from("<5i", b, 56)
    if ftype != 1 or version != 10:
        raise ValueError("unsupported bioseq type/version")
    out = []
    for e in range(nent):
        off = dir_ofs + struct.unpack_from("<i", b, id_idx_ofs + 4 * e)[0]
        data_psn, size, _eid, _ni, data_len, _hash, _flags = struct.unpack_from("<qIiiIHB", b, off)
        name = b[off + 27: b.index(b"\0", off + 27)].decode("latin-1")
        raw = np.frombuffer(b, dtype=np.uint8, count=(data_len + 1) // 2, offset=data_psn)
        codes = np.empty(2 * len(raw), dtype=np.uint8)
        codes[0::2] = raw & 0x0F
        codes[1::2] = raw >> 4
        out.append((name, codes[:data_len].copy()))
    return out


def _name_hash(name: bytes) -> int:
    # only used by name look-ups in the reference (never on the hammings path); any 16-bit value works
    h = 0
    for c in name.lower():
        h = (h * 19 + c) & 0xFFFF
    return h


def write_bioseq(path: str, entries: Sequence[Tuple[str, np.ndarray]], title: str = "k4b") -> None:
    """Minimal writer of the 'bios' container (type 1, version 10) readable by CBioSeqFile."""
    hdr = bytearray(1248)
    data = bytearray()
    dirs = bytearray()
    offs = []
    seq_ofs = 1248
    for eid, (name, codes) in enumerate(entries, 1):
        nb = name.encode("latin-1")
        n = len(codes)
        c = np.asarray(codes, dtype=np.uint8) & 0x0F
        if n & 1:
            c = np.concatenate([c, np.zeros(1, np.uint8)])
        packed = (c[0::2] | (c[1::2] << 4)).astype(np.uint8).tobytes()
        size = 27 + len(nb) + 1 + len(nb) + 1
        offs.append(len(dirs))
        dirs += struct.pack("<qIiiIHB", seq_ofs + len(data), size, eid, 1, n, _name_hash(nb), 0x01)
        dirs += nb + b"\0" + nb + b"\0"
        data += packed
    dir_ofs = seq_ofs + len(data)
    name_idx_ofs = dir_ofs + len(dirs)
    id_idx_ofs = name_idx_ofs + 8 * len(entries)
    file_len = id_idx_ofs + 8 * len(entries)
    order = sorted(range(len(entries)), key=lambda i: entries[i][0].lower())
    hdr[0:4] = b"bios"
    struct.pack_into("<6q", hdr, 8, file_len, seq_ofs, len(data), dir_ofs, id_idx_ofs, name_idx_ofs)
    struct.pack_into("<5i", hdr, 56, 1, 10, 20000000, len(entries), len(dirs))
    t = title.encode("latin-1")[:63]
    hdr[76:76 + len(t)] = t
    hdr[157:157 + len(t)] = t
    hdr[1181:1181 + len(t)] = t
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)
        f.write(dirs)
        f.write(b"".join(struct.pack("<i", offs[i]) for i in order).ljust(8 * len(entries), b"\0"))
        f.write(b"".join(struct.pack("<i", o) for o in offs).ljust(8 * len(entries), b"\0"))
