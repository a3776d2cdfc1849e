// k4b_containers.cpp - input containers of the `hammings` drop-in (bioseq, suffix-array file,
// FASTA) and the concatenated genome layout.  Written from the on-disk formats documented in
// SURVEY.md 8a (a2-a4, a15); see k4b_host.h for the reference loci each function mirrors.
#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>

#include "k4b_host.h"

#include <zlib.h>

namespace k4bhost {

namespace {

bool slurp(const std::string &path, std::vector<uint8_t> &buf, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "unable to open '" + path + "': " + strerror(errno);
        return false;
    }
    fseeko(f, 0, SEEK_END);
    const off_t n = ftello(f);
    fseeko(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) {
        err = "short read on '" + path + "'";
        return false;
    }
    return true;
}

template <typename T>
T rd(const std::vector<uint8_t> &b, size_t off) {
    T v;
    memcpy(&v, b.data() + off, sizeof(T));
    return v;  // files are little-endian, as is every host this runs on
}

std::string cstr(const std::vector<uint8_t> &b, size_t off, size_t maxlen) {
    size_t n = 0;
    while (n < maxlen && off + n < b.size() && b[off + n]) ++n;
    return std::string((const char *)b.data() + off, n);
}

// bioseq header field offsets (tsBSFHeader, #pragma pack(8); SURVEY.md a3, verified there)
constexpr size_t kBsHdrSize = 1248;
constexpr size_t kBsDirFixed = 27;  // bytes before szName in a directory entry (pack(1))

}  // namespace

int read_bioseq(const std::string &path, std::vector<SeqEntry> &entries, std::string &title,
                std::string &err) {
    entries.clear();
    std::vector<uint8_t> b;
    if (!slurp(path, b, err)) return kErrOpnFile;
    if (b.size() < kBsHdrSize || memcmp(b.data(), "bios", 4) != 0) {
        err = "'" + path + "' is not a bioseq file";
        return kErrNotBioseq;
    }
    const int64_t dir_ofs = rd<int64_t>(b, 32);
    const int64_t id_idx_ofs = rd<int64_t>(b, 40);
    const int32_t type = rd<int32_t>(b, 56), version = rd<int32_t>(b, 60);
    const int32_t nent = rd<int32_t>(b, 68), dir_size = rd<int32_t>(b, 72);
    if (type != 1) {
        err = "bioseq file does not hold sequences (type " + std::to_string(type) + ")";
        return kErrFileType;
    }
    if (version != 10) {
        err = "unsupported bioseq version " + std::to_string(version);
        return kErrFileVer;
    }
    title = cstr(b, 1181, 64);
    if (nent < 0 || dir_ofs < 0 || id_idx_ofs < 0 || (uint64_t)dir_ofs + (uint64_t)dir_size > b.size() ||
        (uint64_t)id_idx_ofs + 4ull * (uint64_t)nent > b.size()) {
        err = "bioseq directory lies outside the file";
        return kErrFileAccess;
    }
    entries.reserve((size_t)nent);
    for (int32_t e = 0; e < nent; ++e) {
        const int32_t rel = rd<int32_t>(b, (size_t)id_idx_ofs + 4 * (size_t)e);
        const size_t off = (size_t)dir_ofs + (size_t)rel;
        if (rel < 0 || off + kBsDirFixed + 1 > b.size()) {
            err = "corrupt bioseq directory index";
            return kErrFileAccess;
        }
        const int64_t data_psn = rd<int64_t>(b, off);
        const uint32_t data_len = rd<uint32_t>(b, off + 20);
        const uint8_t flags = b[off + 26];
        if ((flags & 0x0f) != 1) {
            err = "bioseq entry is not nibble-packed sequence data";
            return kErrFileType;
        }
        SeqEntry se;
        se.name = cstr(b, off + kBsDirFixed, 80);  // GetName keeps at most 80 chars
        const size_t nbytes = ((size_t)data_len + 1) / 2;
        if (data_psn < 0 || (uint64_t)data_psn + nbytes > b.size()) {
            err = "bioseq entry data lies outside the file";
            return kErrFileAccess;
        }
        se.codes.resize(data_len);
        const uint8_t *p = b.data() + data_psn;
        for (uint32_t i = 0; i + 1 < data_len; i += 2) {
            const uint8_t v = p[i >> 1];
            se.codes[i] = v & 0x0f;       // base 2i in the low nibble (BioSeqFile.cpp:1155-1159)
            se.codes[i + 1] = v >> 4;
        }
        if (data_len & 1) se.codes[data_len - 1] = p[data_len >> 1] & 0x0f;
        entries.push_back(std::move(se));
    }
    return kOk;
}

int write_bioseq(const std::string &path, const std::vector<SeqEntry> &entries,
                 const std::string &title, std::string &err) {
    std::vector<uint8_t> hdr(kBsHdrSize, 0), data, dirs;
    std::vector<int32_t> offs;
    for (size_t e = 0; e < entries.size(); ++e) {
        const SeqEntry &se = entries[e];
        const std::string nm = se.name.substr(0, 80);
        const uint32_t n = (uint32_t)se.codes.size();
        const int64_t psn = (int64_t)kBsHdrSize + (int64_t)data.size();
        for (uint32_t i = 0; i < n; i += 2) {
            const uint8_t lo = se.codes[i] & 0x0f;
            const uint8_t hi = (i + 1 < n) ? (se.codes[i + 1] & 0x0f) : 0;
            data.push_back((uint8_t)(lo | (hi << 4)));
        }
        const uint32_t size = (uint32_t)(kBsDirFixed + nm.size() + 1 + nm.size() + 1);
        offs.push_back((int32_t)dirs.size());
        uint8_t fixed[kBsDirFixed];
        uint16_t hash = 0;
        for (char c : nm) hash = (uint16_t)(hash * 19 + (uint8_t)tolower(c));
        const int32_t eid = (int32_t)e + 1, inst = 1;
        memcpy(fixed + 0, &psn, 8);
        memcpy(fixed + 8, &size, 4);
        memcpy(fixed + 12, &eid, 4);
        memcpy(fixed + 16, &inst, 4);
        memcpy(fixed + 20, &n, 4);
        memcpy(fixed + 24, &hash, 2);
        fixed[26] = 0x01;  // DType = nibble-packed bases
        dirs.insert(dirs.end(), fixed, fixed + kBsDirFixed);
        for (int rep = 0; rep < 2; ++rep) {
            dirs.insert(dirs.end(), nm.begin(), nm.end());
            dirs.push_back(0);
        }
    }
    const int64_t seq_ofs = kBsHdrSize, seq_size = (int64_t)data.size();
    const int64_t dir_ofs = seq_ofs + seq_size;
    const int64_t name_idx_ofs = dir_ofs + (int64_t)dirs.size();
    const int64_t id_idx_ofs = name_idx_ofs + 8 * (int64_t)entries.size();
    const int64_t file_len = id_idx_ofs + 8 * (int64_t)entries.size();
    memcpy(hdr.data(), "bios", 4);
    const int64_t q[6] = {file_len, seq_ofs, seq_size, dir_ofs, id_idx_ofs, name_idx_ofs};
    memcpy(hdr.data() + 8, q, sizeof(q));
    const int32_t i5[5] = {1, 10, 20000000, (int32_t)entries.size(), (int32_t)dirs.size()};
    memcpy(hdr.data() + 56, i5, sizeof(i5));
    const std::string t = title.substr(0, 63);
    memcpy(hdr.data() + 76, t.data(), t.size());
    memcpy(hdr.data() + 157, t.data(), t.size());
    memcpy(hdr.data() + 1181, t.data(), t.size());
    std::vector<size_t> order(entries.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t c) {
        return strcasecmp(entries[a].name.c_str(), entries[c].name.c_str()) < 0;
    });
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) {
        err = "unable to create '" + path + "': " + strerror(errno);
        return kErrCreateFile;
    }
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
    ok = ok && (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size());
    ok = ok && (dirs.empty() || fwrite(dirs.data(), 1, dirs.size(), f) == dirs.size());
    std::vector<uint8_t> idx(8 * entries.size(), 0);
    for (size_t i = 0; i < order.size(); ++i) memcpy(idx.data() + 4 * i, &offs[order[i]], 4);
    ok = ok && (idx.empty() || fwrite(idx.data(), 1, idx.size(), f) == idx.size());
    std::fill(idx.begin(), idx.end(), 0);
    for (size_t i = 0; i < offs.size(); ++i) memcpy(idx.data() + 4 * i, &offs[i], 4);
    ok = ok && (idx.empty() || fwrite(idx.data(), 1, idx.size(), f) == idx.size());
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        err = "write failed on '" + path + "'";
        return kErrFileAccess;
    }
    return kOk;
}

// FASTA, plain or gzip-compressed (zlib's gzFile reads both).  Rules of the reference's front end
// (genbioseq -> CFasta): descriptor lines start with '>', the entry name is the first
// whitespace-delimited token, at most 80 characters (genbioseq.cpp:402-404, commdefs.h:140);
// sequence symbols aAcCgGtTuU -> 0..3 with the soft-mask flag on lower case, '-' -> InDel (6), any
// other letter -> N (4), everything else is dropped (Fasta.cpp:1167, :1658-1705).
int read_fasta(const std::string &path, std::vector<SeqEntry> &entries, std::string &err) {
    entries.clear();
    gzFile in = gzopen(path.c_str(), "rb");
    if (!in) {
        err = "unable to open '" + path + "'";
        return kErrOpnFile;
    }
    gzbuffer(in, 1 << 20);
    uint8_t map[256];
    memset(map, 0xff, sizeof(map));
    for (int c = 'a'; c <= 'z'; ++c) map[c] = map[c - 32] = 4;
    const char *acgt = "acgt";
    for (int i = 0; i < 4; ++i) {
        map[(uint8_t)acgt[i]] = (uint8_t)(i | 0x08);
        map[(uint8_t)(acgt[i] - 32)] = (uint8_t)i;
    }
    map['u'] = 3 | 0x08;
    map['U'] = 3;
    map['-'] = 6;
    std::vector<unsigned char> buf(4 << 20);
    std::string descr;       // descriptor line being collected (may span read chunks)
    bool in_descr = false, at_line_start = true, have = false;
    int n;
    while ((n = gzread(in, buf.data(), (unsigned)buf.size())) > 0) {
        for (int k = 0; k < n; ++k) {
            const unsigned char ch = buf[k];
            if (in_descr) {
                if (ch == '\n') {
                    SeqEntry se;
                    size_t b = 0;
                    while (b < descr.size() && isspace((unsigned char)descr[b])) ++b;
                    size_t e = b;
                    while (e < descr.size() && !isspace((unsigned char)descr[e])) ++e;
                    se.name = descr.substr(b, std::min<size_t>(e - b, 80));
                    entries.push_back(std::move(se));
                    have = true;
                    in_descr = false;
                    at_line_start = true;
                } else {
                    descr.push_back((char)ch);
                }
                continue;
            }
            if (ch == '\n') {
                at_line_start = true;
                continue;
            }
            if (at_line_start && ch == '>') {
                in_descr = true;
                descr.clear();
                continue;
            }
            at_line_start = false;
            if (have && map[ch] != 0xff) entries.back().codes.push_back(map[ch]);
        }
    }
    int zerr = 0;
    const char *zmsg = gzerror(in, &zerr);
    const bool bad = n < 0 || (zerr != Z_OK && zerr != Z_STREAM_END);
    const std::string zs = zmsg ? zmsg : "";
    gzclose(in);
    if (bad) {
        err = "error while reading '" + path + "': " + zs;
        return kErrFileAccess;
    }
    if (in_descr) {  // descriptor on the last line without a newline: an entry without sequence
        SeqEntry se;
        size_t e = 0;
        while (e < descr.size() && !isspace((unsigned char)descr[e])) ++e;
        se.name = descr.substr(0, std::min<size_t>(e, 80));
        entries.push_back(std::move(se));
        have = true;
    }
    if (!have) {
        err = "'" + path + "' holds no FASTA descriptor line";
        return kErrParse;
    }
    return kOk;
}

// 'b' bioseq container, 's' suffix-array container, 'f' FASTA (plain or gzip), 0 unknown / unreadable
int sniff_format(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return 0;
    unsigned char m[4] = {0, 0, 0, 0};
    const size_t got = fread(m, 1, 4, f);
    fclose(f);
    if (got >= 4 && !memcmp(m, "bios", 4)) return 'b';
    if (got >= 4 && !memcmp(m, "sfx5", 4)) return 's';
    if (got >= 2 && m[0] == 0x1f && m[1] == 0x8b) return 'f';
    for (size_t k = 0; k < got; ++k) {
        if (m[k] == '>') return 'f';
        if (!isspace(m[k])) return 0;
    }
    return 0;
}

// sequences from a bioseq container or straight from FASTA / FASTA.gz (skipping genbioseq)
int read_sequences(const std::string &path, std::vector<SeqEntry> &entries, std::string &title, std::string &err) {
    if (sniff_format(path) == 'f') {
        title = path;
        return read_fasta(path, entries, err);
    }
    return read_bioseq(path, entries, title, err);
}

// the sequence area an `index` run would have produced for these entries: every entry's bases (mask
// stripped) followed by EOS (SfxArray.cpp:1746-1750) - all the GPU engines need of a suffix array file
void sfx_from_entries(const std::vector<SeqEntry> &entries, const std::string &title, SfxData &out) {
    out = SfxData();
    out.title = title;
    out.version = 5;
    uint64_t total = 0;
    for (const SeqEntry &e : entries) total += e.codes.size() + 1;
    out.seq.reserve(total);
    for (const SeqEntry &e : entries) {
        SfxEntry se;
        se.name = e.name;
        se.len = (uint32_t)e.codes.size();
        se.start = out.seq.size();
        for (uint8_t c : e.codes) out.seq.push_back(c & 0x07);
        out.seq.push_back(7);
        out.entries.push_back(std::move(se));
    }
}

int read_sfx(const std::string &path, SfxData &out, std::string &err) {
    // header (tsSfxHeaderV3, pack(4)): only the fixed fields and the name strings are needed
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "unable to open '" + path + "': " + strerror(errno);
        return kErrOpnFile;
    }
    std::vector<uint8_t> h(1224);
    auto bail = [&](int code, const std::string &m) {
        fclose(f);
        err = m;
        return code;
    };
    if (fread(h.data(), 1, h.size(), f) != h.size() || memcmp(h.data(), "sfx5", 4) != 0)
        return bail(kErrFileType, "'" + path + "' is not a suffix array (sfx5) file");
    out.version = rd<int32_t>(h, 4);
    const uint32_t attrs = rd<uint32_t>(h, 8);
    const uint64_t entries_ofs = rd<uint64_t>(h, 20);
    const uint32_t entries_size = rd<uint32_t>(h, 28);
    const uint32_t nblocks = rd<uint32_t>(h, 32);
    const uint64_t block_ofs = rd<uint64_t>(h, 44);
    if (out.version < 4 || out.version > 5)
        return bail(kErrFileVer, "unsupported suffix array file version " + std::to_string(out.version));
    if (attrs & 3u) return bail(kErrFileType, "bisulfite / colorspace suffix arrays are not supported");
    if (nblocks != 1) return bail(kErrFileType, "suffix array file must hold exactly one block");
    out.dataset = cstr(h, 52, 81);
    out.descr = cstr(h, 52 + 81, 1024);
    out.title = cstr(h, 52 + 81 + 1024, 64);
    // entries block: NumEntries u32, MaxEntries u32, then 111-byte tsSfxEntry records (pack(1))
    std::vector<uint8_t> eb(entries_size);
    if (fseeko(f, (off_t)entries_ofs, SEEK_SET) != 0 || fread(eb.data(), 1, eb.size(), f) != eb.size())
        return bail(kErrFileAccess, "unable to read suffix array entries");
    const uint32_t nent = rd<uint32_t>(eb, 0);
    constexpr size_t kEntSize = 4 + 4 + 81 + 2 + 4 + 8 + 8;
    if (8 + (uint64_t)nent * kEntSize > eb.size()) return bail(kErrFileAccess, "corrupt entries block");
    out.entries.clear();
    for (uint32_t i = 0; i < nent; ++i) {
        const size_t off = 8 + (size_t)i * kEntSize;
        SfxEntry e;
        e.name = cstr(eb, off + 8, 81);
        e.len = rd<uint32_t>(eb, off + 8 + 81 + 2);
        e.start = rd<uint64_t>(eb, off + 8 + 81 + 2 + 4);
        out.entries.push_back(std::move(e));
    }
    // block: BlockID u32, NumEntries u32, ConcatSeqLen u64, SfxElSize u32, then the sequence
    uint8_t bh[20];
    if (fseeko(f, (off_t)block_ofs, SEEK_SET) != 0 || fread(bh, 1, sizeof(bh), f) != sizeof(bh))
        return bail(kErrFileAccess, "unable to read suffix block header");
    uint64_t concat_len;
    memcpy(&concat_len, bh + 8, 8);
    out.seq.resize((size_t)concat_len);
    if (concat_len && fread(out.seq.data(), 1, out.seq.size(), f) != out.seq.size())
        return bail(kErrFileAccess, "unable to read suffix block sequence");
    fclose(f);
    for (uint8_t &c : out.seq) c &= 0x07;
    return kOk;
}

void build_genome(const std::vector<SeqEntry> &entries, uint32_t K, Genome &g) {
    g = Genome();
    size_t total = 0;
    for (const SeqEntry &e : entries) total += e.codes.size() + 1;
    g.concat.reserve(total);
    for (size_t i = 0; i < entries.size(); ++i) {
        const SeqEntry &e = entries[i];
        Chrom c;
        c.name = e.name.substr(0, 80);
        c.len = (uint32_t)e.codes.size();
        c.start = (uint32_t)g.concat.size();
        c.num_subseqs = c.len >= K ? c.len - K + 1 : 0;
        g.total_bases += c.len;
        g.num_subseqs += c.num_subseqs;
        for (uint8_t b : e.codes) g.concat.push_back((uint8_t)(b & 0x07));  // strip soft-mask flag
        if (i + 1 < entries.size()) g.concat.push_back(7);                  // eBaseEOS separator
        g.chroms.push_back(std::move(c));
    }
    g.genome_len = (uint32_t)g.concat.size() + 2;
}


// ---- sweep ranges ---------------------------------------------------------------------------------
namespace {
// work-balanced slice boundary L for node k of n (hammings.cpp:1516-1543): total work is
// G*G (Crick) + G*G/2 (Watson); solve G*L + L*L/2 = k/n of it by fixed-point iteration.  The
// float/double mix of the original arithmetic is kept because the integer result is printed
// in the output header.
uint32_t slice_boundary(uint32_t G, int k, int n) {
    if (k < 1) return 0;
    if (k == n) return G;
    const double tot = ((double)G * G) + ((double)G * G) / 2;
    const double want = (tot * k) / n;
    double cur = sqrt((2.0f * k * tot) / n);
    if (cur > G) cur = G;
    double prev;
    do {
        prev = cur;
        const double got = ((uint64_t)G * prev) + (prev * prev) / 2;
        cur = (prev * want) / got;
    } while (labs((long)(((int64_t)prev - (int64_t)cur))) > 1);
    return (uint32_t)prev;
}
}  // namespace

void node_sweep_range(uint32_t genome_len, uint32_t num_chroms, bool watson_only, int num_nodes, int node,
                      uint32_t &sweep_start, uint32_t &sweep_end) {
    uint32_t l1, l2;
    if (!watson_only) {
        l1 = slice_boundary(genome_len, node, num_nodes);
        l2 = slice_boundary(genome_len, node - 1, num_nodes);
    } else {
        const uint64_t area = ((uint64_t)genome_len * (uint64_t)genome_len) / 2;
        l1 = (uint32_t)sqrt((2.0f * node * area) / num_nodes);
        l2 = (uint32_t)sqrt((2.0f * (node - 1) * area) / num_nodes) - 1;
    }
    uint32_t ss = genome_len - l1, se = genome_len - l2;
    // allowance for the inter-chromosome markers plus a safety margin so slices overlap
    if (ss > 2 * num_chroms) ss -= 2 * num_chroms;
    else ss = 0;
    se += 10 + 2 * num_chroms;
    if (se > genome_len) se = genome_len;
    if (ss > 10) ss -= 10;
    else ss = 1;
    sweep_start = ss;
    sweep_end = se;
}

void single_sweep_range(uint32_t genome_len, uint32_t b, uint32_t B, uint32_t &sweep_start, uint32_t &sweep_end) {
    sweep_start = b > genome_len ? genome_len : b;
    sweep_end = B == 0 ? genome_len : (B > genome_len ? genome_len : B);
}

}  // namespace k4bhost
