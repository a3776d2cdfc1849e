// k4b_reports.cpp - output writers of the `hammings` drop-in.  Byte-exact parity with the
// reference lives here, so each loop walks the same flat arrays in the same order as the
// reference and keeps its quirks (a9, a16).
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <iterator>
#include <thread>
#include <vector>

#include "k4b_host.h"

namespace k4bhost {

namespace {

// buffered fd writer; retries short writes like CUtility::RetryWrites (libkit4b/Utility.cpp:713)
class Out {
  public:
    int open_trunc(const std::string &path, std::string &err) {
        fd_ = ::open(path.c_str(), O_RDWR | O_CREAT, S_IRUSR | S_IWUSR);
        if (fd_ < 0 || ftruncate(fd_, 0) != 0) {
            err = "unable to create/truncate output file '" + path + "': " + strerror(errno);
            if (fd_ >= 0) ::close(fd_);
            fd_ = -1;
            return kErrCreateFile;
        }
        buf_.reserve(1 << 20);
        return kOk;
    }
    void put(const char *s, size_t n) {
        buf_.append(s, n);
        if (buf_.size() >= (1 << 20) - 4096) flush();
    }
    void put(const std::string &s) { put(s.data(), s.size()); }
    void put_big(const std::string &s) {  // large block: straight to the fd
        flush();
        size_t off = 0;
        while (off < s.size()) {
            const ssize_t w = ::write(fd_, s.data() + off, s.size() - off);
            if (w < 0) {
                if (errno == EINTR || errno == EAGAIN) continue;
                ok_ = false;
                break;
            }
            off += (size_t)w;
        }
    }
    bool flush() {
        size_t off = 0;
        while (off < buf_.size()) {
            const ssize_t w = ::write(fd_, buf_.data() + off, buf_.size() - off);
            if (w < 0) {
                if (errno == EINTR || errno == EAGAIN) continue;
                ok_ = false;
                break;
            }
            off += (size_t)w;
        }
        buf_.clear();
        return ok_;
    }
    int close_sync(std::string &err) {
        flush();
        if (fd_ >= 0) {
            fsync(fd_);
            ::close(fd_);
            fd_ = -1;
        }
        if (!ok_) {
            err = "write failed";
            return kErrFileAccess;
        }
        return kOk;
    }

  private:
    int fd_ = -1;
    bool ok_ = true;
    std::string buf_;
};

// founder prefix filter: name must start (case-insensitively) with <prefix> followed by '|' and
// '#'; the tag is stripped from printed names (hammings.cpp:1733-1742, :1805-1809,
// seghaplotypes.h:3-5)
struct PrefixFilter {
    std::string tag;
    explicit PrefixFilter(const std::string &prefix) {
        if (!prefix.empty()) tag = prefix + "|#";
    }
    bool skip(const std::string &name) const {
        return !tag.empty() && strncasecmp(name.c_str(), tag.c_str(), tag.size()) != 0;
    }
    const char *printed(const std::string &name) const {
        // the reference prints &szChromName[PrefixLen]; a shorter name would have been skipped
        return name.c_str() + tag.size();
    }
};

}  // namespace

// decimal digits of v at p, returns the end
static inline char *put_u32(char *p, uint32_t v) {
    char tmp[10];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// One run of consecutive lines of the exhaustive CSV: chromosome ci, flat positions
// seq_idx .. seq_idx+count-1, loci first_loci ..
struct CsvRun {
    size_t ci;
    uint64_t seq_idx;
    uint32_t first_loci;
    uint64_t count;
};

int write_exhaustive_csv(const std::string &path, const Genome &g, uint32_t K, const uint16_t *hd,
                         uint32_t sweep_start, uint32_t sweep_end, std::string &err) {
    Out out;
    int rc = out.open_trunc(path, err);
    if (rc) return rc;
    char line[256];
    int n = snprintf(line, sizeof(line), "%u,%d,%d\n", g.genome_len, (int)(sweep_start + 1), (int)sweep_end);
    out.put(line, (size_t)n);
    const uint64_t flat = g.genome_len - 2;
    // The reference walks the flat array with (SeqIdx, CurLoci, chromosome) state
    // (hammings.cpp:2899-2929): when CurLoci reaches the chromosome's NumSubSeqs it moves to the
    // next chromosome and skips K positions (its K-1 tail positions + the separator) - which
    // over-skips after a chromosome shorter than K, then reads mis-aligned values, and emits one
    // line for a chromosome without any K-mer.  The same walk is first done run by run (cheap),
    // then the runs are formatted by several threads and written in order: same bytes.
    std::vector<CsvRun> runs;
    if (!g.chroms.empty()) {
        size_t ci = 0;
        uint64_t seq_idx = 0;
        uint32_t cur_loci = 0;
        while (seq_idx < flat) {
            if (cur_loci >= g.chroms[ci].num_subseqs) {
                if (ci + 1 == g.chroms.size()) break;
                ++ci;
                cur_loci = 0;
                seq_idx += K;
            }
            // this iteration plus the following ones that stay inside the chromosome
            const uint64_t ns = g.chroms[ci].num_subseqs;
            const uint64_t run = ns > (uint64_t)cur_loci + 1 ? ns - cur_loci : 1;
            const uint64_t valid = seq_idx < flat ? std::min(run, flat - seq_idx) : 0;  // past the end: K+1, no line
            if (valid) runs.push_back(CsvRun{ci, seq_idx, cur_loci, valid});
            seq_idx += run;
            cur_loci += (uint32_t)run;
        }
    }
    // tasks of at most kTaskLines lines, formatted kWave at a time
    constexpr uint64_t kTaskLines = 1u << 20;
    std::vector<CsvRun> tasks;
    for (const CsvRun &r : runs)
        for (uint64_t o = 0; o < r.count; o += kTaskLines)
            tasks.push_back(CsvRun{r.ci, r.seq_idx + o, (uint32_t)(r.first_loci + o), std::min(kTaskLines, r.count - o)});
    unsigned hw = std::thread::hardware_concurrency();
    const size_t wave = std::max<size_t>(1, std::min<size_t>(hw ? hw : 4, 16));
    std::vector<std::string> bufs(wave);
    auto format = [&](const CsvRun &t, std::string &buf) {
        const std::string &name = g.chroms[t.ci].name;
        const size_t per_line = name.size() + 3 + 10 + 1 + 10 + 1;
        buf.resize((size_t)t.count * per_line);
        char *p = &buf[0];
        for (uint64_t i = 0; i < t.count; ++i) {
            const uint32_t v = hd[t.seq_idx + i];
            if (v > K) continue;
            *p++ = '"';
            memcpy(p, name.data(), name.size());
            p += name.size();
            *p++ = '"';
            *p++ = ',';
            p = put_u32(p, t.first_loci + (uint32_t)i);
            *p++ = ',';
            p = put_u32(p, v);
            *p++ = '\n';
        }
        buf.resize((size_t)(p - &buf[0]));
    };
    for (size_t t0 = 0; t0 < tasks.size(); t0 += wave) {
        const size_t nt = std::min(wave, tasks.size() - t0);
        std::vector<std::thread> th;
        for (size_t j = 1; j < nt; ++j) th.emplace_back([&, j]() { format(tasks[t0 + j], bufs[j]); });
        format(tasks[t0], bufs[0]);
        for (std::thread &x : th) x.join();
        for (size_t j = 0; j < nt; ++j) out.put_big(bufs[j]);
    }
    return out.close_sync(err);
}

int write_restricted_csv(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K,
                         const uint8_t *h, const std::string &prefix, std::string &err) {
    Out out;
    int rc = out.open_trunc(path, err);
    if (rc) return rc;
    const PrefixFilter pf(prefix);
    out.put("\"Chrom\",\"StartLoci\",\"Len\",\"Hamming\"\n");
    char line[512];
    uint64_t ofs = 0;
    for (const RChrom &c : chroms) {
        const uint8_t *p = h + ofs;
        ofs += c.len;
        if (c.len < K || pf.skip(c.name)) continue;
        // run-length walk with the reference's off-by-ones: the loop starts at loci 1 while
        // still reading H[0], so the last K-mer is never examined and every run after the
        // first is reported one position late (hammings.cpp:1811-1829)
        int run_len = 0, cur = *p, run_loci = 0;
        for (uint32_t loci = 1; loci <= c.len - K; ++loci, ++p) {
            if (*p == cur) {
                ++run_len;
                continue;
            }
            int n = snprintf(line, sizeof(line), "\"%s\",%d,%d,%d\n", pf.printed(c.name), run_loci, run_len, cur);
            out.put(line, (size_t)n);
            run_len = 1;
            cur = *p;
            run_loci = (int)loci;
        }
        int n = snprintf(line, sizeof(line), "\"%s\",%d,%d,%d\n", pf.printed(c.name), run_loci, run_len, cur);
        out.put(line, (size_t)n);
    }
    return out.close_sync(err);
}

int write_restricted_bed(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K, int R,
                         const uint8_t *h, const std::string &prefix, std::string &err) {
    Out out;
    int rc = out.open_trunc(path, err);
    if (rc) return rc;
    const PrefixFilter pf(prefix);
    char line[512];
    int n = snprintf(line, sizeof(line),
                     "track type=bedGraph name=ResHamming%d_%d description=\"Restricted Hammings for K-mer length "
                     "%d and Hamming limit %d\"\n",
                     (int)K, R, (int)K, R);
    out.put(line, (size_t)n);
    uint64_t ofs = 0;
    for (const RChrom &c : chroms) {
        const uint8_t *p = h + ofs;
        ofs += c.len;
        if (c.len < K || pf.skip(c.name)) continue;
        // same walk as the CSV but the run length starts at 1 (hammings.cpp:1945-1965)
        int run_len = 1, cur = *p, run_loci = 0;
        for (uint32_t loci = 1; loci <= c.len - K; ++loci, ++p) {
            if (*p == cur) {
                ++run_len;
                continue;
            }
            n = snprintf(line, sizeof(line), "%s\t%d\t%d\t%d\n", pf.printed(c.name), run_loci, run_loci + run_len - 1, cur);
            out.put(line, (size_t)n);
            run_len = 1;
            cur = *p;
            run_loci = (int)loci;
        }
        n = snprintf(line, sizeof(line), "%s\t%d\t%d\t%d\n", pf.printed(c.name), run_loci, run_loci + run_len - 1, cur);
        out.put(line, (size_t)n);
    }
    return out.close_sync(err);
}

int write_restricted_wiggle(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K,
                            int R, int sensitivity, const uint8_t *h, const std::string &prefix,
                            std::string &err) {
    Out out;
    int rc = out.open_trunc(path, err);
    if (rc) return rc;
    const PrefixFilter pf(prefix);
    std::string hdr(2 * path.size() + 512, '\0');
    int n = snprintf(&hdr[0], hdr.size(),
                     "track type=wiggle_0 color=50,150,255 autoScale=off maxHeightPixels=128:32:8 name=\"Hammings "
                     "(%d,%d,%d) - %s\" description=\"Hammings Sensitivity: %d KMerLen: %d RHamm: %d  for %s\"\n",
                     sensitivity, (int)K, R, path.c_str(), sensitivity, (int)K, R, path.c_str());
    out.put(hdr.data(), (size_t)n);
    char line[512];
    uint64_t ofs = 0;
    for (const RChrom &c : chroms) {
        const uint8_t *p = h + ofs;
        ofs += c.len;
        if (c.len < K || pf.skip(c.name)) continue;
        // correct loop bounds here, but after a change the span start is loci+1 so every span
        // but the first is printed one position late (hammings.cpp:2084-2103)
        uint32_t span_len = 0, span_start = 0, loci = 0;
        uint8_t cur = *p;
        for (loci = 0; loci <= c.len - K; ++loci, ++p) {
            if (*p != cur) {
                n = snprintf(line, sizeof(line), "variableStep chrom=%s span=%d\n%d %d\n", pf.printed(c.name),
                             (int)span_len, (int)(span_start + 1), (int)cur);
                out.put(line, (size_t)n);
                cur = *p;
                span_len = 0;
                span_start = loci + 1;
            }
            ++span_len;
        }
        if (span_start != loci) {
            n = snprintf(line, sizeof(line), "variableStep chrom=%s span=%d\n%d %d\n", pf.printed(c.name),
                         (int)span_len, (int)(span_start + 1), (int)cur);
            out.put(line, (size_t)n);
        }
    }
    return out.close_sync(err);
}

// ---- -m3 merge -----------------------------------------------------------------------------------
namespace {
struct CsvRow {
    bool ok = false;        // "chrom",loci,dist shape
    std::string chrom;
    long loci = 0, dist = 0;
    int nfields = 0;
};
// minimal CSV line split: the reference only accepts rows whose first field is quoted and whose
// second and third are not (hammings.cpp:1243-1248)
CsvRow parse_row(const std::string &line) {
    CsvRow r;
    std::vector<std::string> f;
    std::vector<bool> quoted;
    size_t i = 0;
    while (i <= line.size()) {
        std::string cur;
        bool q = false;
        while (i < line.size() && (line[i] == ' ' || line[i] == '\t')) ++i;
        if (i < line.size() && line[i] == '"') {
            q = true;
            ++i;
            while (i < line.size() && line[i] != '"') cur.push_back(line[i++]);
            if (i < line.size()) ++i;
            while (i < line.size() && line[i] != ',') ++i;
        } else {
            while (i < line.size() && line[i] != ',') cur.push_back(line[i++]);
        }
        f.push_back(cur);
        quoted.push_back(q);
        if (i >= line.size()) break;
        ++i;  // skip the comma
    }
    r.nfields = (int)f.size();
    if (f.size() >= 3 && quoted[0] && !quoted[1] && !quoted[2]) {
        r.ok = true;
        r.chrom = f[0];
        r.loci = strtol(f[1].c_str(), nullptr, 10);
        r.dist = strtol(f[2].c_str(), nullptr, 10);
    }
    return r;
}
bool next_line(std::ifstream &in, std::string &line) {
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        size_t b = 0;
        while (b < line.size() && isspace((unsigned char)line[b])) ++b;
        if (b == line.size()) continue;  // blank lines are skipped by the CSV reader
        return true;
    }
    return false;
}
}  // namespace

int merge_hamming_csv(const std::string &from, const std::string &into, std::string &err) {
    std::ifstream fin(from);
    if (!fin) {
        err = "unable to open '" + from + "'";
        return kErrOpnFile;
    }
    std::ifstream tin(into);
    const bool copy = !tin.good();  // no existing target: the result is a copy of `from`
    const std::string tmp = copy ? into : into + ".tmp";
    Out out;
    int rc = out.open_trunc(tmp, err);
    if (rc) return rc;
    out.put("\"Chrom\",\"Loci\",\"Hamming\"");
    std::string lf, lt;
    int num_errs = 0;
    bool fail = false;
    char line[512];
    while (true) {
        if (!next_line(fin, lf)) break;
        if (!copy && !next_line(tin, lt)) break;
        CsvRow a = parse_row(lf), b;
        if (a.nfields < 3) {
            fail = true;
            break;
        }
        if (!copy) {
            b = parse_row(lt);
            if (b.nfields < 3) {
                fail = true;
                break;
            }
        }
        if (!a.ok) continue;              // descriptor rows (e.g. the "G,2,G" header) are dropped
        if (!copy && !b.ok) continue;
        long d = a.dist;
        if (!copy) {
            if (strcasecmp(a.chrom.c_str(), b.chrom.c_str()) != 0 || a.loci != b.loci) {
                if (num_errs++ > 1) {
                    fail = true;
                    break;
                }
                continue;
            }
            d = std::min(a.dist, b.dist);
        }
        int n = snprintf(line, sizeof(line), "\n\"%s\",%ld,%ld", a.chrom.c_str(), a.loci, d);
        out.put(line, (size_t)n);
    }
    rc = out.close_sync(err);
    if (fail) {
        err = "merge inputs are inconsistent";
        return -1;
    }
    if (rc) return rc;
    if (!copy) {
        tin.close();
        if (remove(into.c_str()) != 0 || rename(tmp.c_str(), into.c_str()) != 0) {
            err = "unable to replace '" + into + "'";
            return -1;
        }
    }
    return kOk;
}


// ---- -m4 / -m5 quick-load binary -----------------------------------------------------------------
namespace {
constexpr size_t kBhamMaxChroms = 1000;
constexpr size_t kBhamHdrSize = 4 + 4 + 4 + 2 + 4 * kBhamMaxChroms;  // tsHHamHdr, pack(1)
constexpr size_t kBhamChromFixed = 4 + 81 + 4;                       // tsHHamChrom before Dists
}  // namespace

int csv_to_bham(const std::string &csv, const std::string &bham, std::string &err) {
    std::ifstream in(csv);
    if (!in) {
        err = "unable to open '" + csv + "'";
        return kErrOpnFile;
    }
    struct ChromBlock {
        std::string name;
        std::vector<uint16_t> dists;
    };
    std::vector<ChromBlock> blocks;
    std::string line;
    while (next_line(in, line)) {
        const CsvRow r = parse_row(line);
        if (r.nfields < 3) {
            err = "expected at least 3 fields per line in '" + csv + "'";
            return kErrParse;
        }
        if (!r.ok) continue;  // descriptor rows
        if (blocks.empty() || strcasecmp(blocks.back().name.c_str(), r.chrom.c_str()) != 0) {
            if (blocks.size() == kBhamMaxChroms) {
                err = "more than 1000 chromosomes";
                return kErrParams;
            }
            blocks.push_back(ChromBlock{r.chrom.substr(0, 80), {}});
        }
        if (r.loci != (long)blocks.back().dists.size()) {  // loci must ascend without gaps
            err = "Hamming loci are not monotonically ascending in '" + csv + "'";
            return kErrParse;
        }
        blocks.back().dists.push_back((uint16_t)r.dist);
    }
    std::vector<uint8_t> img(kBhamHdrSize, 0);
    memcpy(img.data(), "bham", 4);
    const uint32_t version = 1;
    memcpy(img.data() + 4, &version, 4);
    const uint16_t nch = (uint16_t)blocks.size();
    memcpy(img.data() + 12, &nch, 2);
    for (size_t c = 0; c < blocks.size(); ++c) {
        const uint32_t ofs = (uint32_t)img.size();
        memcpy(img.data() + 14 + 4 * c, &ofs, 4);
        const uint32_t id = (uint32_t)c + 1, nels = (uint32_t)blocks[c].dists.size();
        img.resize(img.size() + kBhamChromFixed + 2 * (size_t)nels, 0);
        uint8_t *p = img.data() + ofs;
        memcpy(p, &id, 4);
        memcpy(p + 4, blocks[c].name.c_str(), blocks[c].name.size());
        memcpy(p + 4 + 81, &nels, 4);
        if (nels) memcpy(p + kBhamChromFixed, blocks[c].dists.data(), 2 * (size_t)nels);
    }
    if (img.size() > 0x7fffffffu) {
        err = "Hammings do not fit the 31-bit length field of the bham header";
        return kErrParams;
    }
    const int32_t len = (int32_t)img.size();
    memcpy(img.data() + 8, &len, 4);
    Out out;
    int rc = out.open_trunc(bham, err);
    if (rc) return rc;
    out.put((const char *)img.data(), img.size());
    return out.close_sync(err);
}

int bham_to_csv(const std::string &bham, const std::string &csv, std::string &err) {
    std::ifstream in(bham, std::ios::binary);
    if (!in) {
        err = "unable to open '" + bham + "'";
        return kErrOpnFile;
    }
    std::vector<uint8_t> img((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (img.size() < kBhamHdrSize) {
        err = "'" + bham + "' is too short to be a bham file";
        return kErrParse;
    }
    if (memcmp(img.data(), "bham", 4) != 0) {
        err = "'" + bham + "' is not a bham file";
        return kErrFileType;
    }
    uint16_t nch;
    memcpy(&nch, img.data() + 12, 2);
    if (nch < 1 || nch > kBhamMaxChroms) {
        err = "bham file holds no chromosomes";
        return -59;  // eBSFerrNoEntries
    }
    Out out;
    int rc = out.open_trunc(csv, err);
    if (rc) return rc;
    out.put("\"Chrom\",\"Loci\",\"Hamming\"\n");
    char line[256];
    for (uint16_t c = 0; c < nch; ++c) {
        uint32_t ofs, nels;
        memcpy(&ofs, img.data() + 14 + 4 * (size_t)c, 4);
        if ((size_t)ofs + kBhamChromFixed > img.size()) {
            err = "corrupt bham chromosome offset";
            return kErrParse;
        }
        memcpy(&nels, img.data() + ofs + 4 + 81, 4);
        if ((size_t)ofs + kBhamChromFixed + 2 * (size_t)nels > img.size()) {
            err = "bham chromosome data lies outside the file";
            return kErrParse;
        }
        const char *name = (const char *)img.data() + ofs + 4;
        const std::string nm(name, strnlen(name, 81));
        for (uint32_t l = 0; l < nels; ++l) {
            uint16_t d;
            memcpy(&d, img.data() + ofs + kBhamChromFixed + 2 * (size_t)l, 2);
            const int n = snprintf(line, sizeof(line), "\"%s\",%d,%d\n", nm.c_str(), (int)l, (int)d);
            out.put(line, (size_t)n);
        }
    }
    return out.close_sync(err);
}

// ---- HammingDist, region-less mode (HammingDist/HammingDist.cpp:371-705) ------------------------------
int write_hamming_distribution(const std::string &path, const std::vector<uint64_t> &counts, std::string &err) {
    long maxh = -1;
    for (size_t d = 0; d < counts.size(); ++d)
        if (counts[d]) maxh = (long)d;
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) {
        err = "Unable to create " + path + " - " + strerror(errno);
        return kErrCreateFile;
    }
    if (maxh >= 0) {  // "any to report?" (:616)
        fputs(",\"All\",\"Proportion All\",\"Cumulative All\"", f);
        uint64_t total = 0;
        for (long d = 0; d < maxh; ++d) total += counts[d];  // the reference's loops exclude the largest value (:631)
        double cumulative = 0.0;
        for (long d = 0; d < maxh; ++d) {
            const double prop = total ? (double)counts[d] / (double)total : 0.0;
            cumulative += prop;
            fprintf(f, "\n%ld,%llu,%f,%f", d, (unsigned long long)counts[d], prop, cumulative);
        }
    }
    const bool ok = fflush(f) == 0 && fsync(fileno(f)) == 0;
    fclose(f);
    if (!ok) {
        err = "Error on write to file '" + path + "'";
        return kErrFileAccess;
    }
    return kOk;
}

namespace {
// one CSV line -> fields (double quotes stripped; quoted[i] tells whether field i was quoted)
void split_csv(const std::string &line, std::vector<std::string> &fields, std::vector<bool> &quoted) {
    fields.clear();
    quoted.clear();
    size_t i = 0;
    const size_t n = line.size();
    while (i <= n) {
        while (i < n && (line[i] == ' ' || line[i] == '\t')) ++i;
        std::string v;
        bool q = false;
        if (i < n && line[i] == '"') {
            q = true;
            ++i;
            while (i < n && line[i] != '"') v.push_back(line[i++]);
            while (i < n && line[i] != ',') ++i;
        } else {
            while (i < n && line[i] != ',') v.push_back(line[i++]);
            while (!v.empty() && (v.back() == ' ' || v.back() == '\r' || v.back() == '\t')) v.pop_back();
        }
        fields.push_back(v);
        quoted.push_back(q);
        if (i >= n) break;
        ++i;  // the comma
        if (i == n) {  // trailing comma: one more empty field
            fields.emplace_back();
            quoted.push_back(false);
            break;
        }
    }
}
bool is_number(const std::string &s) {
    if (s.empty()) return false;
    char *end = nullptr;
    strtod(s.c_str(), &end);
    return end && *end == '\0';
}
}  // namespace

int hamming_counts_from_csv(const std::vector<std::string> &files, std::vector<uint64_t> &counts, uint64_t &rows,
                            std::string &err) {
    counts.clear();
    rows = 0;
    std::vector<std::string> fields;
    std::vector<bool> quoted;
    for (const std::string &path : files) {
        std::ifstream in(path);
        if (!in) {
            err = "Unable to open file: " + path;
            return kErrOpnFile;
        }
        std::string line;
        uint64_t processed = 0, lineno = 0;
        while (std::getline(in, line)) {
            ++lineno;
            if (line.empty() || line == "\r") continue;
            split_csv(line, fields, quoted);
            if (fields.size() < 3) {
                err = "Expected 3+ fields in '" + path + "', line " + std::to_string(lineno);
                return kErrParse;
            }
            if (!processed) {
                // CCSVFile::IsLikelyHeaderLine (CSVFile.cpp:757-789): every field quoted or non-numeric, <= 2 empty
                bool header = true;
                int empty = 0;
                for (size_t k = 0; k < fields.size() && header; ++k) {
                    if (quoted[k]) continue;
                    if (fields[k].empty()) {
                        if (++empty > 2) header = false;
                        continue;
                    }
                    if (is_number(fields[k])) header = false;
                }
                if (header) continue;
                // the `m_GenomeLen,SSeqStart,SSeqEnd` descriptor row of -m1 output (hammings.cpp:2899): no chromosome name
                if (!quoted[0] && is_number(fields[0])) continue;
            }
            ++processed;
            const long h = atol(fields[2].c_str());
            if (h < 0 || h > 65535) {
                err = "Hamming distance " + std::to_string(h) + " out of range in '" + path + "', line " + std::to_string(lineno);
                return kErrParse;
            }
            if ((size_t)h >= counts.size()) counts.resize((size_t)h + 1, 0);
            ++counts[(size_t)h];
        }
        rows += processed;
    }
    return kOk;
}

// ---- HammingDist, region mode (HammingDist/HammingDist.cpp:371-496, :606-700) --------------------------
int region_counts_from_csv(const std::vector<std::string> &files, const FeatureSet &fs, int ofs_loci, int reg_len,
                           RegionHistogram &h, std::string &log, std::string &err) {
    std::vector<std::string> fields;
    std::vector<bool> quoted;
    for (const std::string &path : files) {
        std::ifstream in(path);
        if (!in) {
            err = "Unable to open file: " + path;
            return kErrOpnFile;
        }
        std::string line, prev_chrom;
        long processed = 0, lineno = 0;
        int max_h = -1, chrom = -1;
        bool complete = true;  // read to its end: only then do its rows and its maximum count (:490-495)
        while (std::getline(in, line)) {
            ++lineno;
            size_t b = 0;
            while (b < line.size() && (line[b] == ' ' || line[b] == '\t' || line[b] == '\r')) ++b;
            if (b == line.size() || line[b] == '#') continue;  // blank and comment lines (CSVFile.cpp:655-687)
            split_csv(line, fields, quoted);
            if (fields.size() < 3) {
                log += "Expected 3+ fields in '" + path + "', line " + std::to_string(lineno) + ": reading of this file stops\n";
                complete = false;
                break;
            }
            if (!processed) {  // CCSVFile::IsLikelyHeaderLine (CSVFile.cpp:757-789)
                bool header = true;
                int empty = 0;
                for (size_t k = 0; k < fields.size() && header; ++k) {
                    if (quoted[k]) continue;
                    if (fields[k].empty()) {
                        if (++empty > 2) header = false;
                        continue;
                    }
                    if (is_number(fields[k])) header = false;
                }
                if (header) continue;
            }
            ++processed;
            int loci = atoi(fields[1].c_str());
            const int hamming = atoi(fields[2].c_str());
            if (chrom < 0 || fields[0] != prev_chrom) {
                chrom = fs.chrom_id(fields[0]);
                if (chrom < 0) {  // the reference leaves the file at the first unknown chromosome (:449-453)
                    log += "'" + path + "' line " + std::to_string(lineno) + ": chromosome '" + fields[0] +
                           "' is not in the feature file; as in the reference, reading of this file stops here and its rows do not "
                           "count towards the totals\n";
                    complete = false;
                    break;
                }
                prev_chrom = fields[0];
            }
            if (hamming < 0 || hamming > kMaxRegionHamming) {
                err = "Hamming distance " + std::to_string(hamming) + " in '" + path + "', line " + std::to_string(lineno) +
                      " does not fit the region table (0.." + std::to_string(kMaxRegionHamming) + ")";
                return kErrParse;
            }
            loci += ofs_loci;
            if (loci < 0) loci = 0;
            int bits = fs.feature_bits(chrom, loci, loci, kFeatRegionBits, reg_len);
            int region = 0;
            for (; region < kNumRegions - 1; ++region, bits >>= 1)
                if (bits & 1) break;
            if (max_h < hamming) max_h = hamming;
            ++h.counts[region][hamming];
        }
        if (complete) {
            h.total_processed += processed;
            if (h.max_hamming < max_h) h.max_hamming = max_h;
        }
    }
    return kOk;
}

int write_region_distribution(const std::string &path, const RegionHistogram &h, std::string &err) {
    static const char *const kNames[kNumRegions] = {"CDS", "UTR5", "UTR3", "Intron", "UP5", "DN3", "Intergenic"};  // :313-338
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) {
        err = "Unable to create " + path + " - " + strerror(errno);
        return kErrCreateFile;
    }
    if (h.max_hamming >= 0 && h.total_processed > 0) {
        for (const char *pre : {"", "Proportion ", "Cumulative "})
            for (int r = 0; r < kNumRegions; ++r) fprintf(f, ",\"%s%s\"", pre, kNames[r]);
        uint32_t total[kNumRegions] = {};
        double cumulative[kNumRegions] = {};
        for (int d = 0; d < h.max_hamming; ++d)  // the largest distance is left out of totals and rows alike (:631, :640)
            for (int r = 0; r < kNumRegions; ++r) total[r] += h.counts[r][d];
        for (int d = 0; d < h.max_hamming; ++d) {
            fprintf(f, "\n%d", d);
            for (int r = 0; r < kNumRegions; ++r) fprintf(f, ",%d", (int)h.counts[r][d]);
            for (int r = 0; r < kNumRegions; ++r) fprintf(f, ",%f", total[r] > 0 ? h.counts[r][d] / (double)total[r] : 0.0);
            for (int r = 0; r < kNumRegions; ++r) {
                cumulative[r] += total[r] > 0 ? (double)h.counts[r][d] / (double)total[r] : 0.0;
                fprintf(f, ",%f", cumulative[r]);
            }
        }
    }
    const bool ok = fflush(f) == 0 && fsync(fileno(f)) == 0;
    fclose(f);
    if (!ok) {
        err = "Error on write to file '" + path + "'";
        return kErrFileAccess;
    }
    return kOk;
}

}  // namespace k4bhost
