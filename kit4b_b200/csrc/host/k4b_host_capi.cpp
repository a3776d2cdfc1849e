// k4b_host_capi.cpp - extern "C" view of the host-side pieces (containers, genome layout,
// writers, CLI parsing) so that they can be driven without a GPU (tests bind it with ctypes;
// other front ends can reuse the readers/writers).  No numeric work happens here.
#include <string.h>

#include "k4b_host.h"

using namespace k4bhost;

namespace {
thread_local std::string g_err;
int set_err(int rc, const std::string &e) {
    g_err = e;
    return rc;
}
int load_genome(const char *bioseq_path, uint32_t K, Genome &g) {
    std::vector<SeqEntry> entries;
    std::string title, err;
    int rc = read_bioseq(bioseq_path, entries, title, err);
    if (rc) return set_err(rc, err);
    build_genome(entries, K, g);
    return 0;
}
}  // namespace

extern "C" {

const char *k4bh_last_error(void) { return g_err.c_str(); }

// concat layout of a bioseq file (hammings.cpp:2981-3134); returns its length or <0.
// out may be NULL to query the size.  names: '\n'-joined "name\tstart\tlen" lines.
long k4bh_concat_from_bioseq(const char *bioseq_path, uint32_t K, uint8_t *out, size_t cap, char *names,
                             size_t names_cap, uint32_t *genome_len) {
    Genome g;
    int rc = load_genome(bioseq_path, K, g);
    if (rc) return rc;
    if (genome_len) *genome_len = g.genome_len;
    if (out) {
        if (cap < g.concat.size()) return set_err(kErrParams, "buffer too small");
        memcpy(out, g.concat.data(), g.concat.size());
    }
    if (names) {
        std::string s;
        for (const Chrom &c : g.chroms) s += c.name + "\t" + std::to_string(c.start) + "\t" + std::to_string(c.len) + "\n";
        if (s.size() + 1 > names_cap) return set_err(kErrParams, "names buffer too small");
        memcpy(names, s.c_str(), s.size() + 1);
    }
    return (long)g.concat.size();
}

int k4bh_write_exhaustive_csv(const char *bioseq_path, uint32_t K, const uint16_t *hd, size_t hd_len,
                              uint32_t sweep_start, uint32_t sweep_end, const char *out_path) {
    Genome g;
    int rc = load_genome(bioseq_path, K, g);
    if (rc) return rc;
    if (hd_len != g.concat.size()) return set_err(kErrParams, "result array length differs from the genome");
    std::string err;
    if (sweep_end == 0) sweep_end = g.genome_len;
    rc = write_exhaustive_csv(out_path, g, K, hd, sweep_start, sweep_end, err);
    return rc ? set_err(rc, err) : 0;
}

// h: per-loci array over the probe entries (sum of entry lengths); fmt 0 csv, 1 bed, 2 wiggle
int k4bh_write_restricted(const char *probe_bioseq_path, uint32_t K, int R, int sensitivity, int fmt,
                          const char *prefix, const uint8_t *h, size_t h_len, const char *out_path) {
    Genome g;
    int rc = load_genome(probe_bioseq_path, K, g);
    if (rc) return rc;
    if (h_len != g.total_bases) return set_err(kErrParams, "per-loci array length differs from the probe set");
    std::vector<RChrom> rch;
    for (const Chrom &c : g.chroms) rch.push_back(RChrom{c.name, c.len});
    std::string err, pfx = prefix ? prefix : "";
    if (fmt == 0) rc = write_restricted_csv(out_path, rch, K, h, pfx, err);
    else if (fmt == 1) rc = write_restricted_bed(out_path, rch, K, R, h, pfx, err);
    else rc = write_restricted_wiggle(out_path, rch, K, R, sensitivity, h, pfx, err);
    return rc ? set_err(rc, err) : 0;
}

// sequence area of an sfx5 file; returns its length or <0. entries: "name\tstart\tlen" lines
long k4bh_read_sfx(const char *path, uint8_t *out, size_t cap, char *entries, size_t entries_cap) {
    SfxData d;
    std::string err;
    int rc = read_sfx(path, d, err);
    if (rc) return set_err(rc, err);
    if (out) {
        if (cap < d.seq.size()) return set_err(kErrParams, "buffer too small");
        memcpy(out, d.seq.data(), d.seq.size());
    }
    if (entries) {
        std::string s;
        for (const SfxEntry &e : d.entries) s += e.name + "\t" + std::to_string(e.start) + "\t" + std::to_string(e.len) + "\n";
        if (s.size() + 1 > entries_cap) return set_err(kErrParams, "entries buffer too small");
        memcpy(entries, s.c_str(), s.size() + 1);
    }
    return (long)d.seq.size();
}

int k4bh_fasta_to_bioseq(const char *fasta_path, const char *bioseq_path, const char *title) {
    std::vector<SeqEntry> entries;
    std::string err;
    int rc = read_fasta(fasta_path, entries, err);
    if (rc) return set_err(rc, err);
    rc = write_bioseq(bioseq_path, entries, title ? title : "", err);
    return rc ? set_err(rc, err) : 0;
}

// sweep range of a -m2 node slice (num_nodes, node) or of -m1 -b/-B (num_nodes == 0: b, B)
void k4bh_sweep_range(uint32_t genome_len, uint32_t num_chroms, int watson_only, int num_nodes, int node_or_b,
                      uint32_t B, uint32_t *ss, uint32_t *se) {
    if (num_nodes > 0) node_sweep_range(genome_len, num_chroms, watson_only != 0, num_nodes, node_or_b, *ss, *se);
    else single_sweep_range(genome_len, (uint32_t)node_or_b, B, *ss, *se);
}

int k4bh_merge_csv(const char *from, const char *into) {
    std::string err;
    int rc = merge_hamming_csv(from, into, err);
    return rc ? set_err(rc, err) : 0;
}

int k4bh_csv_to_bham(const char *csv, const char *bham) {
    std::string err;
    int rc = csv_to_bham(csv, bham, err);
    return rc ? set_err(rc, err) : 0;
}
int k4bh_bham_to_csv(const char *bham, const char *csv) {
    std::string err;
    int rc = bham_to_csv(bham, csv, err);
    return rc ? set_err(rc, err) : 0;
}

// HammingDist region-less mode: distribution file of the `"chrom",loci,hamming` rows of n CSV files
int k4bh_hamming_dist(int n, const char **csvs, const char *out) {
    std::vector<std::string> files(csvs, csvs + n);
    std::vector<uint64_t> counts;
    uint64_t rows = 0;
    std::string err;
    int rc = hamming_counts_from_csv(files, counts, rows, err);
    if (!rc) rc = write_hamming_distribution(out, counts, err);
    return rc ? set_err(rc, err) : 0;
}

// HammingDist region mode (-I): per-region distribution file; feats = BED text or biobed container.
// The region of one locus is available on its own for tests: -1 = chromosome not in the feature file.
int k4bh_hamming_dist_regions(int n, const char **csvs, const char *feats, int reg_len, int ofs_loci, const char *out) {
    std::vector<std::string> files(csvs, csvs + n);
    std::string err, log;
    FeatureSet fs;
    int rc = read_features(feats, fs, err);
    RegionHistogram hist;
    if (!rc) rc = region_counts_from_csv(files, fs, ofs_loci, reg_len, hist, log, err);
    if (!rc) rc = write_region_distribution(out, hist, err);
    return rc ? set_err(rc, err) : 0;
}
int k4bh_feature_bits(const char *feats, const char *chrom, int n, const int *loci, int reg_len, int *bits) {
    std::string err;
    FeatureSet fs;
    const int rc = read_features(feats, fs, err);
    if (rc) return set_err(rc, err);
    const int id = fs.chrom_id(chrom);
    for (int i = 0; i < n; ++i) bits[i] = id < 0 ? -1 : fs.feature_bits(id, loci[i], loci[i], kFeatRegionBits, reg_len);
    return 0;
}

// parses a command line (argv[0] = program); fills ints[0..15] and strs (4 x 512 chars:
// in, inseq, out, prefix).  Returns 0, or -1 with k4bh_last_error() set.
int k4bh_parse_cli(int argc, char **argv, int *ints, char *strs) {
    std::vector<std::string> expanded;
    std::string err;
    if (expand_param_files(argc, argv, expanded, err) < 0) return set_err(-1, err);
    std::vector<char *> pv;
    for (std::string &s : expanded) pv.push_back(&s[0]);
    pv.push_back(nullptr);
    Options o;
    if (parse_args((int)expanded.size(), pv.data(), o, err)) return set_err(-1, err);
    const int v[16] = {o.mode, o.sensitivity, o.resformat, o.crick, o.intrainterboth, o.rhamm, o.numnodes, o.node,
                       o.sweep_start, o.sweep_end, o.K, o.sample, o.threads, o.gpus, o.help, o.version};
    memcpy(ints, v, sizeof(v));
    const std::string *ss[4] = {&o.in_file, &o.in_seq_file, &o.out_file, &o.prefix};
    for (int i = 0; i < 4; ++i) {
        strncpy(strs + 512 * i, ss[i]->c_str(), 511);
        strs[512 * i + 511] = 0;
    }
    return 0;
}

}  // extern "C"
