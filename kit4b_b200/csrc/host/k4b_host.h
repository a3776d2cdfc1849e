// k4b_host.h - host side of the `hammings` drop-in: input containers, the concatenated
// genome layout, report writers and the CLI.  Everything here mirrors behaviour of the
// reference's ngskit4b/hammings.cpp + libkit4b readers for this one subprocess; the numeric
// work is behind the C ABI in include/k4b_hamm.h.
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace k4bhost {

// result codes follow libkit4b/ErrorCodes.h:15-33
enum : int {
    kOk = 0,
    kErrParams = -100,
    kErrMem = -95,
    kErrNotBioseq = -94,
    kErrFileType = -92,
    kErrOpnFile = -90,
    kErrCreateFile = -89,
    kErrFileVer = -86,
    kErrFileAccess = -85,
    kErrNoFeatures = -60,
    kErrFeature = -53,
    kErrParse = -46,
};

struct SeqEntry {
    std::string name;            // <= 80 chars (commdefs.h:140)
    std::vector<uint8_t> codes;  // one base code per byte, soft-mask flag (0x08) still present
};

// ---- containers -------------------------------------------------------------------------
// 'bios' container, type 1 / version 10 (libkit4b/BioSeqFile.h:28-58, BioSeqFile.cpp:378-470,
// :928-1045, :1174-1220); entries come back in EntryID order (= FASTA order)
int read_bioseq(const std::string &path, std::vector<SeqEntry> &entries, std::string &title,
                std::string &err);
// minimal writer of the same container (used by the FASTA front end and by tests)
int write_bioseq(const std::string &path, const std::vector<SeqEntry> &entries,
                 const std::string &title, std::string &err);
// FASTA text, plain or gzip -> entries (libkit4b/Fasta.cpp:967-1209, :1658-1705 Ascii2Sense, :1167;
// genbioseq.cpp:402-404)
int read_fasta(const std::string &path, std::vector<SeqEntry> &entries, std::string &err);
// 'b' bioseq, 's' suffix array, 'f' FASTA (plain or gzip), 0 unknown; and the reader the CLI uses for
// -i (-m1/-m2) and -I: a bioseq container or, skipping genbioseq, FASTA / FASTA.gz directly
int sniff_format(const std::string &path);
int read_sequences(const std::string &path, std::vector<SeqEntry> &entries, std::string &title, std::string &err);

// 'sfx5' suffix-array container: only the entries and the concatenated sequence are read
// (libkit4b/SfxArray.h:98-123, :194-207; SfxArray.cpp:499-580, :629-825)
struct SfxEntry {
    std::string name;
    uint32_t len = 0;
    uint64_t start = 0;  // offset into the concatenated sequence
};
struct SfxData {
    std::string dataset, descr, title;
    int version = 0;
    std::vector<SfxEntry> entries;
    std::vector<uint8_t> seq;  // ConcatSeqLen bytes: each entry's bases followed by EOS (7)
};
int read_sfx(const std::string &path, SfxData &out, std::string &err);
// what `index` would lay out for these entries (bases + EOS per entry): lets -m0 take its -i assembly
// as bioseq or FASTA, no suffix-array file needed (the GPU engines never read the suffix array)
void sfx_from_entries(const std::vector<SeqEntry> &entries, const std::string &title, SfxData &out);

// ---- concatenated layout (hammings.cpp:2981-3134 LoadGenome) -------------------------------
struct Chrom {
    std::string name;
    uint32_t len = 0;
    uint32_t start = 0;       // offset into concat (== SeqOfs-1 of the reference)
    uint32_t num_subseqs = 0;  // max(0, len-K+1)
};
struct Genome {
    std::vector<uint8_t> concat;  // chr1 EOS chr2 EOS ... chrN (no EOG sentinels), mask stripped
    std::vector<Chrom> chroms;
    uint32_t genome_len = 0;      // the reference's m_GenomeLen == concat.size()+2
    uint64_t total_bases = 0;
    uint64_t num_subseqs = 0;
};
void build_genome(const std::vector<SeqEntry> &entries, uint32_t K, Genome &g);

// ---- sweep ranges (hammings.cpp:2660-2706) ------------------------------------------------------
// -m2: the slice of sweep instances node `node` of `num_nodes` processes, widened so that
// neighbouring slices overlap; -m1: the -b/-B values clamped to the genome length
void node_sweep_range(uint32_t genome_len, uint32_t num_chroms, bool watson_only, int num_nodes, int node,
                      uint32_t &sweep_start, uint32_t &sweep_end);
void single_sweep_range(uint32_t genome_len, uint32_t b, uint32_t B, uint32_t &sweep_start, uint32_t &sweep_end);

// ---- writers ---------------------------------------------------------------------------------
// exhaustive CSV, literal restatement of hammings.cpp:2899-2929 (quirks included)
int write_exhaustive_csv(const std::string &path, const Genome &g, uint32_t K, const uint16_t *hd,
                         uint32_t sweep_start, uint32_t sweep_end, std::string &err);
// restricted reports over the per-loci array H[sum of entry lengths] (hammings.cpp:1711-2118)
struct RChrom {
    std::string name;
    uint32_t len = 0;
};
int write_restricted_csv(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K,
                         const uint8_t *h, const std::string &prefix, std::string &err);
int write_restricted_bed(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K, int R,
                         const uint8_t *h, const std::string &prefix, std::string &err);
int write_restricted_wiggle(const std::string &path, const std::vector<RChrom> &chroms, uint32_t K,
                            int R, int sensitivity, const uint8_t *h, const std::string &prefix,
                            std::string &err);
// -m3: line-wise minimum of two exhaustive CSVs (hammings.cpp:1126-1343)
int merge_hamming_csv(const std::string &from, const std::string &into, std::string &err);

// -m4 / -m5: exhaustive CSV <-> 'bham' quick-load binary (hammings.cpp:79-92, :941-1066,
// :1345-1514).  The layout (pack(1) header with magic 'bham', version 1, <= 1000 chromosome
// offsets; per chromosome id, 81-byte name, NumEls, uint16 distances) is the reference's, but
// lengths and offsets are computed correctly here: the reference advances its length by ONE
// byte per 2-byte distance (hammings.cpp:1464), so its own files are truncated, overlap
// chromosomes and carry uninitialised heap bytes - they are not reproducible even by itself.
int csv_to_bham(const std::string &csv, const std::string &bham, std::string &err);
int bham_to_csv(const std::string &bham, const std::string &csv, std::string &err);

// ---- HammingDist (HammingDist/HammingDist.cpp:371-705), region-less mode; region mode below ------------
// Distribution file of the reference's downstream tool: header `,"All","Proportion All","Cumulative All"`,
// then one row `\n<d>,<count>,<proportion %f>,<cumulative %f>` for d = 0 .. max-1 (the reference's loops
// stop BEFORE the largest value seen, :631-641; kept, so the proportions are over the rows shown).
// counts[d] = number of K-mers whose minimum Hamming distance is d.
int write_hamming_distribution(const std::string &path, const std::vector<uint64_t> &counts, std::string &err);
// counts from `"chrom",loci,hamming` CSV files (hammings -m1 / -m5 output): field 3 is the distance - the
// reference's INTENDED semantics; its region-less mode never assigns the variable it histograms
// (HammingDist.cpp:380, :437-446, :472-478), so there is no reference output to be bit-exact with.
// Leading header lines and the `G,b,B` descriptor row of -m1 files are skipped.
int hamming_counts_from_csv(const std::vector<std::string> &files, std::vector<uint64_t> &counts, uint64_t &rows,
                            std::string &err);

// ---- HammingDist region mode: gene / feature annotation (libkit4b/BEDfile.cpp) -------------------------
// feature bits, BEDfile.h:28-33; the region of a locus is the first set bit in this order, else intergenic
enum : int {
    kFeatCDS = 0x01,
    kFeat5UTR = 0x02,
    kFeat3UTR = 0x04,
    kFeatIntrons = 0x08,
    kFeatUpstream = 0x10,
    kFeatDnstream = 0x20,
    kFeatRegionBits = 0x3f,  // cRegionFeatBits (BEDfile.h:48)
};
struct Feature {
    int32_t start = 0, end = 0;  // chromosome offsets, end inclusive
    int32_t score = 0;
    char strand = '+';           // '+', '-' or '?'
    // gene + exons files only: coding range and exon start/end pairs, all relative to `start`
    int32_t thick_start = 0, thick_end = 0;
    std::vector<int32_t> exons;
};
struct FeatureChrom {
    std::string name;
    std::vector<Feature> feats;  // sorted by start, end
    int32_t max_len = 0;
};
struct FeatureSet {
    bool gene_exons = false;  // features carry exon detail (eBTGeneExons)
    // false: the file was neither BED nor biobed and short enough (< 100 lines) that the reference's GFF3
    // fallback reports "no genes" as success (BEDfile.cpp:421-432 returns its zero count as the result
    // code): the tool then runs with NO chromosome known, which ends every input file at its first row
    bool available = true;
    std::vector<FeatureChrom> chroms;
    std::unordered_map<std::string, int> chrom_index;  // lower-cased name -> index into chroms
    // index of the chromosome, -1 if absent; case-insensitive, chloroplast/ChrC and mitochondria/ChrM are
    // aliases of each other (BEDfile.cpp:3070-3095)
    int chrom_id(const std::string &name) const;
    // OR of the overlap bits of every feature touching [start-updn, end+updn] (BEDfile.cpp:4279-4311)
    int feature_bits(int chrom, int start, int end, int want_bits, int updn) const;
};
// raw BED text (BED3..BED12, tabs or commas), the binary biobed container written by `genbiobed`, or GFF3
// gene models (tried, as in the reference, when the text is no BED)
int read_features(const std::string &path, FeatureSet &fs, std::string &err);

// Region mode of HammingDist (HammingDist.cpp:371-496, :606-700): every row `"chrom",loci,hamming` is put
// into the region of loci+ofs_loci (regulatory length reg_len) and the distances are histogrammed per
// region.  Bit-exact with the reference, its reading rules included: no row is special-cased (the
// `G,b,B` descriptor row of -m1 output names no chromosome of the feature file, and the FIRST row whose
// chromosome is unknown ends the reading of that file, :449-453); distances above 200 do not fit its
// table (:278) and are refused here.
constexpr int kNumRegions = 7, kMaxRegionHamming = 200;
struct RegionHistogram {
    uint32_t counts[kNumRegions][kMaxRegionHamming + 1] = {};
    int max_hamming = -1;      // largest distance in any file that was read to its end
    long total_processed = 0;  // rows of those files
};
int region_counts_from_csv(const std::vector<std::string> &files, const FeatureSet &fs, int ofs_loci, int reg_len,
                           RegionHistogram &h, std::string &log, std::string &err);
int write_region_distribution(const std::string &path, const RegionHistogram &h, std::string &err);

// ---- CLI ----------------------------------------------------------------------------------------
struct Options {
    int mode = 0;            // -m (default 0 = restricted, hammings.cpp:312)
    int sensitivity = 0;     // -s
    int resformat = 0;       // -S
    bool crick = false;      // -c
    int intrainterboth = 0;  // -z
    int rhamm = 3;           // -r
    std::string prefix;      // -p
    int numnodes = 1, node = 0;  // -n -N
    int sweep_start = 0, sweep_end = 0;  // -b -B
    int K = 100;             // -K
    int sample = 1;          // -k
    std::string in_file, in_seq_file, out_file, log_file;  // -i -I -o -F
    int file_log_level = 3;  // -f
    int threads = 0;         // -T
    int gpus = 0;            // --gpus (extension; 0 = all visible)
    std::string dist_file;   // --dist (extension, -m1/-m2): HammingDist-format distribution of the minima
    bool help = false, version = false;
};
// returns 0, or -1 after printing the problem (caller prints usage and exits 1)
int parse_args(int argc, char **argv, Options &o, std::string &err);
// expands @file arguments (libkit4b/Utility.cpp:1200-1313)
int expand_param_files(int argc, char **argv, std::vector<std::string> &out, std::string &err);
// argtable3-style integer: decimal, 0x/0o/0b prefixes, optional KB/MB/GB suffix
bool parse_int_arg(const char *s, long &v);
void print_usage(const char *prog);

}  // namespace k4bhost
