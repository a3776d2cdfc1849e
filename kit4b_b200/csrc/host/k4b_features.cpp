// k4b_features.cpp - gene / feature annotation for the region mode of the HammingDist drop-in.
//
// The reference's HammingDist (HammingDist/HammingDist.cpp:371-705) characterises every K-mer
// start locus into one of seven genomic regions before it histograms the distance:
// CBEDfile::GetFeatureBits (libkit4b/BEDfile.cpp:4279-4311) ORs the overlap bits of every feature
// that touches the locus (+/- the regulatory length), and the FIRST set bit in the order CDS, 5'UTR,
// 3'UTR, intron, 5' upstream, 3' downstream names the region; no bit = intergenic (:463-468).
// This file restates what that path needs, nothing else of CBEDfile:
//   * the raw BED reader (BEDfile.cpp:1172-1337 ProcessBedFile, :1794-2072 AddFeature): tab or comma
//     separated, BED3..BED12, the file becomes a "gene + exons" file at the first line that carries
//     thickStart..blockStarts; and the preprocessed binary 'bios' type-7 container that `genbiobed`
//     writes (BEDfile.h:141-163 header, :105-137 records; BEDfile.cpp:2255-2470 LoadFeatures); and GFF3
//     gene models, which the reference tries when the BED reader gives up (BEDfile.cpp:757-1131);
//   * chromosome lookup, case-insensitive, with the chloroplast/ChrC and mitochondria/ChrM aliases
//     (BEDfile.cpp:3070-3095);
//   * the per-feature overlap rules (BEDfile.cpp:4011-4170 GetFeatureOverlaps) - including that ANY
//     exon overlap raises CDS, 5'UTR and 3'UTR together (the exon mask is the union of the three,
//     BEDfile.h:43), so exonic loci always land in "CDS".
// Features are kept per chromosome sorted by start; a query scans back by the longest feature of the
// chromosome, as the reference's index does (BEDfile.cpp:3247-3305).
#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>

#include "k4b_host.h"

namespace k4bhost {
namespace {

constexpr int kMaxChromName = 81, kMaxFeatName = 81;  // commdefs.h:139-140 (incl. the terminator)
constexpr int kMaxExons = 8000;                       // BEDfile.h:10

std::string lower(const std::string &s) {
    std::string r = s;
    for (char &c : r) c = (char)tolower((unsigned char)c);
    return r;
}

int chrom_slot(FeatureSet &fs, const std::string &name) {
    const std::string key = lower(name);
    auto it = fs.chrom_index.find(key);
    if (it != fs.chrom_index.end()) return it->second;
    const int id = (int)fs.chroms.size();
    fs.chrom_index.emplace(key, id);
    fs.chroms.emplace_back();
    fs.chroms.back().name = name;
    return id;
}

// thickStart thickEnd reserved blockCount blockSizes blockStarts -> gene structure relative to `start`
// (AddFeature, BEDfile.cpp:1886-1931).  Returns false on a malformed line (eBSFerrFeature).
bool parse_gene_detail(const char *supp, int start, int end, Feature &f) {
    int thick_start = 0, thick_end = 0, reserved = 0, blocks = 0, psn = 0;
    const int cnt = (supp && *supp) ? sscanf(supp, " %d %d %d %d %n", &thick_start, &thick_end, &reserved, &blocks, &psn) : 0;
    f.exons.clear();
    if (cnt <= 0) {  // no detail that starts with a number: the feature itself is the one exon
        f.thick_start = 0;
        f.thick_end = end - start;
        f.exons.push_back(0);
        f.exons.push_back(end - start - 1);
        return true;
    }
    if (cnt != 4 || blocks > kMaxExons) return false;
    std::vector<int> sizes((size_t)std::max(blocks, 0), 0), starts((size_t)std::max(blocks, 1), 0);
    int rel = 0;
    for (int i = 0; i < blocks; ++i) {
        rel = 0;
        sscanf(supp + psn, " %d , %n", &sizes[(size_t)i], &rel);
        psn += rel;
    }
    int i = 0;
    for (; i < blocks - 1; ++i) {
        rel = 0;
        sscanf(supp + psn, " %d , %n", &starts[(size_t)i], &rel);
        psn += rel;
    }
    sscanf(supp + psn, " %d", &starts[(size_t)i]);
    for (int b = 0; b < blocks; ++b) {
        f.exons.push_back(starts[(size_t)b]);
        f.exons.push_back(starts[(size_t)b] + sizes[(size_t)b] - 1);
    }
    f.thick_start = thick_start - start;
    f.thick_end = thick_end - start;  // exclusive in BED, kept as is (BEDfile.cpp:1927)
    return true;
}

void finish(FeatureSet &fs) {
    for (FeatureChrom &c : fs.chroms) {
        std::stable_sort(c.feats.begin(), c.feats.end(), [](const Feature &a, const Feature &b) {
            return a.start != b.start ? a.start < b.start : a.end < b.end;
        });
        c.max_len = 0;
        for (const Feature &f : c.feats) c.max_len = std::max(c.max_len, f.end - f.start + 1);
    }
}

int read_bed_text(const std::string &path, FeatureSet &fs, std::string &err) {
    FILE *in = fopen(path.c_str(), "r");
    if (!in) {
        err = "Unable to fopen BED format file " + path + " error: " + strerror(errno);
        return kErrOpnFile;
    }
    std::vector<char> buf(128 + kMaxExons * 8), attrs;  // BEDfile.h:11
    char chrom[2 * kMaxChromName + 1], name[2 * kMaxFeatName + 1];
    int line_no = 0, n_feats = 0, rc = kOk;
    bool csv = false;
    while (fgets(buf.data(), (int)buf.size() - 1, in)) {
        ++line_no;
        for (const char *p = buf.data(); *p; ++p)
            if ((unsigned char)*p > 127) {
                err = "Errors whilst parsing - " + path + " - non-ascii chars at line " + std::to_string(line_no);
                rc = kErrParams;  // fatal: the GFF3 fallback rejects such a file as well (BEDfile.cpp:1025-1030)
                break;
            }
        if (rc) break;
        if (!n_feats && line_no >= 20) {  // no feature within the first 20 lines: not a BED file
            err = path + " does not look like a BED file (no feature line in its first 20 lines)";
            rc = kErrFileType;
            break;
        }
        char *txt = buf.data();
        while (*txt && isspace((unsigned char)*txt)) ++txt;
        size_t len = strlen(txt);
        while (len && isspace((unsigned char)txt[len - 1])) txt[--len] = '\0';
        if (!*txt || *txt == '#') continue;
        int c_start = 0, c_end = 0, score = 0, supp_at = 0, cnt = 0;
        char strand = '+';
        if (!csv) {
            cnt = sscanf(txt, " %140s %d %d %140s %d %c %n", chrom, &c_start, &c_end, name, &score, &strand, &supp_at);
            if (!n_feats && cnt < 3) csv = true;  // perhaps commas separate the fields
        }
        if (csv) cnt = sscanf(txt, " %140s , %d , %d , %140s , %d , %c , %n", chrom, &c_start, &c_end, name, &score, &strand, &supp_at);
        if (cnt < 3) {
            if (!n_feats) {  // a header line such as `track ...` ahead of the features
                csv = false;
                continue;
            }
            err = "Errors whilst parsing - " + path + " line " + std::to_string(line_no);
            rc = kErrParse;
            break;
        }
        chrom[kMaxChromName - 1] = '\0';
        name[kMaxFeatName - 1] = '\0';
        if (cnt == 3) snprintf(name, sizeof(name), "feat%d", n_feats + 1);
        if (cnt <= 4) score = 0;
        if (cnt <= 5) strand = '+';
        const bool has_supp = cnt >= 6 && txt[supp_at] != '\0' && txt[supp_at] != '\n' && txt[supp_at] != '\r';
        const char *supp = nullptr;
        if (fs.gene_exons) {  // every line needs exon detail; without it the feature is its own single exon
            if (has_supp) supp = txt + supp_at;
        } else if (cnt == 6 && has_supp) {  // first line with gene detail: gene + exons file from here on
            supp = txt + supp_at;
            fs.gene_exons = true;
        } else if (cnt > 6) {
            continue;
        }
        if (strand == '.') strand = '+';
        else if (!(strand == '+' || strand == '-' || strand == '?')) strand = '?';
        if (!name[0] || !chrom[0] || c_start < 0 || c_end < c_start) {
            err = "BED feature with bad coordinates or names at line " + std::to_string(line_no) + " of " + path;
            rc = kErrParams;
            break;
        }
        Feature f;
        f.start = c_start;
        f.end = c_end - 1;  // inclusive
        f.score = score;
        f.strand = strand;
        if (fs.gene_exons) {
            attrs.clear();
            if (supp) attrs.assign(supp, supp + strlen(supp));
            attrs.push_back('\0');
            if (!parse_gene_detail(supp ? attrs.data() : nullptr, c_start, c_end, f)) {
                err = "BED gene detail (thickStart..blockStarts) malformed at line " + std::to_string(line_no) + " of " + path;
                rc = kErrFeature;
                break;
            }
        }
        fs.chroms[(size_t)chrom_slot(fs, chrom)].feats.push_back(std::move(f));
        ++n_feats;
    }
    fclose(in);
    if (rc) return rc;
    if (!n_feats) {
        err = "Unable to load any features from '" + path + "'";
        return kErrParse;
    }
    // lines read before the file turned into a gene file carry no exon detail; the reference would read
    // whatever bytes follow their names (undefined) - here they are their own single exon
    if (fs.gene_exons)
        for (FeatureChrom &c : fs.chroms)
            for (Feature &f : c.feats)
                if (f.exons.empty()) parse_gene_detail(nullptr, f.start, f.end + 1, f);
    return kOk;
}

// ---- GFF3 gene models (BEDfile.cpp:757-1131) ---------------------------------------------------
// What the reference makes of a GFF3 file when the BED reader has given up on it: gene (and "orphan"
// mRNA) lines open a gene, exon / CDS / UTR lines paint a 2-bit map of the gene, and when the next gene
// opens the map is turned into ONE BED12-style feature (exon blocks = maximal painted runs).  Kept with
// its reading rules, because they decide which loci count as exonic:
//   * the gene's name is its Name= attribute; a gene without one is dropped (AddFeature refuses the
//     empty name and the caller ignores that, :734);
//   * an mRNA line directly after a gene line is ignored, any LATER mRNA line (second isoform) opens a
//     new "gene" under the mRNA's name (:1042-1046: the flag is cleared by the first exon / CDS / UTR);
//   * coordinates stay 1-based: start = GFF start, end = GFF end (:734), blocks relative to the start;
//   * thickStart / thickEnd are written RELATIVE to the gene (:697-725) but read back as absolute;
//     a gene without painted loci gets one block of end - start bases and absolute thick values;
//   * of a gene's blocks only the first two survive: the writer repeats the second for all later ones.
enum GffType { kGffGene = 1, kGffMrna, kGffExon, kGffIntron, kGffCds, kGffUtr3, kGffUtr5 };
constexpr int kGffAttrLen = 200;  // BEDfile.h:86

struct GffLine {
    char chrom[82], name[kGffAttrLen + 2];
    int start, end, score;
    char strand;
};

// ParseGFFline (:757-957): > 0 the feature type, 0 a type of no interest, < 0 not a feature line
int parse_gff_line(const char *line, GffLine &g) {
    char source[82], feature[52], score_s[52], strand = 0, frame = 0;
    int attr_at = 0;
    g.chrom[0] = g.name[0] = '\0';
    const int n = sscanf(line, " %80s %80s %50s %d %d %50s %c %c %n", g.chrom, source, feature, &g.start, &g.end, score_s, &strand,
                         &frame, &attr_at);
    if (n < 8) return -1;
    if (score_s[0] == '.') g.score = 0;
    else g.score = std::min(999, std::max(0, (int)atof(score_s)));
    g.strand = strand == '-' ? '-' : '+';
    static const char *const kTypes[] = {"gene", "mRNA", "exon", "intron", "CDS", "three_prime_UTR", "five_prime_UTR"};
    int type = 0;
    for (; type < 7; ++type)
        if (!strcasecmp(feature, kTypes[type])) break;
    if (type == 7) return 0;
    // attributes: ID=, Parent= and Name= are recognised wherever they start while no value is being read (also
    // inside other keys and values); a value ends at ';', white space inside it is dropped
    static const char *const kAttrs[] = {"ID", "Parent", "Name"};
    char values[3][kGffAttrLen + 2];
    memset(values, 0, sizeof(values));
    char *val = nullptr;
    int val_len = 0, state = 0;
    const char *txt = line + attr_at;
    const int len = (int)strlen(txt);
    for (int i = 0; i < len; ++i, ++txt) {
        const char c = *txt;
        if (state == 0) {
            if (c == ';' || c == '=') continue;
            int a = 0, ident = 0;
            for (; a < 3; ++a) {
                ident = (int)strlen(kAttrs[a]);
                if (!strncmp(txt, kAttrs[a], (size_t)ident)) break;
            }
            if (a == 3 || txt[ident] != '=') continue;
            val = values[a];
            memset(val, 0, kGffAttrLen + 2);
            val_len = 0;
            i += ident;  // the loop step skips the '='
            txt += ident;
            state = 1;
            continue;
        }
        if (isspace((unsigned char)c)) continue;
        if (c == ';') {
            state = 0;
            continue;
        }
        if (val_len < kGffAttrLen) val[val_len++] = c;
    }
    strcpy(g.name, values[2]);
    return type + 1;
}

// AddFeatExonsBitmap (:565-631): 2 bits per gene locus; 0 none, 1 exon, 2 CDS, 3 UTR; CDS wins, UTR beats exon
bool gff_paint(int type, int f_start, int f_end, int g_start, int g_end, std::vector<uint8_t> &map) {
    if (f_start > f_end) std::swap(f_start, f_end);
    if (f_start < g_start || f_end > g_end) return false;
    const uint32_t want = type == kGffExon ? 1u : type == kGffCds ? 2u : 3u;
    for (int rel = f_start - g_start; rel <= f_end - g_start; ++rel) {
        const uint32_t sh = 2u * ((uint32_t)rel & 3u);
        const uint32_t cur = (map[(size_t)rel >> 2] >> sh) & 3u;
        uint32_t now = want;
        if (cur == 2) now = 2;
        else if (cur == 3 && want == 1) now = 3;
        map[(size_t)rel >> 2] = (uint8_t)((map[(size_t)rel >> 2] & ~(3u << sh)) | (now << sh));
    }
    return true;
}

// AddGFF3Gene2BED (:636-737) + AddFeature: the painted map as one gene feature
void gff_flush_gene(FeatureSet &fs, const GffLine &gene, const std::vector<uint8_t> &map) {
    std::vector<int> rel_start, size;
    int min_cds = -1, max_cds = -1;
    uint32_t prev = 0;
    const int span = 1 + gene.end - gene.start;
    for (int rel = 0; rel < span; ++rel) {
        uint32_t cur = (map[(size_t)rel >> 2] >> (2u * ((uint32_t)rel & 3u))) & 3u;
        if (cur == 3) cur = 1;  // UTRs are exonic
        if (cur == 0) {
            prev = 0;
            continue;
        }
        if (prev == 0) {
            rel_start.push_back(rel);
            size.push_back(1);
            if (cur == 2 && min_cds == -1) min_cds = max_cds = rel;
            prev = cur;
            continue;
        }
        if (prev != 2 && cur == 2) {
            if (min_cds == -1) min_cds = max_cds = rel;
        } else if (prev == 2 && cur == 2) {
            max_cds = rel;
        }
        // (prev keeps the type the block began with, as in the reference: a CDS that follows a UTR inside one
        // block never extends max_cds beyond its first base)
        size.back() += 1;
    }
    if (rel_start.empty()) {
        rel_start.push_back(0);
        size.push_back(gene.end - gene.start);
        min_cds = gene.start;
        max_cds = gene.end + 1;
    }
    if (min_cds == -1) min_cds = gene.start;
    if (max_cds == -1) max_cds = gene.end + 1;
    // AddFeature's own checks (:1822-1830); its refusal is ignored by the caller
    const int start = gene.start, end = gene.end + 1;
    if (!gene.name[0] || !gene.chrom[0] || start < 0 || end < start || strlen(gene.name) > (size_t)kMaxFeatName - 1 ||
        strlen(gene.chrom) > (size_t)kMaxChromName - 1 || rel_start.size() > (size_t)kMaxExons)
        return;
    std::string supp = std::to_string(min_cds) + " " + std::to_string(max_cds) + " 0 " + std::to_string(rel_start.size()) + " ";
    // the reference's writer never advances past the SECOND block (:727-735): blocks 3.. repeat block 2, so a
    // gene keeps its first two exons only and everything after the second one is outside every exon
    for (size_t i = 0; i < size.size(); ++i) supp += (i ? "," : "") + std::to_string(size[std::min<size_t>(i, 1)]);
    supp += ", ";
    for (size_t i = 0; i < rel_start.size(); ++i) supp += (i ? "," : "") + std::to_string(rel_start[std::min<size_t>(i, 1)]);
    Feature f;
    f.start = start;
    f.end = end - 1;
    f.score = gene.score;
    f.strand = gene.strand;
    if (!parse_gene_detail(supp.c_str(), start, end, f)) return;
    fs.chroms[(size_t)chrom_slot(fs, gene.chrom)].feats.push_back(std::move(f));
}

// ParseGFF3FileGFFFeats (:965-1131).  > 0: number of genes opened; 0: none (the caller treats that as
// "no features available"); < 0: error
int read_gff3_text(const std::string &path, FeatureSet &fs, std::string &err) {
    FILE *in = fopen(path.c_str(), "r");
    if (!in) {
        err = "Error accessing GFF file " + path + " - " + strerror(errno);
        return kErrOpnFile;
    }
    fs.gene_exons = true;
    std::vector<char> buf(128 + kMaxExons * 8);
    std::vector<uint8_t> map;
    GffLine gene, line;
    memset(&gene, 0, sizeof(gene));
    int line_no = 0, n_genes = 0, n_feats = 0, rc = kOk;
    bool fresh_gene = false, have_gene = false;
    while (fgets(buf.data(), (int)buf.size() - 1, in)) {
        ++line_no;
        for (const char *p = buf.data(); *p; ++p)
            if ((unsigned char)*p > 127) {
                err = "Errors whilst parsing - " + path + " - non-ascii chars at line " + std::to_string(line_no);
                rc = kErrParse;
                break;
            }
        if (rc) break;
        if (n_genes < 1 && n_feats == 0 && line_no >= 100) {  // no gene within the first 100 lines: not GFF3
            err = path + " is neither a BED nor a GFF3 file (no gene feature in its first 100 lines)";
            rc = kErrFileType;
            break;
        }
        char *txt = buf.data();
        while (*txt && isspace((unsigned char)*txt)) ++txt;
        size_t len = strlen(txt);
        while (len && isspace((unsigned char)txt[len - 1])) txt[--len] = '\0';
        if (!*txt || *txt == '#') continue;
        const int type = parse_gff_line(txt, line);
        if (type <= 0 || type == kGffIntron) continue;
        if (type == kGffMrna && fresh_gene) continue;  // the gene's own transcript
        if (type == kGffGene || type == kGffMrna) {
            if (n_feats) gff_flush_gene(fs, gene, map);
            gene = line;
            if (gene.end < gene.start || (long long)gene.end - gene.start >= 0x7fffffffLL) {  // the reference sizes its map from end - start as an unsigned number
                err = "Errors whilst parsing - " + path + " - at line " + std::to_string(line_no) + ", gene ends before it starts";
                rc = kErrParse;
                break;
            }
            map.assign((size_t)(1 + (1 + gene.end - gene.start) / 4), 0);
            n_feats = 0;
            ++n_genes;
            fresh_gene = have_gene = true;
        } else {
            // (before any gene line the reference compares with uninitialised bounds; refused here)
            if (!have_gene || line.start < gene.start || line.end > gene.end) {
                err = "Errors whilst parsing - " + path + " - at line " + std::to_string(line_no) +
                      ", feature start/end are outside range of gene start/end";
                rc = kErrParse;
                break;
            }
            if (!gff_paint(type, line.start, line.end, gene.start, gene.end, map)) {
                err = "Errors generating packed bitmap of gene exons whilst parsing - " + path + " - at line " + std::to_string(line_no);
                rc = kErrParse;
                break;
            }
            fresh_gene = false;
        }
        ++n_feats;
    }
    fclose(in);
    if (rc) return rc;
    if (n_feats) gff_flush_gene(fs, gene, map);
    return n_genes;
}

template <class T>
bool rd(const std::vector<uint8_t> &img, size_t ofs, T &v) {
    if (ofs + sizeof(T) > img.size()) return false;
    memcpy(&v, img.data() + ofs, sizeof(T));
    return true;
}

int read_biobed(const std::string &path, FeatureSet &fs, std::string &err) {
    std::ifstream in(path, std::ios::binary);
    std::vector<uint8_t> img((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    // tsBEDFileHdr, pack(8), 1176 bytes (BEDfile.h:141-163)
    int64_t names_ofs = 0, feats_ofs = 0;
    uint32_t type = 0, version = 0, feats_size = 0;
    int32_t n_chroms = 0, n_feats = 0, feat_type = 0, names_size = 0;
    if (img.size() < 1176 || !rd(img, 16, names_ofs) || !rd(img, 24, feats_ofs) || !rd(img, 48, type) || !rd(img, 52, version) ||
        !rd(img, 64, n_chroms) || !rd(img, 72, n_feats) || !rd(img, 76, feat_type) || !rd(img, 80, feats_size) ||
        !rd(img, 84, names_size)) {
        err = "Read of file header failed on " + path;
        return kErrFileAccess;
    }
    if (type != 7) {  // cBSFTypeFeat (commdefs.h:64)
        err = path + " opened as a bioseq file - expected type 7, file type is " + std::to_string(type);
        return kErrFileType;
    }
    if (version < 11 || version > 12) {  // BEDfile.h:4-5
        err = path + ": can only handle biobed versions 11 to 12, file version is " + std::to_string(version);
        return kErrFileVer;
    }
    constexpr size_t kNameRec = 24 + kMaxChromName, kFeatFixed = 43;  // pack(1) records, BEDfile.h:105-137
    if (n_chroms < 0 || n_feats < 0 || names_ofs < 0 || feats_ofs < 0 || (size_t)names_ofs + (size_t)n_chroms * kNameRec > img.size() ||
        (size_t)feats_ofs + feats_size > img.size()) {
        err = path + ": biobed tables lie outside the file";
        return kErrParse;
    }
    fs.gene_exons = feat_type == 1;  // eBTGeneExons
    std::vector<int> slot_of_id((size_t)n_chroms + 1, -1);
    for (int c = 0; c < n_chroms; ++c) {
        const size_t o = (size_t)names_ofs + (size_t)c * kNameRec;
        int32_t id = 0;
        rd(img, o, id);
        const char *nm = (const char *)img.data() + o + 24;
        const int slot = chrom_slot(fs, std::string(nm, strnlen(nm, kMaxChromName)));
        if (id >= 1 && id <= n_chroms) slot_of_id[(size_t)id] = slot;
    }
    size_t o = (size_t)feats_ofs;
    const size_t feats_end = o + feats_size;
    for (int i = 0; i < n_feats; ++i) {
        int32_t size = 0, chrom_id = 0, start = 0, end = 0, score = 0;
        uint8_t name_len = 0;
        if (o + kFeatFixed > feats_end || !rd(img, o + 4, size) || size < (int32_t)kFeatFixed || o + (size_t)size > feats_end) {
            err = path + ": biobed feature " + std::to_string(i + 1) + " is truncated";
            return kErrParse;
        }
        rd(img, o + 20, chrom_id);
        rd(img, o + 24, start);
        rd(img, o + 28, end);
        rd(img, o + 32, score);
        rd(img, o + 40, name_len);
        if (chrom_id < 1 || chrom_id > n_chroms || slot_of_id[(size_t)chrom_id] < 0) {
            err = path + ": biobed feature " + std::to_string(i + 1) + " names an unknown chromosome";
            return kErrParse;
        }
        Feature f;
        f.start = start;
        f.end = end;
        f.score = score;
        f.strand = (char)img[o + 41];
        if (fs.gene_exons) {  // tsGeneStructure after the name: Size, NumExons, thickStart, thickEnd, start/end pairs
            const size_t g = o + 42 + name_len + 1;
            int32_t n_exons = 0;
            if (g + 16 > o + (size_t)size || !rd(img, g + 4, n_exons) || n_exons < 0 || g + 16 + (size_t)n_exons * 8 > o + (size_t)size) {
                err = path + ": biobed gene detail of feature " + std::to_string(i + 1) + " is truncated";
                return kErrParse;
            }
            rd(img, g + 8, f.thick_start);
            rd(img, g + 12, f.thick_end);
            f.exons.resize((size_t)n_exons * 2);
            if (n_exons) memcpy(f.exons.data(), img.data() + g + 16, (size_t)n_exons * 8);
        }
        fs.chroms[(size_t)slot_of_id[(size_t)chrom_id]].feats.push_back(std::move(f));
        o += (size_t)size;
    }
    return kOk;
}

}  // namespace

int read_features(const std::string &path, FeatureSet &fs, std::string &err) {
    fs = FeatureSet();
    char magic[4] = {0, 0, 0, 0};
    {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) {
            err = "Unable to open " + path + " - " + strerror(errno);
            return kErrOpnFile;
        }
        const size_t n = fread(magic, 1, 4, f);
        fclose(f);
        if (n < 4) {
            err = "Read of file header failed on " + path;
            return kErrFileAccess;
        }
    }
    const bool bios = tolower(magic[0]) == 'b' && tolower(magic[1]) == 'i' && tolower(magic[2]) == 'o' && tolower(magic[3]) == 's';
    int rc = bios ? read_biobed(path, fs, err) : read_bed_text(path, fs, err);
    if (!bios && (rc == kErrFileType || rc == kErrParse)) {
        // Not BED: the reference tries GFF3 next (BEDfile.cpp:415-432).  A file without GFF3 gene lines fails
        // there only once it has 100 lines (:1031-1035); a shorter one "succeeds" with no features available
        // (the zero gene count is returned as the result code).
        fs = FeatureSet();
        std::string gff_err;
        const int genes = read_gff3_text(path, fs, gff_err);
        if (genes < 0) {
            err = gff_err;
            return genes;
        }
        if (genes == 0) {
            fs = FeatureSet();
            fs.available = false;
            return kOk;
        }
        size_t kept = 0;
        for (const FeatureChrom &c : fs.chroms) kept += c.feats.size();
        if (!kept) {  // genes without a Name= attribute are dropped one by one: SortFeatures then finds nothing (:1530)
            err = "Unable to load any features from '" + path + "'";
            return kErrNoFeatures;
        }
        rc = kOk;
    }
    if (rc) return rc;
    finish(fs);
    return kOk;
}

int FeatureSet::chrom_id(const std::string &name) const {
    if (name.empty() || !available) return -1;
    const std::string key = lower(name);
    auto find = [&](const std::string &k) -> int {
        auto it = chrom_index.find(k);
        return it == chrom_index.end() ? -1 : it->second;
    };
    int id = find(key);
    if (id < 0) {
        if (key == "chloroplast") id = find("chrc");
        else if (key == "mitochondria") id = find("chrm");
        if (id < 0) {
            if (key == "chrc") id = find("chloroplast");
            else if (key == "chrm") id = find("mitochondria");
        }
    }
    return id;
}

// overlap bits of one feature with [s, e] (GetFeatureOverlaps, BEDfile.cpp:4011-4170)
static int feature_overlaps(const Feature &f, bool gene_exons, int want, int s, int e, int dist) {
    if (dist <= 0) want &= ~(kFeatUpstream | kFeatDnstream);
    int got = 0;
    if (want & kFeatUpstream) {
        if (f.strand == '+' ? (s < f.start && e >= f.start - dist) : (s <= f.end + dist && e > f.end)) got |= kFeatUpstream;
    }
    if (want & kFeatDnstream) {
        if (f.strand == '+' ? (s <= f.end + dist && e > f.end) : (s < f.start && e > f.start - dist)) got |= kFeatDnstream;
    }
    if (s > f.end || e < f.start || !gene_exons) return got;
    const int rs = s <= f.start ? 0 : s - f.start;
    const int re = e > f.end ? f.end - f.start : e - f.start;
    const int n_exons = (int)(f.exons.size() / 2);
    constexpr int kExons = kFeatCDS | kFeat5UTR | kFeat3UTR;
    if (want & kExons) {
        for (int i = 0; i < n_exons; ++i) {
            const int xs = f.exons[(size_t)i * 2], xe = f.exons[(size_t)i * 2 + 1];
            if (re < xs) break;
            if (rs <= xe && re >= xs) {
                got |= want & kExons;  // any exon overlap raises every requested exon bit (BEDfile.h:43, :4098-4099)
                if (rs <= f.thick_end && re >= f.thick_start) got |= kFeatCDS;
                if ((got & kExons) == (want & kExons)) break;
            }
        }
    }
    if ((want & kFeatIntrons) && n_exons > 1) {
        for (int i = 0; i < n_exons - 1; ++i) {
            if (rs < f.exons[(size_t)(i + 1) * 2] && re > f.exons[(size_t)i * 2 + 1]) {
                got |= kFeatIntrons;
                break;
            }
            if (re < f.exons[(size_t)i * 2 + 1]) break;
        }
    }
    return got;
}

int FeatureSet::feature_bits(int chrom, int s, int e, int want, int updn) const {
    if (chrom < 0 || (size_t)chrom >= chroms.size()) return 0;
    if (updn <= 0) want &= ~(kFeatUpstream | kFeatDnstream);
    if (!(want & kFeatRegionBits)) return 0;
    int ls = s, le = e;  // the range features must touch to be looked at
    if (want & (kFeatUpstream | kFeatDnstream)) {
        ls = std::max(0, s - updn);
        le = e + updn;
    }
    const FeatureChrom &c = chroms[(size_t)chrom];
    // features starting at or before ls - max_len end before ls
    auto it = std::upper_bound(c.feats.begin(), c.feats.end(), ls - c.max_len,
                               [](int v, const Feature &f) { return v < f.start; });
    int bits = 0;
    for (; it != c.feats.end() && it->start <= le; ++it)
        if (it->end >= ls) bits |= feature_overlaps(*it, gene_exons, want, s, e, updn);
    return bits;
}

}  // namespace k4bhost
