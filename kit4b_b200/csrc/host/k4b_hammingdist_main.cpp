// k4b_hammingdist - drop-in for the reference's `HammingDist` tool (HammingDist/HammingDist.cpp): reads
// `"chrom",loci,hamming` CSV files written by `hammings -m1` / `-m5` and writes the distribution of the
// distances (count, proportion, cumulative proportion per distance), over everything or - with a gene
// feature file, -I - per genomic region (CDS, UTRs, introns, up/downstream, intergenic).
// Same flags as the reference (-m -s -r -R -i -I -o -f -F, @parameter files, -h / -v exit 1).
// Region mode is bit-exact with the reference (goldens in tests/golden/hammingdist_*).  Region-less mode
// takes field 3 of every row as the distance - what the reference means to do; its own region-less mode
// histograms a variable it never assigns (HammingDist.cpp:380, :437-446, :472-478), so its output is
// undefined and cannot be compared bit for bit.
#include <glob.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "k4b_host.h"

using namespace k4bhost;

static void usage(const char *prog) {
    printf("\n%s - Hamming Distance Distributions (B200-native drop-in)\nOptions ---\n", prog);
    printf("  [-hH] [-v] [-f <int>] [-F <file>] [-m <int>] [-s <int>] [-r <int>] [-R <int>] [-i <file>]... [-I <file>] -o <file>\n");
    printf("  -i, --incsv=<file>        input element CSV files (wildcards allowed), rows \"chrom\",loci,hamming\n");
    printf("  -o, --output=<file>       distributions output file\n");
    printf("  -I, --infeats=<file>      input gene or feature biobed / BED file: distributions per genomic region\n");
    printf("  -r, --updnstream=<int>    length of 5'up or 3'down stream regulatory region (default = 0, range 0..1000000)\n");
    printf("  -R, --relofs=<int>        relative loci offset (default = 0, range +/-200)\n");
    printf("  -m, --mode=<int>          input loci file: 0 - CSV (default = 0)\n");
    printf("  -s, --strandproc=<int>    0 - independent, 1 - Watson, 2 - Crick (accepted; the reference never applies it)\n");
    printf("\nNote: Parameters can be entered into a parameter file, one parameter per line.\n");
    printf("      To invoke this parameter file then precede its name with '@'\n");
}

int main(int argc, char **argv) {
    const char *prog = "k4b_hammingdist";
    std::vector<std::string> args;
    std::string err;
    if (expand_param_files(argc, argv, args, err) < 0) {
        printf("\n%s\n", err.c_str());
        return 1;
    }
    std::vector<std::string> specs;
    std::string out, feats;
    long mode = 0, strand = 0, reg_len = 0, rel_ofs = 0, log_level = 3;
    auto value = [&](size_t &i, const std::string &a, const char *shortopt, const char *longopt, std::string &v) -> int {
        // -x<val>, -x <val>, --long=<val>, --long <val>; 1 = matched, 0 = not this option, -1 = value missing
        const std::string s = shortopt, l = std::string("--") + longopt;
        if (a.rfind(s, 0) == 0 && a.size() > s.size() && a[1] != '-') { v = a.substr(s.size()); return 1; }
        if (a.rfind(l + "=", 0) == 0) { v = a.substr(l.size() + 1); return 1; }
        if (a == s || a == l) {
            if (i + 1 >= args.size()) return -1;
            v = args[++i];
            return 1;
        }
        return 0;
    };
    for (size_t i = 1; i < args.size(); ++i) {
        const std::string &a = args[i];
        std::string v;
        int m;
        if (a == "-h" || a == "-H" || a == "--help") { usage(prog); return 1; }
        if (a == "-v" || a == "--version" || a == "--ver") { printf("\n%s Version (B200-native drop-in)\n", prog); return 1; }
        if ((m = value(i, a, "-i", "incsv", v))) { if (m < 0) break; specs.push_back(v); continue; }
        if ((m = value(i, a, "-o", "output", v))) { if (m < 0) break; out = v; continue; }
        if ((m = value(i, a, "-I", "infeats", v))) { if (m < 0) break; feats = v; continue; }
        bool known = false, bad = false;
        struct IntOpt { const char *s, *l; long *v; } int_opts[] = {{"-m", "mode", &mode}, {"-s", "strandproc", &strand},
                                                                    {"-r", "updnstream", &reg_len}, {"-R", "relofs", &rel_ofs},
                                                                    {"-f", "FileLogLevel", &log_level}};
        for (const IntOpt &io : int_opts)
            if ((m = value(i, a, io.s, io.l, v))) {
                known = true;
                if (m < 0 || !parse_int_arg(v.c_str(), *io.v)) {
                    printf("\nError: option '%s' needs an integer value\n", io.s);
                    bad = true;
                }
                break;
            }
        if (!known && (m = value(i, a, "-F", "log", v))) known = true;
        if (bad || !known) {
            if (!known) printf("\nError: unrecognised option '%s'\n", a.c_str());
            usage(prog);
            return 1;
        }
    }
    if (out.empty()) {
        printf("\nError: no output file specified with '-o<file>'\n");
        usage(prog);
        return 1;
    }
    // ranges as HammingDist.cpp:166-202
    if (mode != 0) {
        printf("Error: Processing mode '-m%ld' specified outside of range 0..0\n", mode);
        return 1;
    }
    if (strand < 0 || strand > 2) {
        printf("Error: Strand processing mode '-s%ld' must be in range 0..2\n", strand);
        return 1;
    }
    if (!feats.empty()) {
        if (reg_len < 0 || reg_len > 1000000) {
            printf("Regulatory region length '-r%ld' must be in range 0..1000000\n", reg_len);
            return 1;
        }
        if (rel_ofs < -200 || rel_ofs > 200) {
            printf("Relative offset '-R%ld' must be in range -200..200\n", rel_ofs);
            return 1;
        }
    }
    std::vector<std::string> files;
    for (const std::string &s : specs) {
        glob_t g;
        memset(&g, 0, sizeof(g));
        if (glob(s.c_str(), 0, nullptr, &g) == 0)
            for (size_t k = 0; k < g.gl_pathc; ++k) files.push_back(g.gl_pathv[k]);
        else
            printf("Unable to locate any input loci Hamming file matching '%s'\n", s.c_str());
        globfree(&g);
    }
    int rc;
    if (!feats.empty()) {
        FeatureSet fs;
        printf("Loading: %s\n", feats.c_str());
        rc = read_features(feats, fs, err);
        if (!rc) {
            RegionHistogram hist;
            std::string log;
            rc = region_counts_from_csv(files, fs, (int)rel_ofs, (int)reg_len, hist, log, err);
            fputs(log.c_str(), stdout);
            if (!rc) {
                printf("Processed %ld Hamming rows from %zu file(s)\n", hist.total_processed, files.size());
                rc = write_region_distribution(out, hist, err);
            }
        }
    } else {
        std::vector<uint64_t> counts;
        uint64_t rows = 0;
        rc = hamming_counts_from_csv(files, counts, rows, err);
        if (!rc) {
            printf("Processed %llu Hamming rows from %zu file(s)\n", (unsigned long long)rows, files.size());
            rc = write_hamming_distribution(out, counts, err);
        }
    }
    if (rc) {
        printf("%s\n", err.c_str());
        return 1;
    }
    return 0;
}
