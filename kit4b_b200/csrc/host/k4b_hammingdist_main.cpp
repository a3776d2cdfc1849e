// k4b_hammingdist - drop-in for the reference's `HammingDist` tool (HammingDist/HammingDist.cpp) in its
// region-less mode: reads `"chrom",loci,hamming` CSV files written by `hammings -m1` / `-m5` and writes
// the distribution of the distances (count, proportion, cumulative proportion per distance).
// Same flags as the reference (-m -s -r -R -i -I -o -f -F, @parameter files, -h / -v exit 1).
// Field 3 of every row is the distance - what the reference means to do; its own region-less mode
// histograms a variable it never assigns (HammingDist.cpp:380, :437-446, :472-478), so its output is
// undefined and cannot be compared bit for bit.  The BED-region mode (-I, biobed feature container)
// is outside the hot-path scope and is refused with a message.
#include <glob.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "k4b_host.h"

using namespace k4bhost;

static void usage(const char *prog) {
    printf("\n%s - Hamming Distance Distributions (B200-native drop-in, region-less mode)\nOptions ---\n", prog);
    printf("  [-hH] [-v] [-f <int>] [-F <file>] [-m <int>] [-s <int>] [-r <int>] [-R <int>] [-i <file>]... [-I <file>] -o <file>\n");
    printf("  -i, --incsv=<file>        input element CSV files (wildcards allowed), rows \"chrom\",loci,hamming\n");
    printf("  -o, --output=<file>       distributions output file\n");
    printf("  -I, --infeats=<file>      biobed feature file (region mode): not supported by this build\n");
    printf("  -m, -s, -r, -R            accepted for compatibility; they only matter in region mode\n");
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
        bool known = false;
        for (const char *const *p = (const char *const[]){"-m", "mode", "-s", "strandproc", "-r", "updnstream", "-R", "relofs",
                                                          "-f", "FileLogLevel", "-F", "log", nullptr}; *p; p += 2)
            if ((m = value(i, a, p[0], p[1], v))) { known = true; break; }
        if (!known) {
            printf("\nError: unrecognised option '%s'\n", a.c_str());
            usage(prog);
            return 1;
        }
    }
    if (out.empty()) {
        printf("\nError: no output file specified with '-o<file>'\n");
        usage(prog);
        return 1;
    }
    if (!feats.empty()) {
        printf("\nError: region mode ('-I %s') needs the biobed feature container, which this drop-in does not read\n", feats.c_str());
        return 1;
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
    std::vector<uint64_t> counts;
    uint64_t rows = 0;
    int rc = hamming_counts_from_csv(files, counts, rows, err);
    if (!rc) {
        printf("Processed %llu Hamming rows from %zu file(s)\n", (unsigned long long)rows, files.size());
        rc = write_hamming_distribution(out, counts, err);
    }
    if (rc) {
        printf("%s\n", err.c_str());
        return 1;
    }
    return 0;
}
