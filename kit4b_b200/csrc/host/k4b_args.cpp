// k4b_args.cpp - command line of the `hammings` drop-in: the reference's flag set
// (ngskit4b/hammings.cpp:211-246), `@paramfile` expansion (libkit4b/Utility.cpp:1200-1313) and
// argtable3-style integers (libkit4b/argtable3.cpp:3019-3063).  Parsing rides on glibc
// getopt_long, the same engine argtable3 sits on, so `-K25`, `-K 25`, `--seqlen=25`,
// `--seqlen 25`, grouped literals (`-cv`) and unique long-option prefixes all behave alike.
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "k4b_host.h"

namespace k4bhost {

namespace {
std::string trim(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && isspace((unsigned char)s[b])) ++b;
    while (e > b && isspace((unsigned char)s[e - 1])) --e;
    return s.substr(b, e - b);
}
}  // namespace

int expand_param_files(int argc, char **argv, std::vector<std::string> &out, std::string &err) {
    out.clear();
    for (int i = 0; i < argc; ++i) {
        if (argv[i][0] != '@') {
            out.push_back(argv[i]);
            continue;
        }
        const std::string fname = trim(argv[i] + 1);
        if (fname.empty()) {
            err = "No options file specified following '@' switch";
            return -1;
        }
        FILE *f = fopen(fname.c_str(), "r");
        if (!f) {
            err = "Unable to open options file '" + fname + "'\nError: " + strerror(errno);
            return -1;
        }
        char buf[8192];
        while (fgets(buf, sizeof(buf), f)) {
            const std::string line = trim(buf);
            if (line.empty() || line[0] == '#' || line[0] == ';' || (line[0] == '/' && line.size() > 1 && line[1] == '/'))
                continue;  // comment lines
            // whitespace separates options except inside quotes; the quote characters stay part
            // of the option text, as in the reference
            std::string cur;
            bool in_quotes = false, in_param = false;
            for (char ch : line) {
                if (ch == 0x16) ch = '-';
                if (ch == '"' || ch == '\'') {
                    in_quotes = !in_quotes;
                    in_param = true;
                    cur.push_back(ch);
                } else if ((ch == ' ' || ch == '\t') && !in_quotes) {
                    if (in_param) {
                        out.push_back(cur);
                        cur.clear();
                        in_param = false;
                    }
                } else {
                    cur.push_back(ch);
                    in_param = true;
                }
            }
            if (in_param || !cur.empty()) out.push_back(cur);
        }
        fclose(f);
    }
    return (int)out.size();
}

bool parse_int_arg(const char *s, long &v) {
    // [+-] (0x hex | 0o octal | 0b binary | decimal) [KB|MB|GB], surrounding blanks allowed
    const char *p = s;
    while (isspace((unsigned char)*p)) ++p;
    int sign = 1;
    if (*p == '+' || *p == '-') {
        if (*p == '-') sign = -1;
        ++p;
    }
    int base = 10;
    if (p[0] == '0' && (p[1] == 'x' || p[1] == 'X')) {
        base = 16;
        p += 2;
    } else if (p[0] == '0' && (p[1] == 'o' || p[1] == 'O')) {
        base = 8;
        p += 2;
    } else if (p[0] == '0' && (p[1] == 'b' || p[1] == 'B')) {
        base = 2;
        p += 2;
    }
    char *end = nullptr;
    errno = 0;
    const long mag = strtol(p, &end, base);
    if (end == p || errno == ERANGE) return false;
    long mult = 1;
    if ((end[0] == 'K' || end[0] == 'k') && (end[1] == 'B' || end[1] == 'b')) {
        mult = 1024L;
        end += 2;
    } else if ((end[0] == 'M' || end[0] == 'm') && (end[1] == 'B' || end[1] == 'b')) {
        mult = 1024L * 1024L;
        end += 2;
    } else if ((end[0] == 'G' || end[0] == 'g') && (end[1] == 'B' || end[1] == 'b')) {
        mult = 1024L * 1024L * 1024L;
        end += 2;
    }
    while (isspace((unsigned char)*end)) ++end;
    if (*end) return false;
    v = sign * mag * mult;
    return v >= -2147483647L - 1 && v <= 2147483647L;
}

void print_usage(const char *prog) {
    printf(
        "%s hammings [-hv] [-f <int>] [-F <file>] [-m <int>] [-p <str>] [-s <int>] [-r <int>] [-S <int>] [-c] "
        "[-z <int>] [-n <int>] [-N <int>] [-b <int>] [-B <int>] [-K <int>] [-k <int>] -i <file> [-I <file>] "
        "[-o <file>] [-T <int>] [--gpus=<int>] [--dist=<file>]\n",
        prog);
    printf(
        "  -h, --help                print this help and exit\n"
        "  -v, --version, --ver      print version information and exit\n"
        "  -f, --FileLogLevel=<int>  Level of diagnostics written to screen and logfile 0=fatal,1=errors,2=info,3=diagnostics,4=debug\n"
        "  -F, --log=<file>          diagnostics log file\n"
        "  -m, --mode=<int>          processing mode: 0 - restricted Hammings, 1 - exhaustive single node Hammings, 2 - exhaustive\n"
        "                            multiple node Hammings, 3 - merge multiple Hamming files, 4 - transform Hamming CSV into quick\n"
        "                            load binary format, 5 - transform quick load Hamming binary format into CSV (default = 0)\n"
        "  -s, --sensitivity=<int>   restricted Hamming sensitivity: 0 - normal, 1 - high, 2 - ultra, 3 - low (default = 0); the GPU\n"
        "                            engine is exhaustive, so every level gives the exact answer\n"
        "  -S, --resformat=<int>     restricted Hamming file output format: 0 - csv, 1 - UCSC BED, 2 - UCSC Wiggle (default = 0)\n"
        "  -c, --strandcrick         process Crick in addition to Watson strand\n"
        "  -z, --intrainterboth=<int> 0: both intra and inter sequence hammings, 1: intra only, 2: inter only (default 0)\n"
        "  -r, --rhamm=<int>         restricted hamming upper limit (1..10, default 3) only applies in mode 0\n"
        "  -p, --prefix=<str>        filtering prefix used in restricted mode (alphanumeric only, max 10 chars)\n"
        "  -n, --numnodes=<int>      total number of nodes (2..10000) if processing over multiple nodes\n"
        "  -N, --node=<int>          node instance (1..N) if processing over multiple nodes\n"
        "  -b, --sweepstart=<int>    process starting from this sweep instance inclusive (default = 1 for 1st)\n"
        "  -B, --sweepend=<int>      complete processing at this sweep instance inclusive (default = 0 for all remaining)\n"
        "  -K, --seqlen=<int>        Hamming edit distances for these length k-mer subsequences (range 10..5000, default is 100)\n"
        "  -i, --in=<file>           in mode 0 input sfx file, in mode 1 and 2 bioseq genome assembly file, in mode 3 merge from this file\n"
        "  -I, --seq=<file>          if restricted hamming processing then optional file containing source kmer sequences\n"
        "  -k, --sample=<int>        sample every Nth sweep instance / K-mer (default is 1)\n"
        "  -o, --out=<file>          output (merged) Hamming distances to this file\n"
        "  -T, --threads=<int>       number of host threads 0..128 (accepted for compatibility; the engine runs on the GPUs)\n"
        "      --gpus=<int>          number of B200 GPUs to use (default 0 = all visible)\n"
        "      --dist=<file>         (-m1/-m2) also write the distribution of the minima in the HammingDist tool's format\n");
}

int parse_args(int argc, char **argv, Options &o, std::string &err) {
    static const struct option longopts[] = {
        {"help", no_argument, nullptr, 'h'},           {"version", no_argument, nullptr, 'v'},
        {"ver", no_argument, nullptr, 'v'},            {"FileLogLevel", required_argument, nullptr, 'f'},
        {"log", required_argument, nullptr, 'F'},      {"mode", required_argument, nullptr, 'm'},
        {"sensitivity", required_argument, nullptr, 's'}, {"resformat", required_argument, nullptr, 'S'},
        {"strandcrick", no_argument, nullptr, 'c'},    {"intrainterboth", required_argument, nullptr, 'z'},
        {"rhamm", required_argument, nullptr, 'r'},    {"prefix", required_argument, nullptr, 'p'},
        {"numnodes", required_argument, nullptr, 'n'}, {"node", required_argument, nullptr, 'N'},
        {"sweepstart", required_argument, nullptr, 'b'}, {"sweepend", required_argument, nullptr, 'B'},
        {"seqlen", required_argument, nullptr, 'K'},   {"in", required_argument, nullptr, 'i'},
        {"seq", required_argument, nullptr, 'I'},      {"sample", required_argument, nullptr, 'k'},
        {"out", required_argument, nullptr, 'o'},      {"threads", required_argument, nullptr, 'T'},
        {"gpus", required_argument, nullptr, 1000},    {"dist", required_argument, nullptr, 1001},    {nullptr, 0, nullptr, 0}};
    bool have_in = false;
    optind = 0;  // full re-initialisation of glibc getopt
    opterr = 0;
    int c;
    auto want_int = [&](const char *name, int &dst) -> bool {
        long v;
        if (!parse_int_arg(optarg, v)) {
            err = std::string("invalid argument \"") + optarg + "\" to option " + name;
            return false;
        }
        dst = (int)v;
        return true;
    };
    while ((c = getopt_long(argc, argv, ":hvf:F:m:s:S:cz:r:p:n:N:b:B:K:i:I:k:o:T:", longopts, nullptr)) != -1) {
        switch (c) {
            case 'h': o.help = true; break;
            case 'v': o.version = true; break;
            case 'c': o.crick = true; break;
            case 'f': if (!want_int("-f|--FileLogLevel=<int>", o.file_log_level)) return -1; break;
            case 'm': if (!want_int("-m|--mode=<int>", o.mode)) return -1; break;
            case 's': if (!want_int("-s|--sensitivity=<int>", o.sensitivity)) return -1; break;
            case 'S': if (!want_int("-S|--resformat=<int>", o.resformat)) return -1; break;
            case 'z': if (!want_int("-z|--intrainterboth=<int>", o.intrainterboth)) return -1; break;
            case 'r': if (!want_int("-r|--rhamm=<int>", o.rhamm)) return -1; break;
            case 'n': if (!want_int("-n|--numnodes=<int>", o.numnodes)) return -1; break;
            case 'N': if (!want_int("-N|--node=<int>", o.node)) return -1; break;
            case 'b': if (!want_int("-b|--sweepstart=<int>", o.sweep_start)) return -1; break;
            case 'B': if (!want_int("-B|--sweepend=<int>", o.sweep_end)) return -1; break;
            case 'K': if (!want_int("-K|--seqlen=<int>", o.K)) return -1; break;
            case 'k': if (!want_int("-k|--sample=<int>", o.sample)) return -1; break;
            case 'T': if (!want_int("-T|--threads=<int>", o.threads)) return -1; break;
            case 1000: if (!want_int("--gpus=<int>", o.gpus)) return -1; break;
            case 1001: o.dist_file = optarg; break;
            case 'F': o.log_file = optarg; break;
            case 'p': o.prefix = optarg; break;
            case 'i': o.in_file = optarg; have_in = true; break;
            case 'I': o.in_seq_file = optarg; break;
            case 'o': o.out_file = optarg; break;
            case ':':
                err = std::string("option ") + argv[optind - 1] + " requires an argument";
                return -1;
            default:
                err = std::string("invalid option \"") + argv[optind - 1] + "\"";
                return -1;
        }
    }
    if (optind < argc) {
        err = std::string("unexpected argument \"") + argv[optind] + "\"";
        return -1;
    }
    if (!have_in && !o.help && !o.version) {
        err = "missing option -i|--in=<file>";
        return -1;
    }
    return 0;
}

}  // namespace k4bhost
