// k4b_hammings_main.cpp - `k4b_hammings [hammings] <flags>`: drop-in for `ngskit4b hammings`.
// Same flags, defaults, validation order, input files, output bytes and exit codes as
// ngskit4b/hammings.cpp:169-680 (CLI) and :2598-2966 / :2121-2595 (Process); the engines
// behind it are the CUDA kernels reached through include/k4b_hamm.h.
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <future>

#include "../../../include/k4b_hamm.h"
#include "k4b_host.h"

using namespace k4bhost;

static const char *kProg = "k4b_hammings";
static const char *kVersion = "0.1.0 (B200 engine; kit4b 2.0.2 hammings interface)";
static FILE *g_logf = nullptr;
static int g_screen_level = 3, g_file_level = 0;

static void logmsg(int level, const char *fmt, ...) {
    // level: 0 fatal, 1 errors/warnings, 2/3 info (Diagnostics.h:5-54); same text to screen and log
    char msg[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof(msg), fmt, ap);
    va_end(ap);
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    struct tm tmv;
    localtime_r(&tv.tv_sec, &tmv);
    char ts[64];
    strftime(ts, sizeof(ts), "%b %d %H:%M:%S", &tmv);
    if (level <= g_screen_level) {
        printf("[%s.%06ld %d](%s) %s\n", ts, (long)tv.tv_usec, 1900 + tmv.tm_year, kProg, msg);
        fflush(stdout);
    }
    if (g_logf && level <= g_file_level) {
        fprintf(g_logf, "[%s.%06ld %d](%s) %s\n", ts, (long)tv.tv_usec, 1900 + tmv.tm_year, kProg, msg);
        fflush(g_logf);
    }
}

static int touch_output(const std::string &path) {
    // create + truncate + close up front so an unwritable path fails before the long run
    // (hammings.cpp:2714-2737)
    const int fd = open(path.c_str(), O_RDWR | O_CREAT, S_IRUSR | S_IWUSR);
    if (fd < 0 || ftruncate(fd, 0) != 0) {
        if (fd >= 0) close(fd);
        logmsg(0, "Process: unable to create/truncate output file '%s'", path.c_str());
        return kErrCreateFile;
    }
    close(fd);
    return kOk;
}

// ---- exhaustive (-m1) ------------------------------------------------------------------------
// The CUDA context of a B200 takes seconds to create: k4b_gpu_init runs on a helper thread while
// the input files are read, and is joined right before the first engine call.
static std::future<int> g_gpu_init;
static std::string g_gpu_init_err;  // k4b_last_error() is per thread: copied by the helper thread
static int gpu_ready() {
    if (!g_gpu_init.valid()) return 0;
    const int rc = g_gpu_init.get();
    if (rc) logmsg(0, "Unable to initialise the GPU engine (%d): %s", rc, g_gpu_init_err.c_str());
    return rc;
}

static int run_exhaustive(const Options &o) {
    std::vector<SeqEntry> entries;
    std::string title, err;
    int rc = read_sequences(o.in_file, entries, title, err);  // bioseq, or FASTA / FASTA.gz directly
    if (rc) {
        logmsg(0, "%s", err.c_str());
        logmsg(0, "Unable to open assembly sequence file '%s'", o.in_file.c_str());
        return rc;
    }
    const uint32_t K = (uint32_t)o.K;
    Genome g;
    build_genome(entries, K, g);
    entries.clear();
    entries.shrink_to_fit();
    logmsg(2, "Genome containing %llu total nucleotides loaded with %llu subsequences of length %u...",
           (unsigned long long)g.total_bases, (unsigned long long)g.num_subseqs, K);

    // -m1: -b/-B silently clamped to the genome length; -m2: work-balanced node slice
    // (hammings.cpp:2660-2706)
    uint32_t ss, se;
    if (o.mode == 2)
        node_sweep_range(g.genome_len, (uint32_t)g.chroms.size(), !o.crick, o.numnodes, o.node, ss, se);
    else
        single_sweep_range(g.genome_len, (uint32_t)o.sweep_start, (uint32_t)o.sweep_end, ss, se);
    logmsg(2, "Node sweep start is %u, and sweep end is %u", ss, se);

    if (!o.out_file.empty() && (rc = touch_output(o.out_file))) return rc;

    const uint32_t flat = g.genome_len - 2;
    std::vector<uint16_t> hd(flat, (uint16_t)(K + 1));
    if ((rc = gpu_ready())) return rc;
    logmsg(2, "Starting Hamming edit distance processing on %d GPU(s)", k4b_gpu_count());
    const auto t0 = std::chrono::steady_clock::now();
    if ((rc = gpu_ready())) return rc;
    rc = k4b_hamm_exhaustive(g.concat.data(), flat, K, o.crick ? 1 : 0, ss, se, hd.data());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc) {
        logmsg(0, "Hamming engine failed (%d): %s", rc, k4b_last_error());
        return rc;
    }
    const double cmps = (double)g.num_subseqs * (double)g.num_subseqs * (o.crick ? 2.0 : 1.0);
    logmsg(2, "Engine: %.3f s, %.1f G K-mer comparisons/s", secs, cmps / secs / 1e9);

    if (!o.out_file.empty()) {
        logmsg(2, "Writing Hamming edit distances to file: '%s'", o.out_file.c_str());
        rc = write_exhaustive_csv(o.out_file, g, K, hd.data(), ss, se, err);
        if (rc) {
            logmsg(0, "%s", err.c_str());
            return rc;
        }
    }
    // distribution to the log (hammings.cpp:2939-2962): GPU histogram of the minima; positions without a
    // K-mer hold K+1 and fall into the last bin
    std::vector<uint64_t> hist(K + 2, 0);
    if ((rc = k4b_hamm_histogram(hd.data(), flat, K, hist.data()))) {
        logmsg(0, "Histogram of the minima failed (%d): %s", rc, k4b_last_error());
        return rc;
    }
    if (!o.dist_file.empty()) {  // extension: the HammingDist tool's file without the CSV round trip
        std::vector<uint64_t> counts(hist.begin(), hist.begin() + K + 1);
        logmsg(2, "Writing the distribution of the minima to file: '%s'", o.dist_file.c_str());
        if ((rc = write_hamming_distribution(o.dist_file, counts, err))) {
            logmsg(0, "%s", err.c_str());
            return rc;
        }
    }
    logmsg(2, "Distribution:\nEditDist,Freq,Proportion");
    for (uint32_t d = 0; d < std::min(66u, K); ++d)
        printf("%u,%llu,%1.3f\n", d, (unsigned long long)hist[d],
               g.num_subseqs ? hist[d] * 100.0 / (double)g.num_subseqs : 0.0);
    return kOk;
}

// ---- restricted / targeted (-m0) ------------------------------------------------------------------
static int run_restricted(const Options &o) {
    std::string err, title;
    SfxData sfx;
    int rc;
    if (sniff_format(o.in_file) == 's') {
        logmsg(2, "Loading suffix array file '%s'", o.in_file.c_str());
        rc = read_sfx(o.in_file, sfx, err);
    } else {
        // extension: the assembly as bioseq / FASTA / FASTA.gz - only the sequence area of a suffix
        // array file is ever used here, and it is fully determined by the entries
        logmsg(2, "Loading assembly sequences '%s' (no suffix array file needed)", o.in_file.c_str());
        std::vector<SeqEntry> asm_entries;
        rc = read_sequences(o.in_file, asm_entries, title, err);
        if (!rc) sfx_from_entries(asm_entries, title, sfx);
    }
    if (rc) {
        logmsg(0, "%s", err.c_str());
        logmsg(0, "Unable to open input bioseq suffix array file '%s'", o.in_file.c_str());
        return rc;
    }
    logmsg(2, "Genome Assembly Name: '%s' Descr: '%s' Title: '%s' Version: %d", sfx.dataset.c_str(),
           sfx.descr.c_str(), sfx.title.c_str(), sfx.version);
    const uint32_t K = (uint32_t)o.K;
    const bool sep_probes = !o.in_seq_file.empty();
    Genome g;  // probe set in the LoadGenome layout; for no -I it mirrors the suffix entries
    std::vector<uint8_t> flat;
    double secs = 0;
    if (sep_probes) {
        std::vector<SeqEntry> entries;
        rc = read_sequences(o.in_seq_file, entries, title, err);
        if (rc) {
            logmsg(0, "%s", err.c_str());
            logmsg(0, "Unable to open assembly sequence file '%s'", o.in_seq_file.c_str());
            return rc;
        }
        build_genome(entries, K, g);
        entries.clear();
        logmsg(2, "Genome containing %llu total nucleotides loaded with %llu subsequences of K-mer length %u...",
               (unsigned long long)g.total_bases, (unsigned long long)g.num_subseqs, K);
        const uint32_t plen = (uint32_t)g.concat.size();
        flat.assign(plen, 0xff);
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = gpu_ready())) return rc;
        k4b_set_reference_sensitivity(o.sensitivity);
        rc = k4b_hamm_targeted(sfx.seq.data(), sfx.seq.size(), g.concat.data(), plen, K, o.rhamm, o.crick ? 1 : 0, 0,
                               plen, flat.data());
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } else {
        // K-mers of the indexed assembly against the assembly itself (hammings.cpp:2311-2330)
        for (const SfxEntry &e : sfx.entries) {
            Chrom c;
            c.name = e.name.substr(0, 80);
            c.len = e.len;
            c.start = (uint32_t)e.start;
            c.num_subseqs = c.len >= K ? c.len - K + 1 : 0;
            g.total_bases += c.len;
            g.num_subseqs += c.num_subseqs;
            g.chroms.push_back(std::move(c));
        }
        logmsg(2, "Targeted suffix genome has %zu sequences containing %llu total nucleotides, %llu sequences of at "
                  "least K-mer length %u", sfx.entries.size(), (unsigned long long)g.total_bases,
               (unsigned long long)g.num_subseqs, K);
        flat.assign(sfx.seq.size(), 0xff);
        const auto t0 = std::chrono::steady_clock::now();
        // -z (hammings.cpp:228) takes effect only here: with -I the reference passes entry 0 and
        // the filter never fires (hammings.cpp:1691-1694)
        if ((rc = gpu_ready())) return rc;
        k4b_set_reference_sensitivity(o.sensitivity);
        rc = k4b_hamm_targeted_z(sfx.seq.data(), sfx.seq.size(), K, o.rhamm, o.crick ? 1 : 0, o.intrainterboth, 0, 0,
                                 flat.data());
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (rc) {
        logmsg(0, "Hamming engine failed (%d): %s", rc, k4b_last_error());
        return rc;
    }
    logmsg(2, "Engine: %.3f s for %llu probe K-mers vs %zu target bases", secs,
           (unsigned long long)g.num_subseqs, sfx.seq.size());
    {   // where a file diff against the reference can legitimately differ (its depth cut, SfxArray.cpp:4487-4494)
        uint64_t deep = 0;
        uint32_t cap = 0;
        if (k4b_last_depth_cut(&deep, &cap) == 0) {
            if (deep)
                logmsg(2, "%llu probe K-mers answered below the not-found value hold a core with more than %u copies in the "
                          "assembly: the reference (-s%d) truncates its search of such cores and may report a larger value "
                          "there; the values written here are the exact minima",
                       (unsigned long long)deep, cap, o.sensitivity);
            else
                logmsg(2, "No probe K-mer holds a core with more than %u copies: the reference's depth cut (-s%d) cannot fire",
                       cap, o.sensitivity);
        }
    }

    // per-loci array H[sum of entry lengths] preset 0xFF (hammings.cpp:2336-2354); loci that the
    // reference's chunk scheduler would not sample stay 0xFF (hammings.cpp:1615-1674 with
    // SfxArray.cpp:4262: every SampleN-th K-mer of each chunk, chunks restart the stride)
    std::vector<uint8_t> h((size_t)g.total_bases, 0xff);
    std::vector<RChrom> rch;
    const uint64_t max_chunk = std::max<uint64_t>(10000, g.total_bases / 10000);
    uint64_t hofs = 0;
    for (const Chrom &c : g.chroms) {
        rch.push_back(RChrom{c.name, c.len});
        if (c.len >= K) {
            uint64_t seq_ofs = 0;
            while (seq_ofs + K <= c.len) {
                const uint64_t chunk_len = std::min<uint64_t>(max_chunk, c.len - seq_ofs);
                if (chunk_len < K) break;
                for (uint64_t ofs = 0; ofs + K <= chunk_len; ofs += (uint64_t)o.sample)
                    h[hofs + seq_ofs + ofs] = flat[c.start + seq_ofs + ofs];
                seq_ofs += 1 + chunk_len - K;
            }
        }
        hofs += c.len;
    }
    if (!o.out_file.empty()) {
        logmsg(2, "Writing Hamming edit distances to file: '%s'", o.out_file.c_str());
        switch (o.resformat) {
            case 0: rc = write_restricted_csv(o.out_file, rch, K, h.data(), o.prefix, err); break;
            case 1: rc = write_restricted_bed(o.out_file, rch, K, o.rhamm, h.data(), o.prefix, err); break;
            default:
                rc = write_restricted_wiggle(o.out_file, rch, K, o.rhamm, o.sensitivity, h.data(), o.prefix, err);
                break;
        }
        if (rc) {
            logmsg(0, "%s", err.c_str());
            return rc;
        }
    }
    // distribution 0..R+1 to the log (hammings.cpp:2535-2587)
    uint64_t hist[256] = {0}, sampled = 0, unsampled = 0;
    hofs = 0;
    for (const Chrom &c : g.chroms) {
        if (c.len >= K)
            for (uint32_t i = 0; i + K <= c.len; ++i) {
                const uint8_t v = h[hofs + i];
                if (v < 0x7f) {
                    hist[v]++;
                    sampled++;
                } else {
                    unsampled++;
                }
            }
        hofs += c.len;
    }
    if (o.sample > 1) logmsg(2, "Sampled %llu K-mers, did not sample %llu", (unsigned long long)sampled, (unsigned long long)unsampled);
    logmsg(2, "Distribution:\nEditDist,Freq,Proportion");
    for (int d = 0; d <= o.rhamm + 1; ++d)
        printf("%d%c,%llu,%1.3f\n", d, d == o.rhamm + 1 ? '+' : ' ', (unsigned long long)hist[d],
               sampled ? hist[d] * 100.0 / (double)sampled : 0.0);
    return kOk;
}

int main(int argc, char **argv) {
    // `ngskit4b hammings ...` style invocation: the subprocess word may lead (ngskit4b.cpp:295-312)
    std::vector<char *> av(argv, argv + argc);
    if (argc > 1 && (!strcmp(argv[1], "hammings") || !strcmp(argv[1], "-phammings"))) av.erase(av.begin() + 1);
    std::vector<std::string> expanded;
    std::string err;
    if (expand_param_files((int)av.size(), av.data(), expanded, err) < 0) {
        printf("%s\n", err.c_str());
        printf("\n%s K-mer Hamming distance generator, Version %s\n", kProg, kVersion);
        print_usage(kProg);
        return 1;
    }
    std::vector<char *> pv;
    for (std::string &s : expanded) pv.push_back(&s[0]);
    pv.push_back(nullptr);
    Options o;
    const int perr = parse_args((int)expanded.size(), pv.data(), o, err);
    if (o.help) {  // --help takes precedence over error reporting, and exits 1 (hammings.cpp:255-265)
        printf("\n%s hammings - K-mer Hamming distance generator, Version %s\nOptions ---\n", kProg, kVersion);
        print_usage(kProg);
        printf("\nNote: Parameters can be entered into a parameter file, one parameter per line.");
        printf("\n      To invoke this parameter file then precede its name with '@'");
        printf("\n      e.g. %s @myparams.txt\n\n", kProg);
        return 1;
    }
    if (o.version) {
        printf("\n%s Version %s\n", kProg, kVersion);
        return 1;
    }
    if (perr) {
        printf("\n%s K-mer Hamming distance generator, Version %s\n", kProg, kVersion);
        printf("%s: %s\n", kProg, err.c_str());
        print_usage(kProg);
        printf("\nUse '-h' to view option and parameter usage\n");
        return 1;
    }
    // ---- validation, in the reference's order (hammings.cpp:274-520) ----
    if (o.file_log_level < 0 || o.file_log_level > 4) {
        printf("\nError: FileLogLevel '-l%d' specified outside of range %d..%d\n", o.file_log_level, 0, 4);
        return 1;
    }
    g_screen_level = o.file_log_level;
    if (!o.log_file.empty()) {
        g_file_level = o.file_log_level;
        g_logf = fopen(o.log_file.c_str(), "a");
        if (!g_logf) {
            printf("\nError: Unable to start diagnostics subsystem\n Most likely cause is that logfile '%s' can't be opened/created\n",
                   o.log_file.c_str());
            return 1;
        }
    }
    logmsg(2, "Version: %s", kVersion);
    if (o.mode < 0 || o.mode > 5) {
        logmsg(0, "Error: Processing mode '-m%d' specified outside of range %d..%d", o.mode, 0, 5);
        return 1;
    }
    if (o.mode <= 2 && (o.K < (int)K4B_MIN_K || o.K > (int)K4B_MAX_K)) {
        logmsg(0, "Error: k-mer sequence length '-K%d' specified outside of range %d..%d", o.K, K4B_MIN_K, K4B_MAX_K);
        return 1;
    }
    if (o.mode == 0) {
        if (o.prefix.size() > 10) {
            logmsg(0, "Prefix \"%s\" length must <= %d chars", o.prefix.c_str(), 10);
            return 1;
        }
        for (char ch : o.prefix)
            if (!isalnum((unsigned char)ch)) {
                logmsg(0, "Prefix \"%s\" may only contain alpha-numeric chars", o.prefix.c_str());
                return 1;
            }
        if (o.sensitivity < 0 || o.sensitivity > 3) {
            logmsg(0, "Error: Restricted hamming sensitivity '-s%d' specified outside of range %d..%d", o.sensitivity, 0, 3);
            return 1;
        }
        if (o.rhamm < 1 || o.rhamm > 10) {
            logmsg(0, "Error: Restricted Hamming limit '-r%d' specified outside of range %d..%d", o.rhamm, 1, 10);
            return 1;
        }
        if (o.resformat < 0 || o.resformat > 2) {
            logmsg(0, "Error: Restricted Hamming output file format '-S%d' specified outside of range %d..%d", o.resformat, 0, 2);
            return 1;
        }
        if (o.K / (o.rhamm + 1) < 4) {
            logmsg(0, "Error: Restricted hamming limit '-r%d' is incompatible with k-mer sequence length '-k%d'", o.rhamm, o.K);
            return 1;
        }
        if (o.sample < 1 || o.sample > 100) {
            logmsg(0, "Error: Sampling '-S%d' specified outside of range 1..%d", o.sample, 100);
            return 1;
        }
        if (o.intrainterboth < 0 || o.intrainterboth > 2) {
            logmsg(0, "Error: Hamming processing Intra/Inter/Both '-z%d' specified outside of range 0..2", o.intrainterboth);
            return 1;
        }
        if (o.K > 500) {  // CSfxArray::LocateHammings rejects longer K-mers (SfxArray.cpp:4255)
            logmsg(0, "Error: restricted Hammings support K-mers of at most 500 bases");
            return 1;
        }
    }
    if (o.mode == 1 || o.mode == 2) {
        if (o.mode == 1) {
            if (o.sweep_start < 0) {
                logmsg(0, "Error: Sweep start '-b%d' must be >= 1", o.sweep_start);
                return 1;
            }
            if (o.sweep_start == 0) o.sweep_start = 1;
            if (o.sweep_end != 0 && o.sweep_end < o.sweep_start) {
                logmsg(0, "Error: Sweep end '-B%d' must be either 0 or >= %d", o.sweep_end, o.sweep_start);
                return 1;
            }
        } else {
            if (o.numnodes < 2 || o.numnodes > 10000 || o.node < 1) {
                logmsg(0, "Error: In distributed processing mode both number of nodes '-n<num>' (2..10000) and node instance "
                          "'-N<node>' must be specified");
                return 1;
            }
        }
        if (o.sample < 1 || o.sample > 10000000) {
            logmsg(0, "Error: Sampling '-S%d' specified outside of range 1..10000000", o.sample);
            return 1;
        }
    }
    std::string out_file = o.out_file;
    if (o.sample > 1 && !o.out_file.empty()) {  // every mode: hammings.cpp:508-509
        logmsg(1, "Warning: When sampling no output to file is supported");
        out_file.clear();
    }
    Options run = o;
    run.out_file = out_file;

    const auto t0 = std::chrono::steady_clock::now();
    int rc = 0;
    if (run.mode == 3) {
        rc = merge_hamming_csv(run.in_file, run.out_file, err);
        if (rc) logmsg(0, "Merge failed: %s", err.c_str());
    } else if (run.mode == 4 || run.mode == 5) {
        rc = run.mode == 4 ? csv_to_bham(run.in_file, run.out_file, err) : bham_to_csv(run.in_file, run.out_file, err);
        if (rc) logmsg(0, "Transform failed: %s", err.c_str());
    } else {
        // -k > 1 with -m1 / -m2: the reference then walks every Nth sweep offset of each thread's block, writes
        // no file and only logs the distribution (hammings.cpp:508-509, 901-904; its sample depends on -T).
        // Here the full job takes seconds, so ALL sweeps are computed and the logged distribution is exact;
        // no file either.
        if ((run.mode == 1 || run.mode == 2) && run.sample > 1)
            logmsg(1, "Warning: sweep sampling (-k%d) is not applied: every sweep offset is computed, the logged "
                      "distribution is the exact one", run.sample);
        const int gpus = run.gpus;
        g_gpu_init = std::async(std::launch::async, [gpus]() {
            const int irc = k4b_gpu_init(gpus, nullptr);
            if (irc) g_gpu_init_err = k4b_last_error();
            return irc;
        });
        rc = run.mode != 0 ? run_exhaustive(run) : run_restricted(run);
        if (g_gpu_init.valid()) g_gpu_init.get();  // an input error returned before the engine was needed
        k4b_gpu_shutdown();
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int code = rc >= 0 ? 0 : 1;
    logmsg(2, "Exit code: %d Total processing time: %.3f seconds", code, secs);
    if (g_logf) fclose(g_logf);
    return code;
}
