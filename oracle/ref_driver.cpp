// TEST INFRASTRUCTURE ONLY - not part of the product path.
// Minimal dispatcher that links the UNMODIFIED reference translation units
// (ngskit4b/hammings.cpp, genbioseq.cpp, kit4bax.cpp, genbiobed.cpp) compiled where they lie
// under /root/reference, so the reference's own `hammings`, `genbioseq`, `index` and
// `genbiobed` subprocesses can be run as the parity oracle.  The real ngskit4b/ngskit4b.cpp
// dispatcher is not used because it would pull in all 54 subprocesses
// (ngskit4b/ngskit4b.cpp:136-191); this file only defines the globals those TUs
// expect (ngskit4b/ngskit4b.h:37-47) and forwards argv.
#include "stdafx.h"
#include <sys/mman.h>
#include <pthread.h>
#include "../libkit4b/commhdrs.h"
#include "ngskit4b.h"

extern int hammings(int argc, char *argv[]);
extern int genbioseq(int argc, char *argv[]);
extern int kingsax(int argc, char *argv[]);
extern int genbiobed(int argc, char *argv[]);

const char *cpszProgVer = "oracle";
CStopWatch gStopWatch;
CDiagnostics gDiagnostics;
CSQLiteSummaries gSQLiteSummaries;
int gExperimentID = 0;
int gProcessID = 0;
int gProcessingID = 0;
char gszProcName[_MAX_FNAME];
static tsSubProcess gTable[] = {
    {"hammings", "hammings", "hammings", hammings},
    {"genbioseq", "genbioseq", "genbioseq", genbioseq},
    {"index", "index", "index", kingsax},
    {"genbiobed", "genbiobed", "genbiobed", genbiobed},
};
tsSubProcess *gpszSubProcess = &gTable[0];

#ifdef K4B_ORACLE_NOSLEEP
// Interposes libc sleep(): the reference main thread sleeps 10 s unconditionally
// after starting its workers (ngskit4b/hammings.cpp:2782-2787).  Used only for
// honest small-input timing of -m1; never for -m0 (SfxArray.cpp:1161-1166 relies on it).
extern "C" unsigned int sleep(unsigned int) { return 0; }
#endif

int main(int argc, char *argv[]) {
    strcpy(gszProcName, "ngskit4b");
    if (argc < 2) {
        fprintf(stderr, "usage: %s hammings|genbioseq|index|genbiobed <flags>\n", argv[0]);
        return 2;
    }
    for (auto &e : gTable) {
        if (!strcmp(argv[1], e.pszName)) {
            gpszSubProcess = &e;
            argv[1] = argv[0];
            return e.SubFunct(argc - 1, argv + 1);
        }
    }
    fprintf(stderr, "unknown subprocess %s\n", argv[1]);
    return 2;
}
