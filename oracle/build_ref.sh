#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own hammings/genbioseq/index/genbiobed
# subprocesses and its HammingDist tool from the sources where they lie under /root/reference (nothing is copied
# into this repo) into oracle/_ref/ (git-ignored, NOT gpurun-ignored so it travels to the
# GPU box).  Recipe follows SURVEY.md Appendix A; plain g++, no autotools.
#   oracle/_ref/ngskit4b_ref          stock behaviour (10 s sleep in -m1, 12 s in -m0)
#   oracle/_ref/ngskit4b_ref_nosleep  libc sleep() interposed to return at once (-m1 timing only)
#   oracle/_ref/hammingdist_ref       the reference's HammingDist tool (HammingDist/HammingDist.cpp, own main)
set -euo pipefail
R=${K4B_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
B=$HERE/_ref
J=${JOBS:-$(nproc)}
if [ ! -d "$R/ngskit4b" ]; then echo "reference tree $R not present; keeping prebuilt $B" >&2; exit 0; fi
if [ -x "$B/ngskit4b_ref" ] && [ -x "$B/ngskit4b_ref_nosleep" ] && [ -x "$B/hammingdist_ref" ] && [ -z "${FORCE:-}" ]; then
  echo "oracle/_ref already built"; exit 0; fi
mkdir -p "$B/obj"
CXXF="-O2 -w -std=c++17 -I$R -I$R/libkit4b -I$R/ngskit4b"
ls $R/libzlib/*.c | xargs -P "$J" -I{} sh -c 'gcc -O2 -w -c {} -o '"$B"'/obj/z_$(basename {} .c).o'
ls $R/libkit4b/*.cpp | grep -v -E '/(stdafx|DSsort|FMIndex|MemAlloc|VisData|conservlib)\.cpp$' | \
  xargs -P "$J" -I{} sh -c 'g++ '"$CXXF"' -c {} -o '"$B"'/obj/k_$(basename {} .cpp).o'
rm -f "$B/libk.a"; ar rcs "$B/libk.a" $B/obj/k_*.o $B/obj/z_*.o
for s in hammings genbioseq kit4bax SQLiteSummaries genbiobed; do
  echo $R/ngskit4b/$s.cpp; done | xargs -P "$J" -I{} sh -c 'g++ '"$CXXF"' -c {} -o '"$B"'/obj/n_$(basename {} .cpp).o'
SQLITE=$(ls /usr/lib/x86_64-linux-gnu/libsqlite3.so.0 2>/dev/null || true)
for v in "" _nosleep; do
  D=""; [ -n "$v" ] && D="-DK4B_ORACLE_NOSLEEP"
  g++ $CXXF $D -c "$HERE/ref_driver.cpp" -o "$B/obj/driver$v.o"
  g++ -o "$B/ngskit4b_ref$v" "$B/obj/driver$v.o" $B/obj/n_hammings.o $B/obj/n_genbioseq.o \
      $B/obj/n_kit4bax.o $B/obj/n_SQLiteSummaries.o $B/obj/n_genbiobed.o "$B/libk.a" $SQLITE -lpthread -lrt -ldl
done
g++ $CXXF -I$R/HammingDist -c "$R/HammingDist/HammingDist.cpp" -o "$B/obj/hammingdist.o"
g++ -o "$B/hammingdist_ref" "$B/obj/hammingdist.o" "$B/libk.a" -lpthread -lrt -ldl
rm -rf "$B/obj" "$B/libk.a"
echo "built $B/ngskit4b_ref, $B/ngskit4b_ref_nosleep and $B/hammingdist_ref"
