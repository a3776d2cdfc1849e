// k4b_seed.cu - seed-and-verify engine for the targeted mode (-m0 -I) on sm_100a.
//
// The reference answers a probe K-mer with min(true minimum, "not found") where "not found" =
// K / CoreLen, CoreLen = K / (R+1) (libkit4b/SfxArray.cpp:4462-4463), and finds the hits by the
// pigeonhole principle: a hit with fewer than K/CoreLen mismatches leaves at least one of the
// K/CoreLen disjoint cores of the probe intact, so it is enough to look up every core exactly
// (suffix-array runs, :4480-4527) and to verify the full K bases of each occurrence (:4567-4581).
// Same idea here, without the suffix array and without the reference's depth cut-offs:
//   * index: every target position whose CoreLen window is pure ACGT is bucketed by the code (or a
//     hash, for long cores) of that window - count, exclusive scan, fill.  Both passes stage the
//     planes of a 2048-position chunk in shared memory once (coalesced) and cut every window out of
//     that copy; an entry is ONE 16-byte record {position, flank signature} = the 16 bases before
//     and the 16 bases after the core, written with a single full-width store.  A bucket RANGE can
//     be selected, so that several GPUs each index (and join) their own share of the buckets;
//   * query: one warp per (probe K-mer, strand, core): the lanes walk the bucket of the core
//     (12 coalesced bytes per occurrence); the mismatches between the signature and the probe's
//     own flanks are a lower bound of the distance, so almost every unrelated occurrence is
//     dismissed without touching the target planes; the survivors are verified over the full K
//     bases (target K-mer cut out of the planes, XOR / OR / POPC); the running minimum of the
//     probe is lowered with one atomicMin.
// Exact for every distance below the "not found" value, which is all the output can show;
// buckets that hold unrelated cores (hash collisions) only cost extra verifications.
// Probes must be pure ACGT (wildcard probes need the reference's substitution rule, :4266-4296;
// they stay on the brute-force engines); target N / InDel count as mismatches, windows across an
// entry boundary are rejected through the valid-start plane.
#include "k4b_kernels.cuh"

#include <stdlib.h>

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace k4b {

constexpr uint32_t kSeedMaxBits = 22;  // at most 4 M buckets

// bits [pos, pos+32) of a plane; pos may be slightly negative (front pad) or run past the end
// (back pad): both pads are part of every image
__device__ __forceinline__ uint32_t plane_bits(const uint32_t *pl, long long pos) {
    const long long wi = pos >> 5;  // floor
    const uint32_t sh = (uint32_t)(pos & 31);
    return __funnelshift_r(__ldg(pl + wi), __ldg(pl + wi + 1), sh);
}

// where window bits come from: the planes in global memory, or a chunk of them staged in shared memory
struct GlobalBits {
    const uint32_t *base;  // logical word 0 of plane 0
    uint32_t stride;       // words between planes
    __device__ __forceinline__ explicit GlobalBits(const ImageView &img) : base(img.base), stride(img.stride) {}
    __device__ __forceinline__ GlobalBits(const uint32_t *b, uint32_t s) : base(b), stride(s) {}
    __device__ __forceinline__ uint32_t get(int p, long long pos) const { return plane_bits(base + (size_t)p * stride, pos); }
};
struct SmemBits {
    const uint32_t *pl[3];
    long long origin;  // sequence position of bit 0 of word 0 of the staged copy (a multiple of 32)
    __device__ __forceinline__ uint32_t get(int p, long long pos) const {
        const uint32_t lb = (uint32_t)(pos - origin);
        const uint32_t *w = pl[p] + (lb >> 5);
        return __funnelshift_r(w[0], w[1], lb & 31u);
    }
};

// bucket of the core starting at pos; ok = the window is pure ACGT (plane 2 clear; EOS and the
// pads have it set).  core_len <= 11 with bits == 2*core_len: the code itself, else a hash.
template <class Src>
__device__ __forceinline__ uint32_t core_bucket(const Src &src, long long pos, uint32_t core_len, uint32_t bits,
                                                bool &ok) {
    uint32_t h = 0x9e3779b9u, bad = 0, key = 0;
    for (uint32_t o = 0; o < core_len; o += 32) {
        const uint32_t n = core_len - o < 32 ? core_len - o : 32;
        const uint32_t m = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
        const uint32_t w0 = src.get(0, pos + o) & m;
        const uint32_t w1 = src.get(1, pos + o) & m;
        bad |= src.get(2, pos + o) & m;
        key = w0 | (w1 << (core_len & 31));  // used only when core_len <= 11
        h = (h ^ w0) * 0x85ebca6bu;
        h ^= h >> 13;
        h = (h ^ w1) * 0xc2b2ae35u;
        h ^= h >> 16;
    }
    ok = bad == 0;
    return (2 * core_len == bits) ? key : (h & ((1u << bits) - 1u));
}

// Flank signature of a core: the 16 bases before it (bits 0..15) and the 16 bases after it (bits
// 16..31), x = their plane-0 bits, y = their plane-1 bits.  Planes 0/1 only: a non-ACGT symbol then
// looks like a base, which can only LOWER the mismatch count taken from the signature.
template <class Src>
__device__ __forceinline__ uint2 flank_sig(const Src &src, long long core_pos, uint32_t core_len) {
    const long long before = core_pos - 16, after = core_pos + core_len;  // the front pad makes pos < 16 readable
    return make_uint2((src.get(0, before) & 0xffffu) | (src.get(0, after) << 16),
                      (src.get(1, before) & 0xffffu) | (src.get(1, after) << 16));
}

// Index passes.  A CTA takes kSeedChunk consecutive target positions and stages the words of the
// three planes that their windows touch (one word before the chunk for the 16 bases ahead of a
// core, kSeedHalo words after it for core + flank) in shared memory: the planes are read once,
// coalesced, instead of 14 scattered words per position.  FILL = false counts the cores per
// bucket; FILL = true takes a slot from the bucket's cursor and writes the 16-byte entry.
constexpr int kSeedChunk = 2048;                       // positions per CTA: 8 per thread
constexpr int kSeedHalo = 10;                          // words after the chunk: cores of up to 250 bases + 16 + slack
constexpr int kSeedStage = kSeedChunk / 32 + 1 + kSeedHalo;
template <bool FILL>
__global__ void __launch_bounds__(256) seed_scan_kernel(ImageView t, uint32_t n_pos, uint32_t core_len, uint32_t bits,
                                                        uint32_t b_lo, uint32_t b_hi, uint32_t *__restrict__ counters,
                                                        uint4 *__restrict__ ent) {
    __shared__ uint32_t stage[3][kSeedStage + 1];
    const uint32_t base = blockIdx.x * (uint32_t)kSeedChunk;
    const long long w0 = (long long)(base >> 5) - 1;  // the front pad makes word -1 readable
    for (uint32_t i = threadIdx.x; i < 3u * (kSeedStage + 1); i += 256) {
        const uint32_t p = i / (kSeedStage + 1), w = i - p * (kSeedStage + 1);
        stage[p][w] = __ldg(t.plane((int)p) + w0 + w);
    }
    __syncthreads();
    SmemBits src;
    src.pl[0] = stage[0];
    src.pl[1] = stage[1];
    src.pl[2] = stage[2];
    src.origin = w0 * 32;
#pragma unroll 2
    for (uint32_t j = 0; j < (uint32_t)kSeedChunk / 256; ++j) {
        const uint32_t pos = base + j * 256 + threadIdx.x;  // a warp covers 32 consecutive positions: broadcast reads
        if (pos >= n_pos) break;
        bool ok;
        const uint32_t b = core_bucket(src, pos, core_len, bits, ok);
        if (!ok || b < b_lo || b >= b_hi) continue;
        if (!FILL) {
            atomicAdd(counters + b, 1u);
        } else {
            const uint2 sg = flank_sig(src, pos, core_len);
            const uint32_t slot = atomicAdd(counters + b, 1u);
            ent[slot] = make_uint4(pos, sg.x, sg.y, b);
        }
    }
}

// ---- index by two partition passes (buckets of at most 16 bits) ------------------------------
// The fill above costs one returning L2 atomic and one scattered 16-byte store PER ENTRY.  Here the
// entries are moved twice, but in runs: pass A cuts them out of the planes exactly as the fill does and
// groups them by the TOP 8 bits of the bucket into a scratch array laid out like the index (coarse
// partition p starts at off[p << nlow]); pass B reads a coarse partition in tiles and groups by the
// remaining nlow = bits - 8 bits into the final bucket runs.  A CTA groups its tile of 2048 entries in
// shared memory (ranks from a shared-memory histogram), reserves ONE run per digit with one global atomic
// (a warp's 32 cursors share a line: 8 atomic requests per CTA instead of 2048) and copies the grouped
// tile out with consecutive lanes on consecutive slots.  3x the bytes of the fill, all of them coalesced.
// Order inside a bucket is arbitrary in both builds (the queries take minima).
constexpr int kPartPer = 8;  // entries per thread: a tile is THREADS * 8 entries (2048 or 4096)

// exclusive prefix sums over the first 256 threads of the CTA (one value each); every thread calls
__device__ __forceinline__ uint32_t cta_excl_scan_256(uint32_t v, uint32_t *ws, uint32_t &total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (uint32_t)d) inc += n;
    }
    if (warp < 8 && lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
        const uint32_t s = ws[w];
        all += s;
        if (w < warp) before += s;
    }
    total = all;
    __syncthreads();  // ws is free again
    return before + inc - v;
}

struct PartShared {
    uint32_t hist[256], start[256], gbase[256], ws[8];
};

// groups the entries the CTA's threads hold (e[u].w = bucket < 2^16, bit u of `live` = e[u] is an entry) by
// digit_of(bucket) < 256 in `out` (a tile of entries) and reserves one run per group: reserve(digit, size) is the
// slot of its first entry.  Returns the number of entries; e is dead afterwards.  part_copy_out then writes the
// groups to their runs, consecutive lanes on consecutive slots.
template <class DigitOf, class Reserve>
__device__ __forceinline__ uint32_t part_group(uint4 (&e)[kPartPer], uint32_t live, DigitOf digit_of, Reserve reserve,
                                               uint4 *out, PartShared &sh) {
    const uint32_t tid = threadIdx.x;
    if (tid < 256) sh.hist[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPartPer; ++u)
        if ((live >> u) & 1u) e[u].w |= atomicAdd(&sh.hist[digit_of(e[u].w)], 1u) << 16;  // rank inside the group, < 4096
    __syncthreads();
    uint32_t total;
    const uint32_t h = tid < 256 ? sh.hist[tid] : 0u;
    const uint32_t ex = cta_excl_scan_256(h, sh.ws, total);
    uint32_t g = 0;
    if (tid < 256) {
        sh.start[tid] = ex;
        if (h) g = reserve(tid, h);  // the returning atomic is in flight while the tile is grouped
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPartPer; ++u)
        if ((live >> u) & 1u) {
            const uint32_t b = e[u].w & 0xffffu;
            out[sh.start[digit_of(b)] + (e[u].w >> 16)] = make_uint4(e[u].x, e[u].y, e[u].z, b);
        }
    if (tid < 256) sh.gbase[tid] = g;  // read after the barrier of part_copy_out
    return total;
}
template <int THREADS, class DigitOf>
__device__ __forceinline__ void part_copy_out(uint32_t total, DigitOf digit_of, uint4 *__restrict__ dst, const uint4 *out,
                                              const PartShared &sh) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += THREADS) {
        const uint4 v = out[i];
        const uint32_t d = digit_of(v.w);
        dst[sh.gbase[d] + (i - sh.start[d])] = v;
    }
}
template <int THREADS, class DigitOf, class Reserve>
__device__ __forceinline__ void part_tile_out(uint4 (&e)[kPartPer], uint32_t live, DigitOf digit_of, Reserve reserve,
                                              uint4 *__restrict__ dst, uint4 *out, PartShared &sh) {
    part_copy_out<THREADS>(part_group(e, live, digit_of, reserve, out, sh), digit_of, dst, out, sh);
}

extern __shared__ uint4 part_out[];  // THREADS * 8 entries

// pass A: positions -> entries grouped by the top 8 bits of the bucket.  cur_a: 256 zeroed counters.
// entry of the core at tile offset lt (position pos) cut out of a staged copy of the planes that starts 32 bases
// before the tile: three words per plane hold the 16 + core + 16 bases of a core of at most 8 bases.  false: the
// core holds a non-ACGT symbol or its bucket lies outside [b_lo, b_hi)
__device__ __forceinline__ bool part_fast_entry(const uint32_t *s0, const uint32_t *s1, const uint32_t *s2, uint32_t lt,
                                                uint32_t pos, uint32_t core_len, uint32_t b_lo, uint32_t b_hi, uint4 &e) {
    const uint32_t lb = lt + 16;  // bit of base pos - 16 in the staged copy
    const uint32_t wi = lb >> 5, sh5 = lb & 31u, m = (1u << core_len) - 1u;
    // per plane: first word = bases pos-16 .. pos+15, second = bases pos+16 .. pos+47
    const uint32_t a0 = __funnelshift_r(s0[wi], s0[wi + 1], sh5), a1 = __funnelshift_r(s0[wi + 1], s0[wi + 2], sh5);
    const uint32_t c0 = __funnelshift_r(s1[wi], s1[wi + 1], sh5), c1 = __funnelshift_r(s1[wi + 1], s1[wi + 2], sh5);
    const uint32_t n0 = __funnelshift_r(s2[wi], s2[wi + 1], sh5), n1 = __funnelshift_r(s2[wi + 1], s2[wi + 2], sh5);
    const uint32_t b = (__funnelshift_r(a0, a1, 16) & m) | ((__funnelshift_r(c0, c1, 16) & m) << core_len);
    if ((__funnelshift_r(n0, n1, 16) & m) != 0 || b < b_lo || b >= b_hi) return false;
    e = make_uint4(pos, (a0 & 0xffffu) | (__funnelshift_r(a0, a1, 16 + core_len) << 16),
                   (c0 & 0xffffu) | (__funnelshift_r(c0, c1, 16 + core_len) << 16), b);
    return true;
}

// FAST: bucket and flank signature cut out of three staged words per plane (the 16 + core + 16 bases of a core of at
// most 8 bases span 40 bits) instead of one two-word window per field
template <int THREADS, bool FAST>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) seed_part_a_kernel(ImageView t, uint32_t n_pos, uint32_t core_len,
                                                                              uint32_t bits, uint32_t b_lo, uint32_t b_hi,
                                                                              const uint32_t *__restrict__ off,
                                                                              uint32_t *__restrict__ cur_a,
                                                                              uint4 *__restrict__ tmp) {
    constexpr int kTile = THREADS * kPartPer;
    constexpr int kStage = kTile / 32 + 1 + kSeedHalo + 1;  // words per plane, as in seed_scan_kernel
    __shared__ uint32_t stage[3][kStage];
    __shared__ PartShared sh;
    const uint32_t tid = threadIdx.x;
    const uint32_t base = blockIdx.x * (uint32_t)kTile;
    const long long w0 = (long long)(base >> 5) - 1;
    for (uint32_t i = tid; i < 3u * kStage; i += THREADS) {
        const uint32_t p = i / kStage, w = i - p * kStage;
        stage[p][w] = __ldg(t.plane((int)p) + w0 + w);
    }
    __syncthreads();
    SmemBits src;
    src.pl[0] = stage[0];
    src.pl[1] = stage[1];
    src.pl[2] = stage[2];
    src.origin = w0 * 32;
    uint4 e[kPartPer];
    uint32_t live = 0;
#pragma unroll
    for (int j = 0; j < kPartPer; ++j) {
        const uint32_t pos = base + (uint32_t)j * THREADS + tid;
        e[j] = make_uint4(0, 0, 0, 0);
        if (FAST) {
            if (pos < n_pos && part_fast_entry(stage[0], stage[1], stage[2], pos - base, pos, core_len, b_lo, b_hi, e[j])) live |= 1u << j;
        } else if (pos < n_pos) {
            bool ok;
            const uint32_t b = core_bucket(src, pos, core_len, bits, ok);
            if (ok && b >= b_lo && b < b_hi) {
                const uint2 sg = flank_sig(src, pos, core_len);
                e[j] = make_uint4(pos, sg.x, sg.y, b);
                live |= 1u << j;
            }
        }
    }
    const uint32_t nlow = bits - 8;
    part_tile_out<THREADS>(
        e, live, [nlow](uint32_t b) { return b >> nlow; },
        [&](uint32_t d, uint32_t n) { return __ldg(off + ((size_t)d << nlow)) + atomicAdd(cur_a + d, n); }, tmp, part_out,
        sh);
}

// pass B: tile `blockIdx.x` of the coarse partitions (tiles are numbered partition by partition; every CTA
// derives the numbering from the 257 partition bounds) -> final bucket runs.  cur: a copy of off.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) seed_part_b_kernel(uint32_t bits, const uint32_t *__restrict__ off,
                                                                              uint32_t *__restrict__ cur,
                                                                              const uint4 *__restrict__ tmp,
                                                                              uint4 *__restrict__ ent) {
    constexpr uint32_t kTile = THREADS * kPartPer;
    __shared__ PartShared sh;
    __shared__ uint32_t s_part, s_first, s_n;
    const uint32_t tid = threadIdx.x, nlow = bits - 8;
    uint32_t lo = 0, hi = 0;
    if (tid < 256) {
        lo = __ldg(off + ((size_t)tid << nlow));
        hi = __ldg(off + ((size_t)(tid + 1) << nlow));
    }
    const uint32_t tiles = (hi - lo + kTile - 1) / kTile;
    uint32_t total;
    const uint32_t ex = cta_excl_scan_256(tiles, sh.ws, total);
    if (blockIdx.x >= total) return;  // the grid is an upper bound
    if (blockIdx.x >= ex && blockIdx.x < ex + tiles) {  // tiles = 0 outside the first 256 threads
        const uint32_t first = lo + (blockIdx.x - ex) * kTile;
        s_part = tid;
        s_first = first;
        s_n = hi - first < kTile ? hi - first : kTile;
    }
    __syncthreads();
    const uint32_t part = s_part, first = s_first, n = s_n;
    uint4 e[kPartPer];
    uint32_t live = 0;
#pragma unroll
    for (int j = 0; j < kPartPer; ++j) {
        const uint32_t i = (uint32_t)j * THREADS + tid;
        e[j] = make_uint4(0, 0, 0, 0);
        if (i < n) {
            e[j] = __ldg(tmp + first + i);
            live |= 1u << j;
        }
    }
    const uint32_t low_mask = (1u << nlow) - 1u;
    uint32_t *cur_p = cur + ((size_t)part << nlow);
    part_tile_out<THREADS>(
        e, live, [low_mask](uint32_t b) { return b & low_mask; },
        [&](uint32_t d, uint32_t m) { return atomicAdd(cur_p + d, m); }, ent, part_out, sh);
}

// count pass with the counters in shared memory: one persistent CTA per SM keeps ALL buckets (at most 2^16) as
// 16-bit halves of 2^15 words and counts its chunks of 8192 positions with shared-memory atomics; the global
// counters see one add per bucket and CTA at the end instead of one per position (the one-per-position pass runs
// at the rate of the L2 atomics: 500 M in 3.5 ms).  A half that reaches 2^15 hands 2^15 on to the global counter
// at once (the thread that sees 0x7fff -> 0x8000 subtracts it again), so no half can carry into its neighbour.
// The planes of the next chunk are fetched (one word per thread) while the current one is counted.
constexpr int kCountChunk = 8192, kCountThreads = 1024;
constexpr int kCountStage = kCountChunk / 32 + 2;
extern __shared__ uint32_t count_half[];
__global__ void __launch_bounds__(kCountThreads, 1) seed_count_smem_kernel(ImageView t, uint32_t n_pos, uint32_t core_len,
                                                                           uint32_t bits, uint32_t b_lo, uint32_t b_hi,
                                                                           uint32_t *__restrict__ cnt) {
    static_assert(3 * kCountStage <= kCountThreads, "one staging word per thread");
    __shared__ uint32_t stage[3][kCountStage];
    const uint32_t tid = threadIdx.x, words = 1u << (bits - 1), m = (1u << core_len) - 1u;
    for (uint32_t i = tid; i < words; i += kCountThreads) count_half[i] = 0;
    const uint32_t n_chunks = (n_pos + kCountChunk - 1) / kCountChunk;
    const bool loader = tid < 3 * kCountStage;
    const uint32_t lp = tid / kCountStage, lw = tid - lp * kCountStage;
    uint32_t chunk = blockIdx.x, nxt = 0;
    if (loader && chunk < n_chunks) nxt = __ldg(t.plane((int)lp) + (size_t)chunk * (kCountChunk / 32) + lw);
    while (chunk < n_chunks) {
        __syncthreads();  // the previous chunk has been counted (first round: the counters are zeroed)
        if (loader) stage[lp][lw] = nxt;
        __syncthreads();
        const uint32_t next = chunk + gridDim.x;
        if (loader && next < n_chunks) nxt = __ldg(t.plane((int)lp) + (size_t)next * (kCountChunk / 32) + lw);
        const uint32_t base = chunk * (uint32_t)kCountChunk;
#pragma unroll
        for (int j = 0; j < kCountChunk / kCountThreads; ++j) {
            const uint32_t lb = (uint32_t)j * kCountThreads + tid;
            if (base + lb < n_pos) {
                const uint32_t wi = lb >> 5, sh5 = lb & 31u;
                const uint32_t k0 = __funnelshift_r(stage[0][wi], stage[0][wi + 1], sh5) & m;
                const uint32_t k1 = __funnelshift_r(stage[1][wi], stage[1][wi + 1], sh5) & m;
                const uint32_t bad = __funnelshift_r(stage[2][wi], stage[2][wi + 1], sh5) & m;
                const uint32_t b = k0 | (k1 << core_len);
                if (bad == 0 && b >= b_lo && b < b_hi) {
                    const uint32_t hs = (b & 1u) << 4;
                    const uint32_t old = atomicAdd(&count_half[b >> 1], 1u << hs);
                    if (((old >> hs) & 0xffffu) == 0x7fffu) {
                        atomicSub(&count_half[b >> 1], 0x8000u << hs);
                        atomicAdd(cnt + b, 0x8000u);
                    }
                }
            }
        }
        chunk = next;
    }
    __syncthreads();
    for (uint32_t i = tid; i < words; i += kCountThreads) {
        const uint32_t w = count_half[i];
        if (w & 0xffffu) atomicAdd(cnt + 2 * i, w & 0xffffu);
        if (w >> 16) atomicAdd(cnt + 2 * i + 1, w >> 16);
    }
}

// pass B, persistent: 4 CTAs per SM walk the tiles with a stride of the grid; the loads of a CTA's next tile are
// issued as soon as the current one is grouped in shared memory, so they overlap its copy-out (ncu of the
// one-tile-per-CTA kernel: 23 of 33 stall cycles per issue wait for the tile's loads, issue slots 21 % busy).
__global__ void __launch_bounds__(256, 4) seed_part_b_persistent_kernel(uint32_t bits, const uint32_t *__restrict__ off,
                                                                        uint32_t *__restrict__ cur,
                                                                        const uint4 *__restrict__ tmp,
                                                                        uint4 *__restrict__ ent) {
    constexpr uint32_t kTile = 256 * kPartPer;
    __shared__ PartShared sh;
    __shared__ uint32_t s_ts[257], s_lo[257];  // first tile number and first entry of every coarse partition
    const uint32_t tid = threadIdx.x, nlow = bits - 8, low_mask = (1u << nlow) - 1u;
    uint32_t n_tiles;
    {
        const uint32_t lo = __ldg(off + ((size_t)tid << nlow)), hi = __ldg(off + ((size_t)(tid + 1) << nlow));
        const uint32_t ex = cta_excl_scan_256((hi - lo + kTile - 1) / kTile, sh.ws, n_tiles);
        s_ts[tid] = ex;
        s_lo[tid] = lo;
        if (tid == 255) {
            s_ts[256] = n_tiles;
            s_lo[256] = hi;
        }
    }
    __syncthreads();
    uint4 e[kPartPer];
    uint32_t live, part;
    auto fetch = [&](uint32_t tile) {  // partition of the tile (the last one that starts at or before it), its entries
        uint32_t a = 0, b = 256;
        while (b - a > 1) {
            const uint32_t mid = (a + b) >> 1;
            if (s_ts[mid] <= tile) a = mid;
            else b = mid;
        }
        const uint32_t first = s_lo[a] + (tile - s_ts[a]) * kTile, left = s_lo[a + 1] - first;
        const uint32_t n = left < kTile ? left : kTile;
        part = a;
        live = 0;
#pragma unroll
        for (int j = 0; j < kPartPer; ++j) {
            const uint32_t i = (uint32_t)j * 256 + tid;
            e[j] = make_uint4(0, 0, 0, 0);
            if (i < n) {
                e[j] = __ldg(tmp + first + i);
                live |= 1u << j;
            }
        }
    };
    auto digit_of = [low_mask](uint32_t b) { return b & low_mask; };
    uint32_t tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile);
    while (tile < n_tiles) {
        uint32_t *cur_p = cur + ((size_t)part << nlow);
        const uint32_t total = part_group(e, live, digit_of, [&](uint32_t d, uint32_t m) { return atomicAdd(cur_p + d, m); },
                                          part_out, sh);
        tile += gridDim.x;
        if (tile < n_tiles) fetch(tile);  // in flight during the copy-out
        part_copy_out<256>(total, digit_of, ent, part_out, sh);
        __syncthreads();  // the tile and its tables are free again
    }
}

// mismatches between the K-mer of `a` at pa and the K-mer of `b` at pb (b may hold non-ACGT
// symbols, a is pure ACGT inside the window); stops counting once `limit` is reached
__device__ __forceinline__ uint32_t kmer_mismatches(const ImageView &a, long long pa, const ImageView &b,
                                                    long long pb, uint32_t K, bool three, uint32_t limit) {
    uint32_t d = 0;
    for (uint32_t o = 0; o < K && d < limit; o += 32) {
        uint32_t m = plane_bits(a.plane(0), pa + o) ^ plane_bits(b.plane(0), pb + o);
        m |= plane_bits(a.plane(1), pa + o) ^ plane_bits(b.plane(1), pb + o);
        if (three) m |= plane_bits(b.plane(2), pb + o);
        if (K - o < 32) m &= (1u << (K - o)) - 1u;
        d += __popc(m);
    }
    return d;
}

__device__ __forceinline__ uint32_t seed_entry_of(const SeedSelfRules &r, uint32_t pos) {
    uint32_t lo = 0, hi = r.n_ent;  // largest i with ent_starts[i] <= pos
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(r.ent_starts + mid) <= pos) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ---- what both query kernels share -------------------------------------------------------
struct SeedCtx {  // launch-wide constants
    ImageView q, rcq, t;
    uint32_t K, core_len, n_cores, bits, clamp;
    uint32_t b_lo, b_hi;  // bucket range this launch answers (the index holds only these buckets)
    // depth watch (nullable): deep[p] = 1 for every probe position p that holds a core whose bucket has
    // more than depth_cap entries - where the reference truncates its search (SfxArray.cpp:4487-4494)
    uint32_t depth_cap;
    uint8_t *deep;
    int strands, three, q_impure;
    int core_only;        // >= 0: this pass handles the items of this core only (phase schedule of the join)
    uint32_t done_below;  // items of probe K-mers whose minimum is <= this are skipped (0 outside the phases)
    SeedSelfRules self;
};
struct SeedItem;
// the probe planes an item reads: the forward image or the reverse-complemented one (same geometry).
// Selected as a scalar pointer: brace-initialising an image-holding struct from `strand ? cx.rcq : cx.q`
// made nvcc 12.9 drop the selection and always read the forward planes.
__device__ __forceinline__ GlobalBits probe_bits(const SeedCtx &cx, uint32_t strand) {
    return GlobalBits(strand ? cx.rcq.base : cx.q.base, cx.q.stride);
}
struct SeedItem {  // one (probe K-mer, strand, core)
    uint32_t p, strand, c;  // probe position, 0 sense / 1 antisense, core number
    long long pp;           // position of the (reverse-complemented) K-mer in its image
    uint32_t q0, q1, m;     // the probe's own flank signature of the core and the mask of what lies inside the K-mer
    uint32_t cur;           // running minimum when the item was set up (<= clamp)
};

// which of the 16 + 16 flank bases of core number c lie inside the K-mer: the LAST nl of the 16 bases before
// the core (bits 16-nl .. 15) and the FIRST nr of the 16 after it (bits 16 .. 16+nr-1); depends on c only
__device__ __forceinline__ uint32_t seed_flank_mask(const SeedCtx &cx, uint32_t c) {
    const uint32_t shift = c * cx.core_len;
    const uint32_t nl = shift < 16 ? shift : 16u;
    const uint32_t after = cx.K - shift - cx.core_len;
    const uint32_t nr = after < 16 ? after : 16u;
    const uint32_t ml = nl ? (0xffffu << (16 - nl)) & 0xffffu : 0u;
    const uint32_t mr = nr == 16 ? 0xffffu : ((1u << nr) - 1u);
    return ml | (mr << 16);
}

// decodes item `sub` of probe p; false = nothing to do (no K-mer, wildcard K-mer, already at 0)
__device__ __forceinline__ bool seed_item_setup(const SeedCtx &cx, uint32_t p, uint32_t sub,
                                                const uint32_t *__restrict__ best, SeedItem &it) {
    if (!((cx.q.valid()[p >> 5] >> (p & 31)) & 1u)) return false;  // no K-mer of one entry starts here
    if (cx.q_impure) {  // probe K-mers holding N / InDel are left to the brute-force engines
        uint32_t bad = 0;
        for (uint32_t o = 0; o < cx.K; o += 32) {
            uint32_t w = plane_bits(cx.q.plane(2), (long long)p + o);
            if (cx.K - o < 32) w &= (1u << (cx.K - o)) - 1u;
            bad |= w;
        }
        if (bad) return false;
    }
    it.p = p;
    it.strand = sub / cx.n_cores;
    it.c = sub - it.strand * cx.n_cores;
    // the reverse complement of the K-mer at p is the K-mer at len-K-p of the reverse-complemented planes
    it.pp = it.strand ? (long long)cx.q.len - cx.K - p : p;
    it.cur = __ldg(best + p);            // running minimum so far (other cores / strands / launches)
    if (it.cur > cx.clamp) it.cur = cx.clamp;  // only distances below the "not found" value matter
    if (it.cur <= cx.done_below) return false;  // 0: nothing lower exists; phase j: settled, see the join
    const uint32_t shift = it.c * cx.core_len;
    // the probe's own flanks of this core, cut to what lies inside the K-mer (masked once here, so that
    // a test against a pre-masked entry needs no AND)
    it.m = seed_flank_mask(cx, it.c);
    const uint2 qs = flank_sig(probe_bits(cx, it.strand), it.pp + shift, cx.core_len);
    it.q0 = qs.x & it.m;
    it.q1 = qs.y & it.m;
    return true;
}

// mismatches between an entry's flank signature and the probe's own flanks: a lower bound of the
// distance (3 LOP3 + 1 POPC); almost every unrelated entry is dismissed by it
__device__ __forceinline__ uint32_t seed_sig_bound(const SeedItem &it, uint2 sg) {
    return __popc(((sg.x ^ it.q0) | (sg.y ^ it.q1)) & it.m);
}

// full verification of the entry whose core sits at target position tpos; lowers `mine`
__device__ __noinline__ void seed_verify(const SeedCtx &cx, const SeedItem &it, uint32_t tpos, uint32_t &mine) {
    const long long ts = (long long)tpos - (long long)it.c * cx.core_len;
    if (ts < 0 || ts > (long long)cx.t.len - cx.K) return;
    const ImageView &img = it.strand ? cx.rcq : cx.q;
    const uint32_t d = kmer_mismatches(img, it.pp, cx.t, ts, cx.K, cx.three != 0, mine);
    // windows across an entry boundary are no K-mers: checked only for the rare improvement
    if (d < mine && ((cx.t.valid()[ts >> 5] >> (ts & 31)) & 1u)) {
        // probes drawn from the assembly itself: an exact sense-strand hit at the probe's own
        // position is no hit (SfxArray.cpp:4418-4419, :4585-4594), and -z lets an exact sense
        // hit count only inside (1) / outside (2) the probe's own entry (:4421-4426, :4597-4601)
        if (cx.self.on && it.strand == 0 && d == 0) {
            if (ts == (long long)it.p) return;
            if (cx.self.zfilt) {
                const bool same = seed_entry_of(cx.self, it.p) == seed_entry_of(cx.self, (uint32_t)ts);
                if (cx.self.zfilt == 1 ? !same : same) return;
            }
        }
        mine = d;
    }
}

// ---- query, warp per item: one warp per (probe position, strand, core) streams the bucket ----
__global__ void __launch_bounds__(256) seed_query_kernel(SeedCtx cx, const uint32_t *__restrict__ off,
                                                         const uint4 *__restrict__ ent, uint32_t q_begin,
                                                         uint32_t q_end, uint32_t *__restrict__ best,
                                                         unsigned long long *__restrict__ occ) {
    const unsigned long long warp_id = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t per_probe = (uint32_t)cx.strands * cx.n_cores;
    const unsigned long long pi = warp_id / per_probe;
    if (pi >= (unsigned long long)(q_end - q_begin)) return;
    SeedItem it;
    if (!seed_item_setup(cx, q_begin + (uint32_t)pi, (uint32_t)(warp_id - pi * per_probe), best, it)) return;
    bool ok;
    const uint32_t b = core_bucket(probe_bits(cx, it.strand), it.pp + (long long)it.c * cx.core_len, cx.core_len, cx.bits,
                                   ok);
    if (!ok || b < cx.b_lo || b >= cx.b_hi) return;  // !ok cannot happen for pure-ACGT probes; kept for safety
    const uint32_t lo = __ldg(off + b), hi = __ldg(off + b + 1);
    if (occ && lane == 0) atomicAdd(occ + (blockIdx.x & (kSeedOccSlots - 1)), (unsigned long long)(hi - lo));
    if (cx.deep && lane == 0 && hi - lo > cx.depth_cap) cx.deep[it.p] = 1;
    uint32_t mine = it.cur;
    for (uint32_t i = lo + lane; i < hi && mine; i += 32) {
        const uint4 e = __ldg(ent + i);
        if (seed_sig_bound(it, make_uint2(e.y, e.z)) < mine) seed_verify(cx, it, e.x, mine);
    }
    mine = __reduce_min_sync(0xffffffffu, mine);
    if (lane == 0 && mine < it.cur) atomicMin(best + it.p, mine);
}

// ---- query, bucket-major join --------------------------------------------------------------
// With a few thousand entries per bucket and hundreds of items per bucket, streaming the bucket
// once per item is HBM-bound (16 B per entry test).  Here the items of a probe chunk are sorted by
// bucket; a CTA takes the items of ONE bucket, stages the bucket tile by tile in shared memory and lets
// its warps test their items against the tile: one global read of an entry serves up to ITEMS entry tests.
//
// Phase schedule (whole index on this device): the items of core 0 of every probe K-mer are joined first,
// then core 1, ...  An alignment with d mismatches leaves at least n_cores - d cores intact, so once
// cores 0 .. j-1 are done every alignment with fewer than j mismatches has been seen: a probe K-mer whose
// minimum is <= j needs none of the cores j, j+1, ... any more.  On BASELINE config 4 (half the probes are
// 3 % mutated copies) that drops a third of the entry tests; on unrelated probes it costs three more small
// sorts.  With bucket shards (several GPUs) a rank sees only its share of the hits, so the rule does not
// apply there and all cores go in one pass.
// bucket of every item of the probe chunk [q0, q0+n_probes); inactive items sort to the end.
// cx.core_only >= 0 (phase schedule, see launch_seed_query): items of that core only, n_probes * strands of
// them; an item that is left out only because its probe is already settled for this phase (0 < minimum <=
// cx.done_below) still reports a deep bucket to the depth watch, as it would have when processed.
__global__ void __launch_bounds__(256) seed_item_keys_kernel(SeedCtx cx, uint32_t q0, uint32_t n_probes,
                                                             const uint32_t *__restrict__ best,
                                                             const uint32_t *__restrict__ off,
                                                             uint32_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per_probe = (uint32_t)cx.strands * cx.n_cores;
    const uint32_t per_pass = cx.core_only >= 0 ? (uint32_t)cx.strands : per_probe;
    if (i >= n_probes * per_pass) return;
    const uint32_t pl = i / per_pass;
    const uint32_t sub = cx.core_only >= 0 ? (i - pl * per_pass) * cx.n_cores + (uint32_t)cx.core_only : i - pl * per_pass;
    SeedItem it;
    uint32_t key = 1u << cx.bits;  // inactive: sorts behind every bucket
    SeedCtx probe_cx = cx;
    probe_cx.done_below = 0;  // decode first, the phase rule is applied below
    if (seed_item_setup(probe_cx, q0 + pl, sub, best, it)) {
        bool ok;
        const uint32_t b = core_bucket(probe_bits(cx, it.strand), it.pp + (long long)it.c * cx.core_len, cx.core_len,
                                       cx.bits, ok);
        if (ok && b >= cx.b_lo && b < cx.b_hi) {
            if (it.cur > cx.done_below) key = b;
            else if (cx.deep && __ldg(off + b + 1) - __ldg(off + b) > cx.depth_cap) cx.deep[it.p] = 1;
        }
    }
    keys[i] = key;
    ids[i] = pl * per_probe + sub;
}

// ONE BUCKET PER CTA: CTA b finds the items of bucket b_lo + b in the sorted keys (two binary searches), takes
// them in batches of ITEMS and stages the bucket, TILE entries at a time, once per batch.  (Round 2 first gave
// every CTA 64 consecutive sorted items whatever their buckets: a bucket's items were then split between CTAs,
// each staging the whole bucket again and dealing a ragged handful of items to its 8 warps - 53.7 ms on
// BASELINE config 4 against 33 ms now, 0.67 of the POPC peak.)  With the phase schedule a pass has
// strands * probes / buckets items per bucket (61 on config 4): one batch, every bucket staged once per pass.
template <int ITEMS, int TILE>
__global__ void __launch_bounds__(256, 4) seed_join_kernel(SeedCtx cx, const uint32_t *__restrict__ off,
                                                                   const uint4 *__restrict__ ent, uint32_t q0,
                                                                   uint32_t n_items, const uint32_t *__restrict__ keys,
                                                                   const uint32_t *__restrict__ ids,
                                                                   uint32_t *__restrict__ best,
                                                                   unsigned long long *__restrict__ occ) {
    __shared__ uint32_t tile_pos[TILE];
    __shared__ uint2 tile_sig[TILE];
    __shared__ SeedItem items[ITEMS];
    __shared__ uint32_t s_mine[ITEMS];  // running minimum of the item; 0: nothing (left) to do
    __shared__ uint32_t s_cur[ITEMS];   // its minimum at set-up; 0: the item is inactive
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t per_probe = (uint32_t)cx.strands * cx.n_cores;
    const uint32_t key = cx.b_lo + blockIdx.x;
    const uint32_t lo = __ldg(off + key), hi = __ldg(off + key + 1);
    if (lo == hi) return;
    auto first_at_least = [&](uint32_t k) -> uint32_t {  // first sorted item whose key is >= k
        uint32_t a = 0, b = n_items;
        while (a < b) {
            const uint32_t mid = a + ((b - a) >> 1);
            if (__ldg(keys + mid) < k) a = mid + 1;
            else b = mid;
        }
        return a;
    };
    const uint32_t i_lo = first_at_least(key), i_hi = first_at_least(key + 1);
    for (uint32_t ib = i_lo; ib < i_hi; ib += ITEMS) {
        const uint32_t n_it = i_hi - ib < (uint32_t)ITEMS ? i_hi - ib : (uint32_t)ITEMS;
        __syncthreads();  // the previous batch has been written back
        if (tid < n_it) {
            const uint32_t id = __ldg(ids + ib + tid);
            const uint32_t pl = id / per_probe;
            // the item was active when the keys were made; its minimum may have dropped since
            s_cur[tid] = seed_item_setup(cx, q0 + pl, id - pl * per_probe, best, items[tid]) ? items[tid].cur : 0u;
            s_mine[tid] = s_cur[tid];
            if (cx.deep && hi - lo > cx.depth_cap && s_cur[tid]) cx.deep[items[tid].p] = 1;
        }
        if (occ && tid == 0) atomicAdd(occ + (blockIdx.x & (kSeedOccSlots - 1)), (unsigned long long)(hi - lo) * n_it);
        for (uint32_t t0 = lo; t0 < hi; t0 += TILE) {
            const uint32_t n = hi - t0 < (uint32_t)TILE ? hi - t0 : (uint32_t)TILE;
            __syncthreads();  // the items are set up / the previous tile has been consumed
            for (uint32_t e = tid; e < n; e += 256) {
                const uint4 v = __ldg(ent + t0 + e);  // one 16-byte entry: position + flank signature
                tile_pos[e] = v.x;
                tile_sig[e] = make_uint2(v.y, v.z);
            }
            __syncthreads();
            for (uint32_t j = warp; j < n_it; j += 8) {
                uint32_t mine = s_mine[j];
                if (mine == 0) continue;
                // the item stays in shared memory (the out-of-line verification takes it by reference: a per-thread
                // copy would be written to local memory for every item and tile); the hot fields go to registers
                const SeedItem &it = items[j];
                const uint32_t iq0 = it.q0, iq1 = it.q1, im = it.m;
                const uint2 *sig = tile_sig + lane;
                const uint32_t *tpos = tile_pos + lane;
                // lower bound of the distance from the flank signature (the item's q0 / q1 are masked at set-up)
                auto bound = [&](uint2 sg) -> uint32_t { return __popc(((sg.x ^ iq0) | (sg.y ^ iq1)) & im); };
                // 8 entries per lane and round: independent loads and POPCs in flight, one branch per round;
                // whole rounds run without bounds checks, only the tail of a tile tests e < n
                uint32_t e0 = 0;
                for (; e0 + 256 <= n && mine; e0 += 256) {
                    uint32_t lb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) lb[u] = bound(sig[e0 + u * 32]);
                    const uint32_t lo4 = min(min(lb[0], lb[1]), min(lb[2], lb[3]));
                    const uint32_t hi4 = min(min(lb[4], lb[5]), min(lb[6], lb[7]));
                    if (min(lo4, hi4) < mine) {  // rare
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (lb[u] < mine) seed_verify(cx, it, tpos[e0 + u * 32], mine);
                    }
                }
                for (; e0 < n && mine; e0 += 32)
                    if (e0 + lane < n && bound(sig[e0]) < mine) seed_verify(cx, it, tpos[e0], mine);
                mine = __reduce_min_sync(0xffffffffu, mine);
                if (lane == 0) s_mine[j] = mine;
            }
        }
        __syncthreads();
        if (tid < n_it && s_mine[tid] < s_cur[tid]) atomicMin(best + items[tid].p, s_mine[tid]);
    }
}

// probe K-mers flagged by the depth watch that were answered below the "not found" value: only there
// can a truncated search of the reference have missed the hit this engine reports
__global__ void __launch_bounds__(256) seed_deep_count_kernel(const uint8_t *__restrict__ deep,
                                                              const uint32_t *__restrict__ best, ImageView q, uint32_t n,
                                                              uint32_t clamp, unsigned long long *__restrict__ count) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = p < n && deep[p] && best[p] < clamp && ((q.valid()[p >> 5] >> (p & 31)) & 1u);
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}
cudaError_t launch_seed_deep_count(const uint8_t *d_deep, const uint32_t *d_best, ImageView q, uint32_t n, uint32_t clamp,
                                   unsigned long long *d_count, cudaStream_t st) {
    if (!n) return cudaSuccess;
    seed_deep_count_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_deep, d_best, q, n, clamp, d_count);
    return cudaGetLastError();
}

uint32_t seed_bucket_bits(uint32_t core_len) {
    return 2 * core_len < kSeedMaxBits ? 2 * core_len : kSeedMaxBits;
}

size_t seed_scan_temp_bytes(uint32_t n_buckets) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_buckets + 1);
    return bytes;
}

// partition passes.  FAST_A: three-word field extraction in pass A (a persistent pass A with the next tile's plane
// words prefetched was measured too: 0.2 ms slower, profiles/r02_index_build_variants_cfg4.jsonl); PERSIST_B: persistent pass B with the next tile's
// loads in flight during the copy-out (256 threads only)
template <int THREADS, bool FAST_A, bool PERSIST_B>
static cudaError_t part_launch(ImageView t, uint32_t n_pos, uint32_t core_len, uint32_t bits, uint32_t b_lo, uint32_t b_hi,
                               const uint32_t *d_off, uint32_t *d_cur_a, uint32_t *d_cursor, uint4 *d_part, uint4 *d_ent,
                               cudaStream_t st) {
    constexpr uint32_t kTile = THREADS * kPartPer;
    constexpr size_t kSmem = (size_t)kTile * sizeof(uint4);
    static_assert(!PERSIST_B || THREADS == 256, "the persistent pass B has 256 threads");
    if (kSmem > 48 * 1024) {  // opt-in, per device
        cudaError_t e = cudaFuncSetAttribute(seed_part_a_kernel<THREADS, FAST_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(seed_part_b_kernel<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
    }
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const uint32_t a_tiles = (n_pos + kTile - 1) / kTile;
    seed_part_a_kernel<THREADS, FAST_A><<<a_tiles, THREADS, kSmem, st>>>(t, n_pos, core_len, bits, b_lo, b_hi, d_off, d_cur_a, d_part);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // tiles of all coarse partitions: at most entries / tile + one ragged tile per partition
    const uint32_t max_tiles = n_pos / kTile + 257;
    if (!PERSIST_B) {
        seed_part_b_kernel<THREADS><<<max_tiles, THREADS, kSmem, st>>>(bits, d_off, d_cursor, d_part, d_ent);
    } else {
        const uint32_t grid = std::min<uint32_t>(max_tiles, (uint32_t)sms * 4u);
        seed_part_b_persistent_kernel<<<grid, 256, kSmem, st>>>(bits, d_off, d_cursor, d_part, d_ent);
    }
    return cudaGetLastError();
}

bool seed_index_can_partition(uint32_t core_len) {
    const uint32_t bits = seed_bucket_bits(core_len);
    return bits >= 9 && bits <= 16;
}

// d_cnt: n_buckets+1 counters (zeroed by the caller), turned into bucket offsets d_off
// (n_buckets+1 entries, d_off[n_buckets] = number of indexed cores); d_cursor: n_buckets+1
// scratch; d_ent: t.len entries.  Only cores whose bucket lies in [b_lo, b_hi) are indexed.
// d_part: nullptr = the entries are placed one by one (atomic cursor + scattered store), else t.len
// entries of scratch for the two partition passes (seed_index_can_partition(core_len) must hold), in tiles
// of 2048 entries; part_mode is a set of kSeedIndex* flags (k4b_kernels.cuh) that select the kernel variants.
cudaError_t launch_seed_index(ImageView t, uint32_t core_len, uint32_t b_lo, uint32_t b_hi, uint32_t *d_cnt,
                              uint32_t *d_off, uint32_t *d_cursor, uint4 *d_ent, uint4 *d_part, int part_mode, void *d_temp,
                              size_t temp_bytes, cudaStream_t st, int *n_launches) {
    if (n_launches) *n_launches = 0;
    if (t.len < core_len) return cudaSuccess;
    const uint32_t bits = seed_bucket_bits(core_len), nb = 1u << bits;
    const uint32_t n_pos = t.len - core_len + 1;
    const uint32_t grid = (n_pos + kSeedChunk - 1) / kSeedChunk;
    if (d_part && !seed_index_can_partition(core_len)) return cudaErrorInvalidValue;
    cudaError_t e;
    if (d_part && (part_mode & kSeedIndexSmemCount)) {
        int dev = 0, sms = 0;
        const size_t smem = (size_t)nb * 2;  // 16 bits per bucket
        e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(seed_count_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const uint32_t n_chunks = (n_pos + kCountChunk - 1) / kCountChunk;
        seed_count_smem_kernel<<<std::min<uint32_t>(n_chunks, (uint32_t)sms), kCountThreads, smem, st>>>(t, n_pos, core_len, bits,
                                                                                                       b_lo, b_hi, d_cnt);
    } else {
        seed_scan_kernel<false><<<grid, 256, 0, st>>>(t, n_pos, core_len, bits, b_lo, b_hi, d_cnt, nullptr);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_cnt, d_off, (int)nb + 1, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d_cursor, d_off, ((size_t)nb + 1) * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    if (!d_part) {
        seed_scan_kernel<true><<<grid, 256, 0, st>>>(t, n_pos, core_len, bits, b_lo, b_hi, d_cursor, d_ent);
        if (n_launches) *n_launches = 2;
        return cudaGetLastError();
    }
    e = cudaMemsetAsync(d_cnt, 0, 256 * 4, st);  // the counts are spent: 256 coarse cursors for pass A
    if (e != cudaSuccess) return e;
    if (n_launches) *n_launches = 3;
#define K4B_PART(T, FA, PB) part_launch<T, FA, PB>(t, n_pos, core_len, bits, b_lo, b_hi, d_off, d_cnt, d_cursor, d_part, d_ent, st)
    const bool fast_a = (part_mode & kSeedIndexFastA) != 0, persist_b = (part_mode & kSeedIndexPersistB) != 0;
    if (part_mode & kSeedIndexTile4096) return fast_a ? K4B_PART(512, true, false) : K4B_PART(512, false, false);
    if (persist_b) return fast_a ? K4B_PART(256, true, true) : K4B_PART(256, false, true);
    return fast_a ? K4B_PART(256, true, false) : K4B_PART(256, false, false);
#undef K4B_PART
}

// kernel shapes of the join: {items per batch, entries per tile}; K4B_SEED_JOIN_VARIANT picks one (tests, measurements).
// Also measured and dropped (profiles/r02_join_variants_cfg4_5ctas.jsonl): 5 CTAs per SM (register cap 51: 55 ms against
// 41 ms for config 4) and batches of 256 items (45.4 ms).
constexpr int kJoinDefaultVariant = 0;
static cudaError_t join_launch(int variant, const SeedCtx &cx, const uint32_t *d_off, const uint4 *d_ent, uint32_t q0,
                               uint32_t n_items, const uint32_t *keys, const uint32_t *ids, uint32_t *d_best,
                               unsigned long long *d_occ, cudaStream_t st) {
    const uint32_t grid = cx.b_hi - cx.b_lo;  // one CTA per bucket of this launch's share
    switch (variant) {
        case 0: seed_join_kernel<128, 3072><<<grid, 256, 0, st>>>(cx, d_off, d_ent, q0, n_items, keys, ids, d_best, d_occ); break;
        case 1: seed_join_kernel<64, 2048><<<grid, 256, 0, st>>>(cx, d_off, d_ent, q0, n_items, keys, ids, d_best, d_occ); break;
        case 2: seed_join_kernel<128, 1024><<<grid, 256, 0, st>>>(cx, d_off, d_ent, q0, n_items, keys, ids, d_best, d_occ); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_seed_query(ImageView q, ImageView rcq, ImageView t, uint32_t K, uint32_t core_len,
                              const uint32_t *d_off, const uint4 *d_ent, uint32_t q_begin, uint32_t q_end,
                              uint32_t b_lo, uint32_t b_hi, uint32_t clamp, bool crick, bool three, bool q_impure,
                              SeedSelfRules self, uint32_t depth_cap, uint8_t *d_deep, uint32_t *d_best,
                              unsigned long long *d_occ, cudaStream_t st, int *n_launches) {
    int own = 0;  // kernels of this file enqueued (the library sort launches are not counted)
    struct Report {
        int *dst, &n;
        ~Report() { if (dst) *dst = n; }
    } report{n_launches, own};
    if (q_begin >= q_end || t.len < K || q.len < K) return cudaSuccess;
    SeedCtx cx;
    cx.q = q;
    cx.rcq = rcq;
    cx.t = t;
    cx.K = K;
    cx.core_len = core_len;
    cx.n_cores = K / core_len;
    cx.bits = seed_bucket_bits(core_len);
    cx.clamp = clamp;
    cx.b_lo = b_lo;
    cx.b_hi = b_hi;
    cx.depth_cap = depth_cap;
    cx.deep = d_deep;
    cx.strands = crick ? 2 : 1;
    cx.three = three ? 1 : 0;
    cx.q_impure = q_impure ? 1 : 0;
    cx.core_only = -1;
    cx.done_below = 0;
    cx.self = self;
    const uint32_t per_probe = (uint32_t)cx.strands * cx.n_cores;
    // bucket-major join when buckets are long enough to be worth staging (short cores on long
    // targets); otherwise one warp per item streams its (short) bucket
    const char *env = getenv("K4B_SEED_JOIN");
    const bool join = env ? atoi(env) != 0 : (t.len >> cx.bits) >= 256;
    if (!join) {
        // 1-D grids hold 2^31-1 CTAs of 8 warps: split the probe range if needed
        const unsigned long long max_warps = 0x7fffffffull * 8ull;
        const uint32_t max_probes = (uint32_t)std::min<unsigned long long>(0xffffffffull, max_warps / per_probe);
        for (uint32_t b = q_begin; b < q_end;) {
            const uint32_t e = (q_end - b > max_probes) ? b + max_probes : q_end;
            const unsigned long long w = (unsigned long long)(e - b) * per_probe;
            seed_query_kernel<<<(unsigned)((w + 7) / 8), 256, 0, st>>>(cx, d_off, d_ent, b, e, d_best, d_occ);
            ++own;
            const cudaError_t err = cudaGetLastError();
            if (err != cudaSuccess) return err;
            b = e;
        }
        return cudaSuccess;
    }
    // probe chunks of at most 2^25 items: keys -> radix sort by bucket -> join, core by core when this
    // device holds the whole index (phase schedule, see above; K4B_SEED_PHASES=0 joins all cores in one pass)
    const char *ph = getenv("K4B_SEED_PHASES");
    const bool phases = (ph ? atoi(ph) != 0 : true) && cx.n_cores > 1 && b_lo == 0 && b_hi == (1u << cx.bits);
    const char *jv = getenv("K4B_SEED_JOIN_VARIANT");  // kernel shape, for measurements (see join_launch)
    const int variant = jv ? atoi(jv) : kJoinDefaultVariant;
    const uint32_t chunk_probes = std::max(1u, (1u << 25) / per_probe);
    const uint32_t max_items = std::min<unsigned long long>((unsigned long long)chunk_probes,
                                                            (unsigned long long)(q_end - q_begin)) * per_probe;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)max_items, 0, (int)cx.bits + 1, st);
    uint32_t *d_buf = nullptr;
    void *d_sort = nullptr;
    cudaError_t err = cudaMallocAsync(&d_buf, (size_t)max_items * 16, st);
    if (err == cudaSuccess) err = cudaMallocAsync(&d_sort, sort_bytes ? sort_bytes : 4, st);
    if (err != cudaSuccess) {
        if (d_buf) cudaFreeAsync(d_buf, st);
        return err;
    }
    uint32_t *k_in = d_buf, *k_out = d_buf + max_items, *i_in = k_out + max_items, *i_out = i_in + max_items;
    for (uint32_t b = q_begin; b < q_end && err == cudaSuccess;) {
        const uint32_t e = (q_end - b > chunk_probes) ? b + chunk_probes : q_end;
        for (uint32_t pass = 0; pass < (phases ? cx.n_cores : 1u) && err == cudaSuccess; ++pass) {
            cx.core_only = phases ? (int)pass : -1;
            cx.done_below = phases ? pass : 0u;
            const uint32_t n_items = (e - b) * (phases ? (uint32_t)cx.strands : per_probe);
            seed_item_keys_kernel<<<(n_items + 255) / 256, 256, 0, st>>>(cx, b, e - b, d_best, d_off, k_in, i_in);
            err = cudaGetLastError();
            if (err == cudaSuccess)
                err = cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, k_in, k_out, i_in, i_out, (int)n_items, 0,
                                                      (int)cx.bits + 1, st);
            if (err == cudaSuccess) err = join_launch(variant, cx, d_off, d_ent, b, n_items, k_out, i_out, d_best, d_occ, st);
            own += 2;  // item keys + join
        }
        b = e;
    }
    cudaFreeAsync(d_sort, st);
    cudaFreeAsync(d_buf, st);
    return err;
}

}  // namespace k4b
