// k4b_kernels.cuh - sm_100a kernels of the K-mer Hamming-distance engine.
//
// Replaces the reference's sliding-diagonal CPU loops (ngskit4b/hammings.cpp:3183-3287
// GHamDistWatson, :3300-3489 GHamDistCrick) with a query-centric all-pairs scan:
//   - sequences live in HBM as bit-planes (plane b holds bit b of every base code, 1 bit per
//     base, little-endian bit order inside 32-bit words) plus a valid-K-mer-start plane;
//   - a thread keeps Q query K-mers x {forward, reverse-complement} in registers;
//   - all 32 lanes of a warp walk the SAME candidate stream, staged tile by tile into shared
//     memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier), double buffered;
//   - a candidate K-mer is cut out of the planes with one funnel shift per plane word;
//   - one word-compare is (q0^c0) | (q1^c1) [| (q2^c2)] -> POPC, summed over W words, then a
//     running per-query minimum in a register.
// Integer-pipe work only: no tensor cores (deliberately).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace k4b {

constexpr int kThreads = 256;        // threads per CTA
constexpr int kTileGroups = 256;     // 32-candidate groups per smem tile (8192 candidates)
constexpr int kTilePad = 8;          // extra plane words staged so funnel shifts can look ahead
constexpr int kTileWords = kTileGroups + kTilePad;
constexpr int kMaxRegW = 4;          // W (words per K-mer) handled by the register-resident path
constexpr uint32_t kNoDist = 0xffffu;

constexpr int kFrontPadWords = 320;  // EOS-filled words before logical position 0 (diagonal engine
                                     // reads columns at slightly negative offsets)
constexpr int kBackPadWords = 576;   // EOS-filled words after the last sequence word

// Layout of one packed image: 4 arrays (plane0, plane1, plane2, valid) of `stride` words each,
// logical word 0 of an array sits kFrontPadWords into it; then the 2-bit CODE array of 2 * stride
// words (bits 2i, 2i+1 of its logical word i/16 = plane-0 / plane-1 bit of base i; 16 bases per
// word, logical word 0 sits 2 * kFrontPadWords into it): the band engine reads the base codes of
// 32 consecutive rows as one 64-bit window of it.  kImageArrays words of stride in total.
constexpr int kImageArrays = 6;
struct ImageView {
    const uint32_t *base;  // device pointer to logical word 0 of plane 0
    uint32_t stride;       // words between consecutive arrays
    uint32_t nwl;          // words addressable from logical word 0 (sequence + back pad)
    uint32_t len;          // bases (incl. EOS separators)
    __host__ __device__ const uint32_t *plane(int p) const { return base + (size_t)p * stride; }
    __host__ __device__ const uint32_t *valid() const { return base + (size_t)3 * stride; }
    __host__ __device__ const uint32_t *code2() const { return base + (size_t)4 * stride + kFrontPadWords; }
};

struct AllPairsParams {
    ImageView q;             // query set
    ImageView t;             // target set
    uint32_t K;
    uint32_t q_begin, q_end; // flat query start positions [q_begin, q_end)
    uint32_t tiles_total;    // number of kTileGroups tiles covering the target set
    uint32_t tiles_per_chunk;
    uint32_t groups_limit;   // generic (K > 128) kernel only: stop after this many 32-candidate groups (0 = all)
    uint32_t *out;           // [q_end-q_begin] running minima (pre-set to K+1)
    int self_exclude;        // skip forward pair with query pos == target pos
    int wildcard;            // targeted mode: query symbols >= N match any target ACGT base
    // sweep sub-range (-b/-B, -m2 slices; hammings.cpp:924-928): a forward pair (q,t) counts iff
    // w_lo <= |q-t| <= w_hi, a reverse-complement pair iff c = q+t+K-1 lies in [c1_lo,c1_hi] or
    // [c2_lo,c2_hi] (the two anti-diagonals one Crick sweep offset covers)
    int ranged;
    long long w_lo, w_hi, c1_lo, c1_hi, c2_lo, c2_hi;
    // -z intra/inter filter (probes drawn from the indexed assembly; SfxArray.cpp:4421-4426,
    // :4597-4601): an EXACT forward hit counts only if it lies in the same (1) / another (2)
    // entry than the query.  ent_starts: ascending flat start positions of the n_ent entries.
    int zfilt;
    const uint32_t *ent_starts;
    uint32_t n_ent;
};

// diagonal-band engine (k4b_diag.cu)
enum DiagMode : int {
    kDiagWatson = 0,  // A = B = sequence, diagonals s >= 1, every cell lowers both K-mers
    kDiagCrick = 1,   // B = reverse complement of A; first half of each mirror-symmetric diagonal
    kDiagRect = 2,    // A = probes (or their reverse complement), B = targets; rows only
};
struct DiagParams {
    ImageView a;             // row sequence planes
    ImageView b;             // column sequence planes
    ImageView va, vb;        // images whose valid-start planes validate row / column POSITIONS
    uint32_t K;
    uint32_t Mrow, Mcol;     // last K-mer start of the row / column sequence (len - K)
    int mode;                // DiagMode
    int row_flip, col_flip;  // position of a row (column) index r is Mrow - r (Mcol - c): the
                             // sequence in a (b) is the reverse complement of the one indexed by best[]
    int update_cols;         // also lower best[] at the column position (symmetric modes)
    int wild;                // row symbols >= N are wildcards (targeted rules, needs 3 planes)
    uint32_t t_fixed;        // kDiagRect: fixed threshold (the clamp); else thresholds come from blockmax
    long long s_first;       // diagonal (column - row) of lane 0 of group 0
    // CTA groups of this launch: g = part + nparts * q for the q with (q mod q_period) in
    // [q_lo, q_lo + q_span), enumerated from index l_first (parts of the pair matrix x slabs)
    uint32_t part, nparts, q_period, q_lo, q_span, l_first;
    uint32_t n_seg, rows_per_seg;  // row segments per group
    uint32_t *best;          // running minima per position (atomicMin)
    const uint32_t *blockmax;  // max of best over 2^bm_shift positions (valid starts only)
    uint32_t bm_shift;
    // kernel-variant selection without a host round trip: the same launch is enqueued with two
    // counter widths and each instance checks the global maximum threshold on the device
    const uint32_t *tmax_ptr;  // max over all blockmax entries (written by blockmax_kernel)
    uint32_t sel_limit;
    const uint32_t *low_ptr;   // number of blocks whose maximum lies below the narrow instance's floor
    uint32_t low_max;          // (their thresholds would be lifted: more flags); narrow only if <= low_max
    int sel;                   // 0 run always, 1 run iff narrow fits (*tmax_ptr <= sel_limit and few
                               // low blocks), 2 iff not
};
constexpr int kDiagGroupDiagonals = 8 * 1024;  // diagonals per CTA group (8 warps x 1024)
// d_tmax (nullable): receives the maximum over all blocks; d_n_low (nullable): the number of
// blocks with 0 < maximum < low_floor; both must be zeroed by the caller
cudaError_t launch_blockmax(const uint32_t *d_best, ImageView a, uint32_t n_pos, uint32_t shift,
                            uint32_t *d_blockmax, uint32_t n_blocks, uint32_t *d_tmax, uint32_t low_floor,
                            uint32_t *d_n_low, cudaStream_t st);
// np = number of counter planes (5..14); thresholds must fit 2^(np-1)
cudaError_t launch_diag(const DiagParams &p, bool three_planes, int np, uint32_t n_groups, cudaStream_t st,
                        unsigned long long *n_ctas);
int diag_planes_for_k(uint32_t K);

// host-callable launchers (defined in k4b_kernels.cu)
// img.base must point into a writable allocation (the launchers cast constness away)
cudaError_t launch_pack(const uint8_t *d_concat, ImageView img, uint32_t *d_flags, cudaStream_t st);
cudaError_t launch_valid(ImageView img, uint32_t K, unsigned long long *d_count, cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t *d, uint32_t n, uint32_t v, cudaStream_t st);
cudaError_t launch_finalize(const uint32_t *d_min32, ImageView q, uint32_t q_begin, uint32_t n,
                            uint32_t K, uint32_t clamp, int max_wild, uint16_t *d_out16,
                            cudaStream_t st);
// d_hist: nbins counters, zeroed by the caller; values >= nbins land in the last bin
cudaError_t launch_histogram_u16(const uint16_t *d_v, uint32_t n, uint32_t nbins, unsigned long long *d_hist,
                                 cudaStream_t st);
// returns cudaErrorInvalidValue when K needs more than the supported words
cudaError_t launch_allpairs(const AllPairsParams &p, bool three_planes, bool crick,
                            cudaStream_t st, int *n_ctas);
// K > 32*kMaxRegW: rc = 3 reverse-complemented planes of the query set (same geometry as q)
cudaError_t launch_allpairs_generic(const AllPairsParams &p, bool three_planes, bool crick,
                                    ImageView rc, cudaStream_t st, int *n_ctas);
// rc.base: logical word 0 of a 3-array allocation with q's stride and pads
cudaError_t launch_revcomp_planes(ImageView q, ImageView rc, cudaStream_t st);
int queries_per_thread(uint32_t W, bool three_planes);

// seed-and-verify engine for the targeted mode (k4b_seed.cu)
struct SeedSelfRules {          // probes drawn from the indexed assembly itself (no -I)
    int on;                     // skip the exact sense-strand hit at the probe's own position
    int zfilt;                  // -z: 1 intra only, 2 inter only (exact sense hits only)
    const uint32_t *ent_starts; // ascending flat start positions of the entries
    uint32_t n_ent;
};
uint32_t seed_bucket_bits(uint32_t core_len);
size_t seed_scan_temp_bytes(uint32_t n_buckets);
// only cores whose bucket lies in [b_lo, b_hi) are indexed / answered (bucket shards of a multi-GPU run)
// d_part: nullptr = entries placed one by one; else t.len entries of scratch for the two partition passes
// (only when seed_index_can_partition(core_len)); *n_launches = kernels of k4b_seed.cu enqueued
bool seed_index_can_partition(uint32_t core_len);
cudaError_t launch_seed_index(ImageView t, uint32_t core_len, uint32_t b_lo, uint32_t b_hi, uint32_t *d_cnt,
                              uint32_t *d_off, uint32_t *d_cursor, uint4 *d_ent, uint4 *d_part, int part_mode, void *d_temp,
                              size_t temp_bytes, cudaStream_t st, int *n_launches);
cudaError_t launch_seed_query(ImageView q, ImageView rcq, ImageView t, uint32_t K, uint32_t core_len,
                              const uint32_t *d_off, const uint4 *d_ent, uint32_t q_begin, uint32_t q_end,
                              uint32_t b_lo, uint32_t b_hi, uint32_t clamp, bool crick, bool three, bool q_impure,
                              SeedSelfRules self, uint32_t depth_cap, uint8_t *d_deep, uint32_t *d_best,
                              unsigned long long *d_occ, cudaStream_t st, int *n_launches);
// d_deep (nullable, one byte per probe position, zeroed by the caller): set to 1 where a core of the probe
// K-mer has more than depth_cap index entries; launch_seed_deep_count adds to *d_count the flagged valid
// probe K-mers whose minimum lies below clamp
cudaError_t launch_seed_deep_count(const uint8_t *d_deep, const uint32_t *d_best, ImageView q, uint32_t n, uint32_t clamp,
                                   unsigned long long *d_count, cudaStream_t st);
// index build, K4B_SEED_INDEX overrides: 0 = entries placed one by one; else partition passes (flag 1) with
// tiles of 4096 instead of 2048 entries (2), the counters of the count pass in shared memory (4), the three-word
// field extraction in pass A (8), the persistent pass B (16; tiles of 2048 only)
constexpr int kSeedIndexPartition = 1, kSeedIndexTile4096 = 2, kSeedIndexSmemCount = 4, kSeedIndexFastA = 8,
              kSeedIndexPersistB = 16;
constexpr int kSeedIndexDefault = kSeedIndexPartition | kSeedIndexSmemCount | kSeedIndexFastA | kSeedIndexPersistB;
constexpr int kSeedOccSlots = 1024;  // d_occ (nullable): bucket entries streamed, summed over these slots
cudaError_t launch_microbench(int which, int iters, uint32_t *d_sink, int *blocks, int *threads,
                              int *ops_per_thread_iter, cudaStream_t st);

}  // namespace k4b
