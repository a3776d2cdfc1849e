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

// Layout of one packed image: 4 arrays of nwp words: plane0, plane1, plane2, valid.
struct ImageView {
    const uint32_t *base;  // device pointer
    uint32_t nwp;          // padded words per array
    uint32_t len;          // bases (incl. EOS separators)
    __host__ __device__ const uint32_t *plane(int p) const { return base + (size_t)p * nwp; }
    __host__ __device__ const uint32_t *valid() const { return base + (size_t)3 * nwp; }
};

struct AllPairsParams {
    ImageView q;             // query set
    ImageView t;             // target set
    uint32_t K;
    uint32_t q_begin, q_end; // flat query start positions [q_begin, q_end)
    uint32_t tiles_total;    // number of kTileGroups tiles covering the target set
    uint32_t tiles_per_chunk;
    uint32_t *out;           // [q_end-q_begin] running minima (pre-set to K+1)
    int self_exclude;        // skip forward pair with query pos == target pos
    int wildcard;            // targeted mode: query symbols >= N match any target ACGT base
    // sweep sub-range (-b/-B, -m2 slices; hammings.cpp:924-928): a forward pair (q,t) counts iff
    // w_lo <= |q-t| <= w_hi, a reverse-complement pair iff c = q+t+K-1 lies in [c1_lo,c1_hi] or
    // [c2_lo,c2_hi] (the two anti-diagonals one Crick sweep offset covers)
    int ranged;
    long long w_lo, w_hi, c1_lo, c1_hi, c2_lo, c2_hi;
};

// host-callable launchers (defined in k4b_kernels.cu)
cudaError_t launch_pack(const uint8_t *d_concat, uint32_t len, uint32_t *d_image, uint32_t nwp,
                        uint32_t *d_flags, cudaStream_t st);
cudaError_t launch_valid(uint32_t *d_image, uint32_t nwp, uint32_t len, uint32_t K,
                         unsigned long long *d_count, cudaStream_t st);
cudaError_t launch_fill_u32(uint32_t *d, uint32_t n, uint32_t v, cudaStream_t st);
cudaError_t launch_finalize(const uint32_t *d_min32, ImageView q, uint32_t q_begin, uint32_t n,
                            uint32_t K, uint32_t clamp, int max_wild, uint16_t *d_out16,
                            cudaStream_t st);
// returns cudaErrorInvalidValue when K needs more than the supported words
cudaError_t launch_allpairs(const AllPairsParams &p, bool three_planes, bool crick,
                            cudaStream_t st, int *n_ctas);
// K > 32*kMaxRegW: d_q_rc_planes = 3 reverse-complemented planes of the query set (nwp words each)
cudaError_t launch_allpairs_generic(const AllPairsParams &p, bool three_planes, bool crick,
                                    const uint32_t *d_q_rc_planes, uint32_t q_rc_nwp,
                                    cudaStream_t st, int *n_ctas);
cudaError_t launch_revcomp_planes(ImageView q, uint32_t *d_rc_planes, cudaStream_t st);
int queries_per_thread(uint32_t W, bool three_planes);
cudaError_t launch_microbench(int which, int iters, uint32_t *d_sink, int *blocks, int *threads,
                              int *ops_per_thread_iter, cudaStream_t st);

}  // namespace k4b
