// k4b_kernels.cu - sm_100a kernels: bit-plane pack, valid-start plane, all-pairs minimum
// Hamming distance, finalize, integer-pipe microbenchmarks.  See k4b_kernels.cuh.
#include "k4b_kernels.cuh"

namespace k4b {

// ---------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS / UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------
// pack: 1 byte/base concat (codes 0..7) -> three bit-planes.  Streaming, HBM-bound:
// reads len bytes, writes 3*len/8 bytes.  Positions >= len are written as EOS (7) so the
// padding can never be part of a valid K-mer.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gather4(uint32_t x, int p) {
    // bit p of each of the 4 bytes of x -> 4-bit nibble (byte 0 -> bit 0)
    uint32_t t = (x >> p) & 0x01010101u;
    return (t * 0x10204080u) >> 28;
}

// 16 bits -> 32 bits with a zero between neighbours (bit i -> bit 2i)
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    return (x | (x << 1)) & 0x55555555u;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ concat, ImageView img,
                                                   uint32_t *__restrict__ image, uint32_t *__restrict__ flags) {
    // thread i packs 16 bases into 16 bits per plane; lane pairs merge to a word.  Array words
    // run from -kFrontPadWords to nwl-1 (logical), i.e. `image` points at the first pad word.
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w_abs = i >> 1;
    const long long b0 = ((long long)w_abs - kFrontPadWords) * 32 + (long long)(i & 1) * 16;
    const long long len = img.len;
    uint32_t x[4];
    if (b0 >= 0 && b0 + 16 <= len) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(concat + b0));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long pos = b0 + 4 * k + j;
                const uint32_t c = (pos >= 0 && pos < len) ? (uint32_t)concat[pos] : 7u;
                w |= c << (8 * j);
            }
            x[k] = w;
        }
    }
    uint32_t frag[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        frag[p] = gather4(x[0], p) | (gather4(x[1], p) << 4) | (gather4(x[2], p) << 8) |
                  (gather4(x[3], p) << 12);
    }
    // codes 4..6 (N, Undef, InDel): bit 2 set but not all three bits
    const uint32_t non_acgt = frag[2] & ~(frag[0] & frag[1]);
    if (__any_sync(0xffffffffu, non_acgt != 0) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
    // 2-bit code array: this thread's 16 bases are exactly one of its words
    if (w_abs < img.stride) image[(size_t)4 * img.stride + i] = spread16(frag[0]) | (spread16(frag[1]) << 1);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint32_t hi = __shfl_down_sync(0xffffffffu, frag[p], 1);
        if ((threadIdx.x & 1) == 0 && w_abs < img.stride)
            image[(size_t)p * img.stride + w_abs] = frag[p] | (hi << 16);
    }
}

// ---------------------------------------------------------------------------------------
// valid plane: bit i set iff the window [i, i+K) holds no EOS (code 7 = all plane bits set)
// i.e. the K-mer lies inside one chromosome (hammings.cpp:3084-3094 NumSubSeqs, :3260 EOS
// counter).  Also counts the valid K-mers.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t eos_word(const ImageView &img, uint32_t w) {
    // branch-free (the loads of several words can be in flight together): words past the end are
    // clamped onto the last word of the back pad, which is EOS-filled like everything after the sequence
    w = w < img.nwl ? w : img.nwl - 1;
    return __ldg(img.plane(0) + w) & __ldg(img.plane(1) + w) & __ldg(img.plane(2) + w);
}
// first EOS position >= pos found while scanning the words that overlap [pos, pos+K); kNoEos
// when those words hold none (a hit may lie at or beyond pos+K: callers compare)
constexpr uint64_t kNoEos = ~0ull;
__device__ uint64_t next_eos(const ImageView &img, uint64_t pos, uint32_t K) {
    const uint64_t lim = pos + K;
    uint32_t w = (uint32_t)(pos >> 5);
    uint32_t e = eos_word(img, w) & (0xffffffffu << (pos & 31));
    while (true) {
        if (e) return ((uint64_t)w << 5) + (__ffs(e) - 1);
        ++w;
        if (((uint64_t)w << 5) >= lim) return kNoEos;
        e = eos_word(img, w);
    }
}

__global__ void __launch_bounds__(256) valid_kernel(ImageView img, uint32_t *__restrict__ valid_arr,
                                                    uint32_t K, unsigned long long *__restrict__ count) {
    // thread per array word; valid_arr points at the first (front pad) word of the valid array
    const uint32_t w_abs = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    if (w_abs < img.stride) {
        if (w_abs >= (uint32_t)kFrontPadWords) {
            const uint32_t g = w_abs - kFrontPadWords;
            const uint64_t base = (uint64_t)g << 5;
            if (base < img.len) {
                // common case: no EOS anywhere in [base, base+31+K) -> all 32 starts are valid.  For K <= 225
                // that window is at most 8 words: their loads are issued together (no dependent scan)
                const uint32_t span = K + 31, nwin = (span + 31) / 32;
                bool clean = false;
                uint64_t ne = kNoEos;
                if (nwin <= 8) {
                    uint32_t any = 0;
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        if (k < nwin) {
                            uint32_t e = eos_word(img, g + k);
                            const uint32_t left = span - 32 * k;  // bits of this word inside the window
                            if (left < 32) e &= (1u << left) - 1u;
                            any |= e;
                        }
                    }
                    clean = any == 0;
                    if (!clean) ne = next_eos(img, base, span);
                } else {
                    ne = next_eos(img, base, span);
                    clean = ne == kNoEos || ne >= base + span;
                }
                if (clean) {
                    v = 0xffffffffu;
                } else {
                    for (int i = 0; i < 32; ++i) {
                        const uint64_t pos = base + i;
                        if (ne == kNoEos || ne < pos) ne = next_eos(img, pos, K);
                        if (ne == kNoEos || ne >= pos + K) v |= 1u << i;
                    }
                }
            }
        }
        valid_arr[w_abs] = v;
    }
    // number of valid K-mers: one global atomic per CTA (one per warp made 0.5 M atomics on a single address
    // the bottleneck of this kernel at 500 Mbp)
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const uint32_t c = __reduce_add_sync(0xffffffffu, __popc(v));
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0 && s_count) atomicAdd(count, (unsigned long long)s_count);
}

__global__ void fill_u32_kernel(uint32_t *d, uint32_t n, uint32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}

// minima -> uint16 results; positions that are not valid K-mer starts report K+1 (the
// reference's "never lowered" fill value, hammings.cpp:3120-3122)
// max_wild >= 0 (targeted mode): a query K-mer holding more than max_wild symbols >= N reports 0
// (SfxArray.cpp:4322-4326)
__global__ void finalize_kernel(const uint32_t *__restrict__ min32, ImageView q, uint32_t q_begin,
                                uint32_t n, uint32_t K, uint32_t clamp, int max_wild,
                                uint16_t *__restrict__ out16) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pos = q_begin + i;
    const bool ok = pos < q.len && ((q.valid()[pos >> 5] >> (pos & 31)) & 1u);
    uint32_t v = min32[i];
    if (clamp && v > clamp) v = clamp;
    if (ok && max_wild >= 0) {
        const uint32_t *p2 = q.plane(2);
        uint32_t cnt = 0;
        for (uint32_t done = 0; done < K; done += 32) {
            const uint32_t at = pos + done, wi = at >> 5, sh = at & 31;
            uint32_t bits = __funnelshift_r(p2[wi], p2[wi + 1], sh);
            if (K - done < 32) bits &= (1u << (K - done)) - 1u;
            cnt += __popc(bits);
        }
        if ((int)cnt > max_wild) v = 0;
    }
    out16[i] = ok ? (uint16_t)v : (uint16_t)(K + 1);
}

// Distribution of the minima (the table `hammings` logs, hammings.cpp:2939-2962, and the
// downstream HammingDist tool builds from the CSV): bin d counts the positions whose minimum is d,
// bin K+1 collects the positions where no K-mer starts.  Shared-memory privatised histogram,
// HBM-bound: reads 2 B per position.
__global__ void __launch_bounds__(256) histogram_u16_kernel(const uint16_t *__restrict__ v, uint32_t n, uint32_t nbins,
                                                            unsigned long long *__restrict__ hist) {
    extern __shared__ uint32_t sh_bins[];
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) sh_bins[b] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t d = v[i];
        atomicAdd(&sh_bins[d < nbins ? d : nbins - 1], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x)
        if (sh_bins[b]) atomicAdd(hist + b, (unsigned long long)sh_bins[b]);
}

// ---------------------------------------------------------------------------------------
// all-pairs minimum
// ---------------------------------------------------------------------------------------
template <int W, int P>
struct Kmer {
    uint32_t w[P][W];
};

// cut the K-mer starting at bit position pos out of the planes (global memory)
template <int W, int P>
__device__ __forceinline__ void load_kmer(const ImageView &img, uint32_t pos, uint32_t tail_mask,
                                          Kmer<W, P> &k) {
    const uint32_t wi = pos >> 5, sh = pos & 31;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t *pl = img.plane(p) + wi;
        uint32_t lo = __ldg(pl);
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const uint32_t hi = __ldg(pl + w + 1);
            k.w[p][w] = __funnelshift_r(lo, hi, sh);
            lo = hi;
        }
        k.w[p][W - 1] &= tail_mask;
    }
}

// reverse complement in plane form.  rc[p] = cpl(q[K-1-p]); cpl flips the two low code bits
// of A,C,G,T and leaves codes >= 4 unchanged (MapCpl, hammings.cpp:3173-3180).
template <int W, int P>
__device__ __forceinline__ void revcomp(const Kmer<W, P> &q, uint32_t K, uint32_t tail_mask,
                                        Kmer<W, P> &r) {
    const uint32_t sh = 32u * W - K;  // 0..31
#pragma unroll
    for (int p = 0; p < P; ++p) {
        uint32_t f[W + 1];
#pragma unroll
        for (int w = 0; w < W; ++w) f[w] = __brev(q.w[p][W - 1 - w]);
        f[W] = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) r.w[p][w] = __funnelshift_r(f[w], f[w + 1], sh);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t m = (w == W - 1) ? tail_mask : 0xffffffffu;
        if (P == 3) {
            const uint32_t flip = ~r.w[P - 1][w] & m;
            r.w[0][w] ^= flip;
            r.w[1][w] ^= flip;
        } else {
            r.w[0][w] = ~r.w[0][w] & m;
            r.w[1][w] = ~r.w[1][w] & m;
        }
    }
}

// a = query, b = candidate.  WILD (targeted mode, three planes): query symbols >= N are
// wildcards that match any target ACGT base but never a target N (SfxArray.cpp:4266-4296);
// the query was prepared by make_wild(): plane 2 holds the NOT-wild mask, planes 0/1 are zero
// at wildcard positions.
template <int W, int P, bool WILD>
__device__ __forceinline__ uint32_t kmer_dist(const Kmer<W, P> &a, const Kmer<W, P> &b) {
    uint32_t d = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        uint32_t m = a.w[0][w] ^ b.w[0][w];
        m |= a.w[1][w] ^ b.w[1][w];
        if (P == 3) {
            if (WILD) m = (m & a.w[2][w]) | b.w[2][w];
            else m |= a.w[2][w] ^ b.w[2][w];
        }
        d += __popc(m);
    }
    return d;
}

template <int W, int P>
__device__ __forceinline__ void make_wild(Kmer<W, P> &q, uint32_t tail_mask) {
    if (P != 3) return;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t nw = ~q.w[P - 1][w] & ((w == W - 1) ? tail_mask : 0xffffffffu);
        q.w[0][w] &= nw;
        q.w[1][w] &= nw;
        q.w[P - 1][w] = nw;
    }
}

// candidate K-mer at shift s of the group whose plane words are cw[p][0..W]
template <int W, int P>
__device__ __forceinline__ void cut_candidate(const uint32_t (&cw)[P][W + 1], uint32_t s,
                                              uint32_t tail_mask, Kmer<W, P> &c) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
        for (int w = 0; w < W; ++w) c.w[p][w] = __funnelshift_r(cw[p][w], cw[p][w + 1], s);
        c.w[p][W - 1] &= tail_mask;
    }
}

// -z: does the exact forward hit of the K-mer at qpos on the K-mer at cpos count?
__device__ __noinline__ bool zfilter_pass(const AllPairsParams &prm, uint32_t qpos, uint32_t cpos) {
    auto entry_of = [&](uint32_t pos) {
        uint32_t lo = 0, hi = prm.n_ent;  // largest i with ent_starts[i] <= pos
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(prm.ent_starts + mid) <= pos) lo = mid;
            else hi = mid;
        }
        return lo;
    };
    const bool same = entry_of(qpos) == entry_of(cpos);
    return prm.zfilt == 1 ? same : !same;
}

template <int W, int P, int Q, bool CRICK, bool WILD, bool RANGED>
__global__ void __launch_bounds__(kThreads, 2) allpairs_min_kernel(const AllPairsParams prm) {
    constexpr int S = CRICK ? 2 : 1;
    constexpr int NARR = P + 1;  // planes + valid
    __shared__ __align__(128) uint32_t tile[2][NARR][kTileWords];
    __shared__ __align__(8) uint64_t full_bar[2];

    const uint32_t tid = threadIdx.x;
    const uint32_t K = prm.K;
    const uint32_t tail_mask = (K & 31) ? ((1u << (K & 31)) - 1u) : 0xffffffffu;
    const uint32_t qbase = prm.q_begin + blockIdx.y * (uint32_t)(kThreads * Q);

    // ---- queries into registers ----
    Kmer<W, P> qk[Q][S];
    uint32_t qpos[Q];
    uint32_t best[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        qpos[j] = qbase + j * kThreads + tid;
        load_kmer<W, P>(prm.q, qpos[j], tail_mask, qk[j][0]);
        if (CRICK) revcomp<W, P>(qk[j][0], K, tail_mask, qk[j][1]);
        if (WILD) {
            make_wild<W, P>(qk[j][0], tail_mask);
            if (CRICK) make_wild<W, P>(qk[j][1], tail_mask);
        }
        // start from the current global minimum (monotone, so a stale read is still an upper
        // bound): lets a query that already reached the floor 0 skip work
        best[j] = qpos[j] < prm.q_end ? prm.out[qpos[j] - prm.q_begin] : 0u;
    }
    const uint32_t qg_lo = qbase >> 5, qg_hi = (qbase + kThreads * Q - 1) >> 5;

    // ---- target chunk ----
    const uint32_t tile0 = blockIdx.x * prm.tiles_per_chunk;
    uint32_t ntiles = prm.tiles_total > tile0 ? prm.tiles_total - tile0 : 0u;
    if (ntiles > prm.tiles_per_chunk) ntiles = prm.tiles_per_chunk;

    auto all_zero = [&]() {
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < Q; ++j) o |= best[j];
        return o == 0;
    };
    // early exit at the reporting floor: nothing can beat distance 0
    if (__syncthreads_and(all_zero()) || ntiles == 0) return;

    auto issue = [&](uint32_t t) {
        const uint32_t st = t & 1;
        const size_t off = (size_t)(tile0 + t) * kTileGroups;
        constexpr uint32_t bytes = kTileWords * 4;
        mbar_expect_tx(&full_bar[st], bytes * NARR);
#pragma unroll
        for (int p = 0; p < P; ++p) tma_bulk_g2s(&tile[st][p][0], prm.t.plane(p) + off, bytes, &full_bar[st]);
        tma_bulk_g2s(&tile[st][P][0], prm.t.valid() + off, bytes, &full_bar[st]);
    };
    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) issue(0);

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t st = t & 1;
        if (tid == 0 && t + 1 < ntiles) issue(t + 1);
        mbar_wait(&full_bar[st], (t >> 1) & 1);

        const uint32_t gabs0 = (tile0 + t) * kTileGroups;
        for (uint32_t g = 0; g < (uint32_t)kTileGroups; ++g) {
            const uint32_t v = tile[st][P][g];
            if (v == 0) continue;  // warp-uniform: every lane reads the same word
            uint32_t cw[P][W + 1];
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int w = 0; w <= W; ++w) cw[p][w] = tile[st][p][g + w];
            const uint32_t gabs = gabs0 + g;
            // -z sends every group down the per-candidate path (positions are needed)
            const bool self_here = (prm.self_exclude && gabs >= qg_lo && gabs <= qg_hi) || prm.zfilt;
            if (RANGED) {
                // sub-range sweeps: skip groups no query of this CTA can pair with
                const long long t0 = (long long)gabs << 5, t1 = t0 + 31;
                const long long qlo = qbase, qhi = (long long)qbase + kThreads * Q - 1;
                const long long kk = (long long)K - 1;
                bool any = (t1 >= qlo + prm.w_lo && t0 <= qhi + prm.w_hi) ||
                           (t1 >= qlo - prm.w_hi && t0 <= qhi - prm.w_lo);
                if (CRICK)
                    any = any || (t1 >= prm.c1_lo - kk - qhi && t0 <= prm.c1_hi - kk - qlo) ||
                          (t1 >= prm.c2_lo - kk - qhi && t0 <= prm.c2_hi - kk - qlo);
                if (!any) continue;
            }
            if (!RANGED && v == 0xffffffffu && !self_here) {
                // fast path: 32 valid candidates, no self pair possible
#pragma unroll 4
                for (uint32_t s = 0; s < 32; ++s) {
                    Kmer<W, P> c;
                    cut_candidate<W, P>(cw, s, tail_mask, c);
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        uint32_t d = kmer_dist<W, P, WILD>(qk[j][0], c);
                        if (CRICK) d = min(d, kmer_dist<W, P, WILD>(qk[j][1], c));
                        best[j] = min(best[j], d);
                    }
                }
            } else {
                // slow path: chromosome ends (invalid candidates) and the CTA's own positions
                const uint32_t cpos0 = gabs << 5;
#pragma unroll 1
                for (uint32_t s = 0; s < 32; ++s) {
                    if (!((v >> s) & 1u)) continue;
                    Kmer<W, P> c;
                    cut_candidate<W, P>(cw, s, tail_mask, c);
                    const uint32_t cpos = cpos0 + s;
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        uint32_t d = kmer_dist<W, P, WILD>(qk[j][0], c);
                        // the reference skips only EXACT self hits (SfxArray.cpp:4418-4419,
                        // :4585-4594); for identity comparison the self pair is always exact
                        if (self_here && cpos == qpos[j] && d == 0) d = kNoDist;
                        if (!RANGED && d == 0 && prm.zfilt && !zfilter_pass(prm, qpos[j], cpos)) d = kNoDist;
                        if (RANGED) {
                            const long long dl = (long long)cpos - (long long)qpos[j];
                            const long long ad = dl < 0 ? -dl : dl;
                            if (ad < prm.w_lo || ad > prm.w_hi) d = kNoDist;
                        }
                        if (CRICK) {
                            uint32_t dr = kmer_dist<W, P, WILD>(qk[j][1], c);
                            if (RANGED) {
                                const long long cc = (long long)cpos + (long long)qpos[j] + K - 1;
                                if (!((cc >= prm.c1_lo && cc <= prm.c1_hi) || (cc >= prm.c2_lo && cc <= prm.c2_hi)))
                                    dr = kNoDist;
                            }
                            d = min(d, dr);
                        }
                        best[j] = min(best[j], d);
                    }
                }
            }
        }
        // all lanes are done with tile[st] (buffer reuse) + floor test
        if (__syncthreads_and(all_zero())) {
            if (t + 1 < ntiles) mbar_wait(&full_bar[st ^ 1], ((t + 1) >> 1) & 1);  // drain TMA
            break;
        }
    }

#pragma unroll
    for (int j = 0; j < Q; ++j)
        if (qpos[j] < prm.q_end) atomicMin(&prm.out[qpos[j] - prm.q_begin], best[j]);
}

// ---------------------------------------------------------------------------------------
// generic path for K > 32*kMaxRegW: one query per thread, distance accumulated word by word
// for the 32 candidates of a group (accumulators in registers, K-mers cut on the fly).
// ---------------------------------------------------------------------------------------
template <int P, bool CRICK, bool WILD>
__global__ void __launch_bounds__(kThreads, 2) allpairs_min_generic_kernel(const AllPairsParams prm,
                                                                           const ImageView rcq) {
    // q_rc_image: 3 planes of the reverse-complemented query concat (position i of the rc
    // sequence is base len-1-i complemented) so rc(K-mer at pos) = K-mer at len-K-pos.
    constexpr int S = CRICK ? 2 : 1;
    const uint32_t tid = threadIdx.x;
    const uint32_t K = prm.K;
    const uint32_t W = (K + 31) >> 5;
    const uint32_t tail_mask = (K & 31) ? ((1u << (K & 31)) - 1u) : 0xffffffffu;
    const uint32_t qpos = prm.q_begin + blockIdx.y * kThreads + tid;
    const bool q_ok = qpos < prm.q_end && qpos + K <= prm.q.len;
    uint32_t best = q_ok ? prm.out[qpos - prm.q_begin] : 0u;
    const uint32_t qp[2] = {q_ok ? qpos : 0u, q_ok ? prm.q.len - K - qpos : 0u};
    const uint32_t *qimg[2] = {prm.q.base, rcq.base};
    const uint32_t qnwp[2] = {prm.q.stride, rcq.stride};

    const uint32_t g_begin = blockIdx.x * prm.tiles_per_chunk * kTileGroups;
    uint32_t g_end = g_begin + prm.tiles_per_chunk * kTileGroups;
    uint32_t g_tot = prm.tiles_total * kTileGroups;
    if (prm.groups_limit && g_tot > prm.groups_limit) g_tot = prm.groups_limit;
    if (g_end > g_tot) g_end = g_tot;
    const uint32_t qg = qpos >> 5;

    for (uint32_t g = g_begin; g < g_end; ++g) {
        const uint32_t v = __ldg(prm.t.valid() + g);
        if (v == 0) continue;
        uint32_t acc[S][32];
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2)
#pragma unroll
            for (int s = 0; s < 32; ++s) acc[s2][s] = 0;
        for (uint32_t w = 0; w < W; ++w) {
            const uint32_t m = (w == W - 1) ? tail_mask : 0xffffffffu;
            uint32_t qw[S][P];
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) {
                const uint32_t wi = (qp[s2] >> 5) + w, sh = qp[s2] & 31;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint32_t *pl = qimg[s2] + (size_t)p * qnwp[s2] + wi;
                    qw[s2][p] = __funnelshift_r(pl[0], pl[1], sh) & m;
                }
                if (WILD && P == 3) {
                    const uint32_t nw = ~qw[s2][P - 1] & m;
                    qw[s2][0] &= nw;
                    qw[s2][1] &= nw;
                    qw[s2][P - 1] = nw;
                }
            }
            uint32_t c0[P], c1[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                c0[p] = __ldg(prm.t.plane(p) + g + w);
                c1[p] = __ldg(prm.t.plane(p) + g + w + 1);
            }
#pragma unroll
            for (uint32_t s = 0; s < 32; ++s) {
                uint32_t c[P];
#pragma unroll
                for (int p = 0; p < P; ++p) c[p] = __funnelshift_r(c0[p], c1[p], s) & m;
#pragma unroll
                for (int s2 = 0; s2 < S; ++s2) {
                    uint32_t x = qw[s2][0] ^ c[0];
                    x |= qw[s2][1] ^ c[1];
                    if (P == 3) {
                        if (WILD) x = (x & qw[s2][P - 1]) | c[P - 1];
                        else x |= qw[s2][P - 1] ^ c[P - 1];
                    }
                    acc[s2][s] += __popc(x);
                }
            }
        }
        const bool self_here = prm.self_exclude && g == qg;
#pragma unroll
        for (uint32_t s = 0; s < 32; ++s) {
            if (!((v >> s) & 1u)) continue;
            uint32_t d = acc[0][s];
            const uint32_t cpos = (g << 5) + s;
            if (self_here && cpos == qpos && d == 0) d = kNoDist;
            if (d == 0 && prm.zfilt && !zfilter_pass(prm, qpos, cpos)) d = kNoDist;
            if (prm.ranged) {
                const long long dl = (long long)cpos - (long long)qpos;
                const long long ad = dl < 0 ? -dl : dl;
                if (ad < prm.w_lo || ad > prm.w_hi) d = kNoDist;
            }
            if (CRICK) {
                uint32_t dr = acc[S - 1][s];
                if (prm.ranged) {
                    const long long cc = (long long)cpos + (long long)qpos + K - 1;
                    if (!((cc >= prm.c1_lo && cc <= prm.c1_hi) || (cc >= prm.c2_lo && cc <= prm.c2_hi))) dr = kNoDist;
                }
                d = min(d, dr);
            }
            best = min(best, d);
        }
    }
    if (q_ok) atomicMin(&prm.out[qpos - prm.q_begin], best);
}

// planes of the reverse-complemented sequence: rc[i] = cpl(x[len-1-i]); positions >= len are
// EOS in all planes.  One thread per output word.  rc_arr has the geometry of a packed image
// (kImageArrays arrays; the valid array stays unused).
__device__ __forceinline__ uint32_t bits_at(const uint32_t *pl, int64_t start) {
    // 32 bits of the plane starting at bit offset start (may be negative: those bits read 0)
    if (start <= -32) return 0;
    if (start < 0) return pl[0] << (uint32_t)(-start);
    const uint32_t wi = (uint32_t)(start >> 5), sh = (uint32_t)(start & 31);
    return __funnelshift_r(pl[wi], pl[wi + 1], sh);
}
__global__ void __launch_bounds__(256) revcomp_planes_kernel(ImageView q, uint32_t *__restrict__ rc_arr) {
    // thread per array word; rc_arr points at the first (front pad) word of rc plane 0
    const uint32_t w_abs = blockIdx.x * blockDim.x + threadIdx.x;
    if (w_abs >= q.stride) return;
    const long long j = (long long)w_abs - kFrontPadWords;  // logical word of the rc sequence
    uint32_t o[3] = {0, 0, 0};
    uint32_t in = 0;
    if (j >= 0 && j * 32 < (long long)q.len) {
        const int64_t start = (int64_t)q.len - 32 * (j + 1);
#pragma unroll
        for (int p = 0; p < 3; ++p) o[p] = __brev(bits_at(q.plane(p), start));
        const int64_t n = (int64_t)q.len - j * 32;
        in = n >= 32 ? 0xffffffffu : ((1u << (uint32_t)n) - 1u);
        const uint32_t flip = ~o[2] & in;
        o[0] ^= flip;
        o[1] ^= flip;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        o[p] |= ~in;
        rc_arr[(size_t)p * q.stride + w_abs] = o[p];
    }
    // 2-bit code array of the reverse-complemented sequence (rows of the rectangular band mode)
    uint32_t *code = rc_arr + (size_t)4 * q.stride + 2 * (size_t)w_abs;
    code[0] = spread16(o[0]) | (spread16(o[1]) << 1);
    code[1] = spread16(o[0] >> 16) | (spread16(o[1] >> 16) << 1);
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
cudaError_t launch_revcomp_planes(ImageView q, ImageView rc, cudaStream_t st) {
    revcomp_planes_kernel<<<(q.stride + 255) / 256, 256, 0, st>>>(
        q, const_cast<uint32_t *>(rc.base) - kFrontPadWords);
    return cudaGetLastError();
}
cudaError_t launch_pack(const uint8_t *d_concat, ImageView img, uint32_t *d_flags, cudaStream_t st) {
    const uint32_t threads = img.stride * 2;
    pack_kernel<<<(threads + 255) / 256, 256, 0, st>>>(
        d_concat, img, const_cast<uint32_t *>(img.base) - kFrontPadWords, d_flags);
    return cudaGetLastError();
}
cudaError_t launch_valid(ImageView img, uint32_t K, unsigned long long *d_count, cudaStream_t st) {
    valid_kernel<<<(img.stride + 255) / 256, 256, 0, st>>>(
        img, const_cast<uint32_t *>(img.valid()) - kFrontPadWords, K, d_count);
    return cudaGetLastError();
}
cudaError_t launch_fill_u32(uint32_t *d, uint32_t n, uint32_t v, cudaStream_t st) {
    if (!n) return cudaSuccess;
    fill_u32_kernel<<<(n + 255) / 256, 256, 0, st>>>(d, n, v);
    return cudaGetLastError();
}
cudaError_t launch_histogram_u16(const uint16_t *d_v, uint32_t n, uint32_t nbins, unsigned long long *d_hist,
                                 cudaStream_t st) {
    if (!n || !nbins) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t want = (n + 255) / 256;
    const uint32_t grid = want < (uint32_t)sms * 8 ? want : (uint32_t)sms * 8;  // a few CTAs per SM, grid-stride
    histogram_u16_kernel<<<grid, 256, nbins * sizeof(uint32_t), st>>>(d_v, n, nbins, d_hist);
    return cudaGetLastError();
}
cudaError_t launch_finalize(const uint32_t *d_min32, ImageView q, uint32_t q_begin, uint32_t n,
                            uint32_t K, uint32_t clamp, int max_wild, uint16_t *d_out16,
                            cudaStream_t st) {
    if (!n) return cudaSuccess;
    finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_min32, q, q_begin, n, K, clamp, max_wild, d_out16);
    return cudaGetLastError();
}

template <int W, int P, int Q>
static cudaError_t launch_ap(const AllPairsParams &p, bool crick, dim3 grid, cudaStream_t st) {
    const bool wild = P == 3 && p.wildcard;
    if (p.ranged) {  // sub-range sweeps exist only in exhaustive (identity) mode
        if (crick) allpairs_min_kernel<W, P, Q, true, false, true><<<grid, kThreads, 0, st>>>(p);
        else allpairs_min_kernel<W, P, Q, false, false, true><<<grid, kThreads, 0, st>>>(p);
    } else if (wild) {
        if (crick) allpairs_min_kernel<W, P, Q, true, (P == 3), false><<<grid, kThreads, 0, st>>>(p);
        else allpairs_min_kernel<W, P, Q, false, (P == 3), false><<<grid, kThreads, 0, st>>>(p);
    } else {
        if (crick) allpairs_min_kernel<W, P, Q, true, false, false><<<grid, kThreads, 0, st>>>(p);
        else allpairs_min_kernel<W, P, Q, false, false, false><<<grid, kThreads, 0, st>>>(p);
    }
    return cudaGetLastError();
}

int queries_per_thread(uint32_t W, bool three_planes) {
    if (W == 1) return 8;
    if (W == 2) return 4;
    return three_planes ? 2 : 4;
}

cudaError_t launch_allpairs(const AllPairsParams &p, bool three_planes, bool crick,
                            cudaStream_t st, int *n_ctas) {
    const uint32_t W = (p.K + 31) / 32;
    if (W > (uint32_t)kMaxRegW) return cudaErrorInvalidValue;
    const uint32_t nq = p.q_end - p.q_begin;
    const int Q = queries_per_thread(W, three_planes);
    const uint32_t qb = (nq + kThreads * Q - 1) / (kThreads * Q);
    const uint32_t chunks = (p.tiles_total + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    dim3 grid(chunks, qb);
    if (n_ctas) *n_ctas = (int)(chunks * qb);
    if (!nq || !chunks) return cudaSuccess;
#define K4B_CASE(WW, PP, QQ) return launch_ap<WW, PP, QQ>(p, crick, grid, st)
    if (!three_planes) {
        switch (W) {
            case 1: K4B_CASE(1, 2, 8);
            case 2: K4B_CASE(2, 2, 4);
            case 3: K4B_CASE(3, 2, 4);
            case 4: K4B_CASE(4, 2, 4);
        }
    } else {
        switch (W) {
            case 1: K4B_CASE(1, 3, 8);
            case 2: K4B_CASE(2, 3, 4);
            case 3: K4B_CASE(3, 3, 2);
            case 4: K4B_CASE(4, 3, 2);
        }
    }
#undef K4B_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_allpairs_generic(const AllPairsParams &p, bool three_planes, bool crick,
                                    ImageView rc, cudaStream_t st, int *n_ctas) {
    const uint32_t nq = p.q_end - p.q_begin;
    const uint32_t qb = (nq + kThreads - 1) / kThreads;
    const uint32_t chunks = (p.tiles_total + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    dim3 grid(chunks, qb);
    if (n_ctas) *n_ctas = (int)(chunks * qb);
    if (!nq || !chunks) return cudaSuccess;
    if (three_planes && p.wildcard) {
        if (crick) allpairs_min_generic_kernel<3, true, true><<<grid, kThreads, 0, st>>>(p, rc);
        else allpairs_min_generic_kernel<3, false, true><<<grid, kThreads, 0, st>>>(p, rc);
    } else if (three_planes) {
        if (crick) allpairs_min_generic_kernel<3, true, false><<<grid, kThreads, 0, st>>>(p, rc);
        else allpairs_min_generic_kernel<3, false, false><<<grid, kThreads, 0, st>>>(p, rc);
    } else {
        if (crick) allpairs_min_generic_kernel<2, true, false><<<grid, kThreads, 0, st>>>(p, rc);
        else allpairs_min_generic_kernel<2, false, false><<<grid, kThreads, 0, st>>>(p, rc);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// integer-pipe microbenchmarks (register resident, no memory traffic)
// ---------------------------------------------------------------------------------------
constexpr int kMbChains = 8;
constexpr int kMbUnroll = 32;

template <int WHICH>
__global__ void __launch_bounds__(256) microbench_kernel(int iters, uint32_t *sink) {
    uint32_t x[kMbChains];
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) x[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
    uint32_t a = threadIdx.x | 0x10101u, b = blockIdx.x * 77u + 5u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kMbUnroll; ++u) {
#pragma unroll
            for (int k = 0; k < kMbChains; ++k) {
                if (WHICH == 0) {
                    asm volatile("popc.b32 %0, %0;" : "+r"(x[k]));
                } else if (WHICH == 1) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
                } else if (WHICH == 2) {
                    uint32_t t, m, d;
                    asm volatile("lop3.b32 %0, %1, %2, %2, 0x3c;" : "=r"(t) : "r"(x[k]), "r"(a));   // x^a
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0xbe;" : "=r"(m) : "r"(b), "r"(x[k]), "r"(t));
                    asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(m));
                    asm volatile("min.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(d));
                } else if (WHICH == 3) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
                } else if (WHICH == 4) {  // IMAD alone (FMA pipe)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                } else if (WHICH == 5) {  // LOP3 and IMAD 1:1 - do the ALU and FMA pipes overlap?
                    if (k & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
                } else if (WHICH == 6) {  // funnel shift alone
                    asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[k]) : "r"(a));
                } else if (WHICH == 7) {  // IMAD.WIDE alone
                    unsigned long long w;
                    asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x[k]), "r"(a), "l"((unsigned long long)b << 32));
                    x[k] = (uint32_t)(w >> 32);
                } else if (WHICH == 9) {  // IMAD.HI alone
                    asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                } else if (WHICH == 10) {  // LOP3 : IMAD.HI 3 : 1
                    if ((k & 3) == 3) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
                } else {                  // LOP3 : IMAD 3 : 1
                    if ((k & 3) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
                    else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < kMbChains; ++k) r ^= x[k];
    if (r == 0x12345678u) sink[0] = r;  // keep the chains alive
}

cudaError_t launch_microbench(int which, int iters, uint32_t *d_sink, int *blocks, int *threads,
                              int *ops_per_thread_iter, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int nb = sms * 8;
    *blocks = nb;
    *threads = 256;
    *ops_per_thread_iter = kMbChains * kMbUnroll;
    switch (which) {
        case 0: microbench_kernel<0><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 1: microbench_kernel<1><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 2: microbench_kernel<2><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 3: microbench_kernel<3><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 4: microbench_kernel<4><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 5: microbench_kernel<5><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 6: microbench_kernel<6><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 7: microbench_kernel<7><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 8: microbench_kernel<8><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 9: microbench_kernel<9><<<nb, 256, 0, st>>>(iters, d_sink); break;
        case 10: microbench_kernel<10><<<nb, 256, 0, st>>>(iters, d_sink); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace k4b
