// k4b_diag.cu - diagonal-band engine for the exhaustive all-vs-all minimum (sm_100a).
//
// The reference walks one diagonal of the K-mer pair matrix at a time and updates the
// distance incrementally: +1 for a mismatch entering at the 3' end, -1 for one leaving at the
// 5' end, and every cell lowers the minima of BOTH K-mers of the pair
// (ngskit4b/hammings.cpp:3183-3287 GHamDistWatson, :3300-3489 GHamDistCrick).  This kernel
// keeps that O(1)-per-pair recurrence and that symmetry but runs 32 diagonals per thread in
// bit-sliced form:
//   * a thread owns 32 consecutive diagonals s0..s0+31; the 32 running distances live as
//     NP bit-planes c[0..NP) ("vertical counters"), one bit per diagonal per plane;
//   * one row step: the entering base A[e] is compared with the 32 column bases
//     B[e+s0 .. e+s0+31] in two LOP3 (bit-plane XOR/OR against the broadcast row base), the
//     same for the leaving base A[e-K]; the +1/-1/0 update of all 32 counters is a ripple of
//     2 LOP3 per plane; no POPC, no per-pair minimum;
//   * counters are stored biased so that "distance < T" is simply "top plane bit clear":
//     T is an upper bound of every running minimum this warp can still improve (max over
//     256-position blocks, refreshed between launches).  Only flagged cells - a few per
//     million on non-repetitive sequence - take the slow path that rebuilds the exact
//     distance and issues atomicMin on both K-mers of the pair;
//   * a warp = 32 threads = 1024 consecutive diagonals over the same rows, so column words are
//     read fully coalesced and row words are warp-uniform; a CTA = 8 such warps.
// Watson: A = B = sequence, diagonals s >= 1.  Crick: B = reverse-complemented sequence Y
// (Y[a] = cpl(x[len-1-a])), so that d(i, rc j) = HD(x-Kmer i, Y-Kmer M-j); the diagonal
// s = j' - i is mirror symmetric about its middle, only its first half is walked.
// Exact, bit-identical results; measured against the same oracle as the POPC engine.
#include "k4b_kernels.cuh"

#include <stdlib.h>

namespace k4b {

constexpr int kDiagWarps = 8;        // super-bands (of 1024 diagonals) per CTA
constexpr int kSuperBand = 1024;     // diagonals per warp
constexpr int kGroupDiags = kDiagWarps * kSuperBand;

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}

__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {  // a * b + c on the FMA pipe
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ uint32_t sext_bit(uint32_t x, uint32_t t) {
    return (uint32_t)((int32_t)(x << (31u - t)) >> 31);
}

// running-minimum upper bounds per 2^shift positions (valid K-mer starts only)
__global__ void __launch_bounds__(256) blockmax_kernel(const uint32_t *__restrict__ best, ImageView a,
                                                       uint32_t n_pos, uint32_t shift,
                                                       uint32_t *__restrict__ blockmax, uint32_t n_blocks,
                                                       uint32_t *__restrict__ tmax, uint32_t low_floor,
                                                       uint32_t *__restrict__ n_low) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_blocks) return;
    const uint32_t base = warp << shift, span = 1u << shift;
    uint32_t m = 0;
    for (uint32_t o = lane; o < span; o += 32) {
        const uint32_t pos = base + o;
        if (pos < n_pos && ((a.valid()[pos >> 5] >> (pos & 31)) & 1u)) m = max(m, best[pos]);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) {
        blockmax[warp] = m;
        if (tmax && m) atomicMax(tmax, m);
        // blocks whose threshold the narrow-counter instance would have to lift (more flags)
        if (n_low && m && m < low_floor) atomicAdd(n_low, 1u);
    }
}

template <int P>
struct SideWords {      // words of one side (enter or leave) for a block of up to 32 row steps
    uint32_t ra[P];     // row bases: bit t = plane bit of A[idx + t]           (warp uniform)
    uint32_t xa[P];     // column window: bits [idx+s0, idx+s0+32) of B planes   (per lane)
    uint32_t xb[P];     //                bits [idx+s0+32, idx+s0+64)
};

template <int P>
__device__ __forceinline__ void load_side(const ImageView &a, const ImageView &b, long long idx,
                                          long long s0, SideWords<P> &w) {
    const uint32_t wi = (uint32_t)(idx >> 5), sh = (uint32_t)(idx & 31);
    const long long wb = idx + s0;
    const long long wj = wb >> 5;  // floor: may be slightly negative (front pad)
    const uint32_t ws = (uint32_t)(wb & 31);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t *ap = a.plane(p) + wi;
        w.ra[p] = __funnelshift_r(__ldg(ap), __ldg(ap + 1), sh);
        const uint32_t *bp = b.plane(p) + wj;
        const uint32_t w0 = __ldg(bp), w1 = __ldg(bp + 1), w2 = __ldg(bp + 2);
        w.xa[p] = __funnelshift_r(w0, w1, ws);
        w.xb[p] = __funnelshift_r(w1, w2, ws);
    }
}

// mismatch word of row step t: bit k = [A[idx+t] != B[idx+t+s0+k]].  WILD (targeted rules):
// a row symbol >= N matches any column ACGT base but never a column N (SfxArray.cpp:4266-4296)
template <int P, bool WILD>
__device__ __forceinline__ uint32_t mism_word(const SideWords<P> &w, uint32_t t) {
    uint32_t m = __funnelshift_r(w.xa[0], w.xb[0], t) ^ sext_bit(w.ra[0], t);
    m |= __funnelshift_r(w.xa[1], w.xb[1], t) ^ sext_bit(w.ra[1], t);
    if (P == 3) {
        const uint32_t xn = __funnelshift_r(w.xa[2], w.xb[2], t), bn = sext_bit(w.ra[2], t);
        if (WILD) m = (m & ~bn) | xn;
        else m |= xn ^ bn;
    }
    return m;
}

// slow path, out of line so that the unrolled hot loop stays small: exact distance of the
// flagged diagonals of `row`, then min-update of both K-mers of each pair
template <int NP>
struct Counters {
    uint32_t c[NP];
};
struct FlushCtx {  // what the slow path needs, by value
    const uint32_t *valid_row, *valid_col;
    uint32_t *best;
    long long Mrow, Mcol;
    int row_flip, col_flip, update_cols;
};
// one row of the slow path: flagged diagonals k of `row`, distance = stored - bias + bit k of adj
template <int NP>
__device__ __forceinline__ void diag_flush_row(const Counters<NP> &cs, uint32_t flags, uint32_t bias,
                                               long long row, long long s0, uint32_t adj, const FlushCtx &fc) {
    if (row > fc.Mrow) return;
    const long long prow = fc.row_flip ? fc.Mrow - row : row;
    if (!((fc.valid_row[prow >> 5] >> (prow & 31)) & 1u)) return;
    while (flags) {
        const uint32_t k = __ffs(flags) - 1;
        flags &= flags - 1;
        const long long col = row + s0 + k;
        if (col < 0 || col > fc.Mcol) continue;
        const long long pcol = fc.col_flip ? fc.Mcol - col : col;
        if (!((fc.valid_col[pcol >> 5] >> (pcol & 31)) & 1u)) continue;
        uint32_t val = 0;
#pragma unroll
        for (int b = 0; b < NP; ++b) val |= ((cs.c[b] >> k) & 1u) << b;
        const uint32_t d = val - bias + ((adj >> k) & 1u);
        if (d < fc.best[prow]) atomicMin(&fc.best[prow], d);
        if (fc.update_cols && d < fc.best[pcol]) atomicMin(&fc.best[pcol], d);
    }
}
// The counters hold min(d(row), d(row+1)) of a row pair (see diag_min_kernel): d(row) = stored +
// adj0 bit, d(row+1) = stored + adj1 bit.  rows = 1: single row, the counters hold d(row) itself.
template <int NP>
__device__ __noinline__ void diag_flush(Counters<NP> cs, uint32_t flags, uint32_t bias, long long row,
                                        long long s0, uint32_t adj0, uint32_t adj1, int rows, FlushCtx fc) {
    diag_flush_row<NP>(cs, flags, bias, row, s0, adj0, fc);
    if (rows == 2) diag_flush_row<NP>(cs, flags, bias, row + 1, s0, adj1, fc);
}

// EWIN (two planes only): the mismatch words of a 32-row block come from a per-thread table of the
// four possible mismatch WINDOWS (one per row base) in shared memory - see the main loop.
template <int NP, int P, bool WILD, bool EWIN>
__global__ void __launch_bounds__(kDiagWarps * 32) diag_min_kernel(const DiagParams prm) {
    static_assert(!EWIN || (P == 2 && !WILD), "the window-table path handles pure ACGT sets");
    if (prm.sel) {  // device-side choice between the narrow- and the full-counter instance
        const uint32_t tg = __ldg(prm.tmax_ptr);
        const bool narrow = tg <= prm.sel_limit && __ldg(prm.low_ptr) <= prm.low_max;
        if ((prm.sel == 1) != narrow) return;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t grp_local = blockIdx.x / prm.n_seg, seg = blockIdx.x - grp_local * prm.n_seg;
    const uint32_t l = prm.l_first + grp_local;
    const long long q = (long long)(l / prm.q_span) * prm.q_period + prm.q_lo + l % prm.q_span;
    const long long grp = (long long)prm.part + q * prm.nparts;
    const long long S0cta = prm.s_first + grp * kGroupDiags;
    const long long S1cta = S0cta + kGroupDiags - 1;
    const long long Mrow = prm.Mrow, Mcol = prm.Mcol;
    // rows of diagonal s: max(0,-s) .. min(Mrow, Mcol-s); the CTA takes the union over its
    // diagonals (cells outside a diagonal's own range are invalid or mirror duplicates)
    long long row_lo = S1cta < 0 ? -S1cta : 0;
    long long row_hi = Mcol - S0cta;
    if (prm.mode == kDiagCrick) row_hi >>= 1;  // first half of every mirror-symmetric diagonal
    if (row_hi > Mrow) row_hi = Mrow;
    const long long r_start = row_lo + (long long)seg * prm.rows_per_seg;
    if (r_start > row_hi) return;
    long long r_end = r_start + prm.rows_per_seg;
    if (r_end > row_hi + 1) r_end = row_hi + 1;
    const long long S0w = S0cta + (long long)warp * kSuperBand;
    const long long s0 = S0w + lane * 32;
    const uint32_t K = prm.K;

    // ---- threshold: upper bound of every minimum this warp can still lower ----
    uint32_t tmax = prm.t_fixed;
    if (prm.mode != kDiagRect) {
        auto scan = [&](long long lo, long long hi, long long top) {
            if (lo < 0) lo = 0;
            if (hi > top) hi = top;
            if (lo > hi) return;
            for (long long bk = (lo >> prm.bm_shift) + lane; bk <= (hi >> prm.bm_shift); bk += 32)
                tmax = max(tmax, __ldg(prm.blockmax + bk));
        };
        scan(r_start, r_end - 1, Mrow);
        const long long c_lo = r_start + S0w, c_hi = r_end - 1 + S0w + kSuperBand - 1;
        if (!prm.col_flip) scan(c_lo, c_hi, Mcol);
        else scan(Mcol - c_hi, Mcol - c_lo, Mcol);
        tmax = __reduce_max_sync(0xffffffffu, tmax);
    }
    if (tmax == 0) return;  // every K-mer in reach already sits at the floor 0
    // a larger threshold is always safe (more flags, never a miss): lift it so that the biased
    // maximum K + 2^(NP-1) - T still fits NP planes when the narrow-counter instance runs
    const uint32_t t_floor = (K + 1 > (1u << (NP - 1))) ? K + 1 - (1u << (NP - 1)) : 0u;
    const uint32_t T = tmax > t_floor ? tmax : t_floor;
    const uint32_t bias = (1u << (NP - 1)) - T;  // stored = distance + bias; distance < T <=> top bit clear

    uint32_t c[NP];
#pragma unroll
    for (int b = 0; b < NP; ++b) c[b] = ((bias >> b) & 1u) ? 0xffffffffu : 0u;

    FlushCtx fc;
    fc.valid_row = prm.va.valid();
    fc.valid_col = prm.vb.valid();
    fc.best = prm.best;
    fc.Mrow = Mrow;
    fc.Mcol = Mcol;
    fc.row_flip = prm.row_flip;
    fc.col_flip = prm.col_flip;
    fc.update_cols = prm.update_cols;
    auto flush = [&](uint32_t flags, long long row, uint32_t adj0, uint32_t adj1, int rows) {
        Counters<NP> cs;
#pragma unroll
        for (int b = 0; b < NP; ++b) cs.c[b] = c[b];
        diag_flush<NP>(cs, flags, bias, row, s0, adj0, adj1, rows, fc);
    };

    // ---- warm-up: the first K bases of the window enter, nothing leaves ----
    long long e = r_start;
    {
        uint32_t remaining = K;
        while (remaining) {
            const uint32_t n = remaining < 32 ? remaining : 32;
            SideWords<P> we;
            load_side<P>(prm.a, prm.b, e, s0, we);
#pragma unroll 1
            for (uint32_t t = 0; t < n; ++t) {
                uint32_t act = mism_word<P, WILD>(we, t);
#pragma unroll
                for (int b = 0; b < NP; ++b) {  // ripple increment
                    const uint32_t old = c[b];
                    c[b] = old ^ act;
                    act &= old;
                }
            }
            e += n;
            remaining -= n;
        }
    }
    {
        const uint32_t f = ~c[NP - 1];
        if (f) flush(f, r_start, 0u, 0u, 1);
    }

    // ---- main: row r: base r+K-1 enters (mismatch word en), base r-1 leaves (lv) ----
    // Rows are taken in PAIRS (r, r+1) and the counters hold C = min(d(r), d(r+1)) = d(r+1) - p,
    // p = [the step r -> r+1 was +1]: one ripple and one threshold test serve two rows, and the
    // test "C < T" is exact for both (no slack).  Update over a pair, from the previous pair's p:
    //   C' = C + p_prev + (en1 - lv1) - q2,   q2 = [the step r -> r+1 is -1] = lv2 & ~en2
    // i.e. a signed delta in [-2, 2], added in two's complement: 2 LOP3 per plane per PAIR.
    // On a flag the slow path rebuilds d(r) = C + q2 and d(r+1) = C + p.
    uint32_t pprev = 0;  // the counters start as d(r_start) itself
    // Written as explicit LOP3s (a = 0xF0, b = 0xCC, c = 0xAA): left to itself the compiler
    // re-associates the carry chain into a wider, longer form (34 instead of 28 LOP3 per pair).
    // (u1 u0) = p_prev + en1, (v1 v0) = lv1 + q2, both 0..2; pn = next pair's p_prev
    auto pair_core = [&](uint32_t u0, uint32_t u1, uint32_t v0, uint32_t v1, uint32_t pn) {
        const uint32_t b0 = lop3<0x0c>(u0, v0, 0u);       // u - v: borrow out of bit 0 = ~u0 & v0
        const uint32_t d1 = lop3<0x96>(u1, v1, b0);       // bit 1 of u - v
        const uint32_t sg = lop3<0x8e>(u1, v1, b0);       // borrow out of bit 1 = sign = maj(~u1, v1, b0)
        uint32_t old = c[0];
        c[0] = lop3<0x96>(old, u0, v0);
        uint32_t carry = lop3<0x60>(old, u0, v0);         // old & (u0 ^ v0)
        old = c[1];
        c[1] = lop3<0x96>(old, d1, carry);
        carry = lop3<0xe8>(old, d1, carry);               // majority
#pragma unroll
        for (int b = 2; b < NP; ++b) {
            old = c[b];
            c[b] = lop3<0x96>(old, sg, carry);
            if (b + 1 < NP) carry = lop3<0xe8>(old, sg, carry);
        }
        pprev = pn;
    };
    auto pair_step = [&](uint32_t e1, uint32_t l1, uint32_t e2, uint32_t l2, uint32_t &q2) {
        const uint32_t pn = lop3<0x30>(e2, l2, 0u);       // e2 & ~l2
        q2 = lop3<0x30>(l2, e2, 0u);                      // l2 & ~e2
        pair_core(lop3<0x3c>(pprev, e1, 0u), lop3<0xc0>(pprev, e1, 0u), lop3<0x3c>(l1, q2, 0u),
                  lop3<0xc0>(l1, q2, 0u), pn);
    };
    // Two-plane fast path: the XOR of a column window with the broadcast row bit runs on the FMA
    // pipe, which this kernel otherwise leaves idle: x ^ R = x * s + R for R in {0, ~0}, s = R | 1
    // (IMAD; measured to overlap LOP3 fully, profiles/r01_intpipe_microbench.json).  (R, s) of the
    // 32 rows of a block come from a per-warp shared-memory table (one LDS.128 per row and side,
    // broadcast) instead of 2 uniform-pipe shifts per word, and the OR of the two plane terms is
    // folded into the LOP3s that consume the mismatch words: 2*NP + 18 ALU ops per row pair
    // (two planes; three planes: + 4 SHF + 4 LOP3).
    // Three planes ride the same path: identity compare adds a third IMAD term (x2 ^ R2) and one OR
    // per word; the wildcard rule ((m & ~R2) | x2, mism_word above) masks with the row's plane-2 bit
    // inside that OR and takes the column's plane-2 window as the second term.  Either way a mismatch
    // word is the OR of two terms (t1, t2), which the consumers fold in.
    constexpr bool kFma = true;
    __shared__ uint4 rowtab[EWIN ? 1 : kDiagWarps][EWIN ? 1 : 32][2];
    __shared__ uint4 rowtab2[P == 3 ? kDiagWarps : 1][32];  // plane 2: {R2 enter, s2 enter, R2 leave, s2 leave}
    // Window-table path (EWIN).  For a block of 32 rows the column window of a thread is fixed
    // (64 bits per plane), and the mismatch word of row step t is bits [t, t+32) of
    //     M_b = (x0 ^ b0) | (x1 ^ b1)          b = the row's base (2 bits, warp uniform)
    // so each thread keeps M_A, M_C, M_G, M_T of the enter and of the leave side (2 x 4 x 8 bytes) in its
    // OWN shared-memory slots (no synchronisation) and a row step is one LDS.64 at [slot + base offset]
    // and one funnel shift: 4 SHF + (2*NP + 8) LOP3 + 1 ISETP per row PAIR instead of 8 SHF + (2*NP + 9)
    // LOP3 + 1 ISETP + 8 IMAD, and the per-warp row table disappears from the block prologue.  The base
    // offsets of the 32 rows are packed 2 bits per row into a 64-bit word that REDUX leaves in uniform
    // registers, so the per-row address arithmetic runs on the uniform datapath.
    __shared__ uint2 mtab[EWIN ? 2 : 1][EWIN ? 4 : 1][EWIN ? kDiagWarps * 32 : 1];
    // the two terms of the mismatch word of row step t of one side; ER = the side's row of rowtab,
    // r2 / s2 = its plane-2 entries of rowtab2
    auto terms = [&](const SideWords<P> &w, uint32_t t, const uint4 &ER, uint32_t r2, uint32_t s2, uint32_t &t1,
                     uint32_t &t2) {
        const uint32_t a = __funnelshift_r(w.xa[0], w.xb[0], t) * ER.y + ER.x;
        const uint32_t b = __funnelshift_r(w.xa[1], w.xb[1], t) * ER.w + ER.z;
        if constexpr (P == 2) {
            t1 = a;
            t2 = b;
        } else {
            const uint32_t x2 = __funnelshift_r(w.xa[2], w.xb[2], t);
            if constexpr (WILD) {
                t1 = lop3<0x54>(a, b, r2);  // (a | b) & ~R2: a wildcard row symbol mismatches only a column N
                t2 = x2;
            } else {
                t1 = lop3<0xfc>(a, b, 0u);
                t2 = x2 * s2 + r2;
            }
        }
    };
    const uint32_t lane_mul = 1u << (31u - lane);
    long long row = r_start + 1;
    while (row < r_end) {
        const long long left = r_end - row;
        if constexpr (EWIN) {
            if (left >= 32) {
                uint32_t ce_lo, ce_hi, cl_lo, cl_hi;  // 2-bit base codes of the 32 enter / leave rows (uniform)
                auto build = [&](long long idx, int side, uint32_t &c_lo, uint32_t &c_hi) {
                    // base codes of rows idx .. idx+31: a 64-bit window of the 2-bit code array (warp
                    // uniform; REDUX leaves it in uniform registers)
                    const uint32_t *cw = prm.a.code2() + (idx >> 4);
                    const uint32_t csh = (uint32_t)(idx & 15) * 2u;
                    const uint32_t k0 = __ldg(cw), k1 = __ldg(cw + 1), k2 = __ldg(cw + 2);
                    c_lo = __reduce_or_sync(0xffffffffu, __funnelshift_r(k0, k1, csh));
                    c_hi = __reduce_or_sync(0xffffffffu, __funnelshift_r(k1, k2, csh));
                    const long long wb = idx + s0;
                    const long long wj = wb >> 5;  // floor: may be slightly negative (front pad)
                    const uint32_t ws = (uint32_t)(wb & 31);
                    const uint32_t *b0 = prm.b.plane(0) + wj, *b1 = prm.b.plane(1) + wj;
                    const uint32_t p0 = __ldg(b0), p1 = __ldg(b0 + 1), p2 = __ldg(b0 + 2);
                    const uint32_t q0 = __ldg(b1), q1 = __ldg(b1 + 1), q2w = __ldg(b1 + 2);
                    const uint32_t x0a = __funnelshift_r(p0, p1, ws), x0b = __funnelshift_r(p1, p2, ws);
                    const uint32_t x1a = __funnelshift_r(q0, q1, ws), x1b = __funnelshift_r(q1, q2w, ws);
                    // mismatch windows (x0 ^ b0) | (x1 ^ b1) for the four row bases, one LOP3 per word
                    mtab[side][0][threadIdx.x] = make_uint2(lop3<0xfc>(x0a, x1a, 0u), lop3<0xfc>(x0b, x1b, 0u));  // A (00):  x0 |  x1
                    mtab[side][1][threadIdx.x] = make_uint2(lop3<0xcf>(x0a, x1a, 0u), lop3<0xcf>(x0b, x1b, 0u));  // C (01): ~x0 |  x1
                    mtab[side][2][threadIdx.x] = make_uint2(lop3<0xf3>(x0a, x1a, 0u), lop3<0xf3>(x0b, x1b, 0u));  // G (10):  x0 | ~x1
                    mtab[side][3][threadIdx.x] = make_uint2(lop3<0x3f>(x0a, x1a, 0u), lop3<0x3f>(x0b, x1b, 0u));  // T (11): ~x0 | ~x1
                };
                build(row + K - 1, 0, ce_lo, ce_hi);
                build(row - 1, 1, cl_lo, cl_hi);
                const char *slot = reinterpret_cast<const char *>(&mtab[0][0][threadIdx.x]);
                constexpr uint32_t kBaseStride = kDiagWarps * 32 * sizeof(uint2);  // 2048 B: bits 11, 12 select the base
                constexpr uint32_t kSideStride = 4 * kBaseStride;
                static_assert(kBaseStride == 2048, "offset extraction below assumes a 2 KB base stride");
                // byte offset of the window of row step t: (code of row t) * 2048
                auto off = [&](uint32_t c_lo, uint32_t c_hi, uint32_t t) -> uint32_t {
                    const uint32_t w = t < 16 ? c_lo : c_hi, b = (t & 15u) * 2u;  // the code sits at bits b, b+1 of w
                    return (b <= 11 ? (w << (11 - b)) : (w >> (b - 11))) & 0x1800u;
                };
                auto word = [&](uint32_t side_off, uint32_t o, uint32_t t) -> uint32_t {
                    const uint2 w = *reinterpret_cast<const uint2 *>(slot + side_off + o);
                    return __funnelshift_r(w.x, w.y, t);
                };
#pragma unroll
                for (uint32_t t = 0; t < 32; t += 2) {
                    const uint32_t e1 = word(0, off(ce_lo, ce_hi, t), t);
                    const uint32_t l1 = word(kSideStride, off(cl_lo, cl_hi, t), t);
                    const uint32_t e2 = word(0, off(ce_lo, ce_hi, t + 1), t + 1);
                    const uint32_t l2 = word(kSideStride, off(cl_lo, cl_hi, t + 1), t + 1);
                    uint32_t q2;
                    pair_step(e1, l1, e2, l2, q2);
                    const uint32_t f = ~c[NP - 1];
                    if (__builtin_expect(f != 0, 0)) flush(f, row + t, q2, pprev, 2);
                }
                row += 32;
                continue;
            }
        }
        SideWords<P> we, wl;
        load_side<P>(prm.a, prm.b, row + K - 1, s0, we);
        load_side<P>(prm.a, prm.b, row - 1, s0, wl);
        if (!EWIN && left >= 32) {
            if constexpr (kFma) {
                __syncwarp();
                {
                    // R = all-ones iff bit `lane` of the row word; the multiply (FMA pipe) brings the bit to
                    // the top, one arithmetic shift spreads it; s = 2R + 1 = R | 1 on the FMA pipe as well
                    auto rs = [&](uint32_t w, uint32_t &R, uint32_t &sgn) {
                        R = (uint32_t)((int32_t)imad(w, lane_mul, 0u) >> 31);
                        sgn = imad(R, 2u, 1u);
                    };
                    uint4 E, L;
                    rs(we.ra[0], E.x, E.y);
                    rs(we.ra[1], E.z, E.w);
                    rs(wl.ra[0], L.x, L.y);
                    rs(wl.ra[1], L.z, L.w);
                    rowtab[warp][lane][0] = E;
                    rowtab[warp][lane][1] = L;
                    if constexpr (P == 3) {
                        uint4 X;
                        rs(we.ra[2], X.x, X.y);
                        rs(wl.ra[2], X.z, X.w);
                        rowtab2[warp][lane] = X;
                    }
                }
                __syncwarp();
#pragma unroll
                for (uint32_t t = 0; t < 32; t += 2) {
                    const uint4 E1 = rowtab[warp][t][0], L1 = rowtab[warp][t][1];
                    const uint4 E2 = rowtab[warp][t + 1][0], L2 = rowtab[warp][t + 1][1];
                    uint4 X1 = make_uint4(0, 0, 0, 0), X2 = X1;
                    if constexpr (P == 3) {
                        X1 = rowtab2[warp][t];
                        X2 = rowtab2[warp][t + 1];
                    }
                    uint32_t a1, b1, c1, d1, a2, b2, c2, d2;
                    terms(we, t, E1, X1.x, X1.y, a1, b1);
                    terms(wl, t, L1, X1.z, X1.w, c1, d1);
                    terms(we, t + 1, E2, X2.x, X2.y, a2, b2);
                    terms(wl, t + 1, L2, X2.z, X2.w, c2, d2);
                    const uint32_t l2 = lop3<0xfc>(c2, d2, 0u);     // lv2 = c2 | d2
                    const uint32_t pn = lop3<0x54>(a2, b2, l2);     // (a2 | b2) & ~lv2
                    const uint32_t q2 = lop3<0x02>(a2, b2, l2);     // lv2 & ~(a2 | b2)
                    pair_core(lop3<0x1e>(pprev, a1, b1), lop3<0xe0>(pprev, a1, b1), lop3<0x1e>(q2, c1, d1),
                              lop3<0xe0>(q2, c1, d1), pn);
                    const uint32_t f = ~c[NP - 1];
                    if (__builtin_expect(f != 0, 0)) flush(f, row + t, q2, pprev, 2);
                }
            } else {
#pragma unroll
                for (uint32_t t = 0; t < 32; t += 2) {
                    uint32_t q2;
                    pair_step(mism_word<P, WILD>(we, t), mism_word<P, WILD>(wl, t), mism_word<P, WILD>(we, t + 1),
                              mism_word<P, WILD>(wl, t + 1), q2);
                    const uint32_t f = ~c[NP - 1];
                    if (__builtin_expect(f != 0, 0)) flush(f, row + t, q2, pprev, 2);
                }
            }
            row += 32;
        } else {
#pragma unroll 1
            for (uint32_t t = 0; t < (uint32_t)left; t += 2) {
                const bool two = t + 1 < (uint32_t)left;
                uint32_t q2;
                pair_step(mism_word<P, WILD>(we, t), mism_word<P, WILD>(wl, t),
                          two ? mism_word<P, WILD>(we, t + 1) : 0u, two ? mism_word<P, WILD>(wl, t + 1) : 0u, q2);
                const uint32_t f = ~c[NP - 1];
                if (f) flush(f, row + t, q2, pprev, two ? 2 : 1);
            }
            row += left;
        }
    }
}

// ---- two diagonal words per thread ---------------------------------------------------------------
// Same algorithm as the window-table path of diag_min_kernel (two planes, pure ACGT), but a thread owns
// 64 consecutive diagonals = two counter words, a warp 2048 diagonals and a CTA (4 warps, 128 threads)
// the same 8192-diagonal group.  The two words share what is per thread rather than per word: the
// column windows overlap (three aligned words serve both: (a,b) and (b,c)), so the block prologue
// builds 3 instead of 4 mismatch words per row base and side, ONE LDS.128 per row step and side fetches
// the windows of both words, the per-row address arithmetic on the uniform datapath is done once, and
// two independent ripple chains are in flight per thread.  ALU work per word is unchanged
// (4 SHF + 2*NP+8 LOP3 + 1 ISETP per row pair).
template <int NP>
__device__ __forceinline__ void pair_core_fn(uint32_t (&c)[NP], uint32_t &pprev, uint32_t u0, uint32_t u1, uint32_t v0,
                                             uint32_t v1, uint32_t pn) {
    const uint32_t b0 = lop3<0x0c>(u0, v0, 0u);       // u - v: borrow out of bit 0 = ~u0 & v0
    const uint32_t d1 = lop3<0x96>(u1, v1, b0);       // bit 1 of u - v
    const uint32_t sg = lop3<0x8e>(u1, v1, b0);       // borrow out of bit 1 = sign = maj(~u1, v1, b0)
    uint32_t old = c[0];
    c[0] = lop3<0x96>(old, u0, v0);
    uint32_t carry = lop3<0x60>(old, u0, v0);         // old & (u0 ^ v0)
    old = c[1];
    c[1] = lop3<0x96>(old, d1, carry);
    carry = lop3<0xe8>(old, d1, carry);               // majority
#pragma unroll
    for (int b = 2; b < NP; ++b) {
        old = c[b];
        c[b] = lop3<0x96>(old, sg, carry);
        if (b + 1 < NP) carry = lop3<0xe8>(old, sg, carry);
    }
    pprev = pn;
}
template <int NP>
__device__ __forceinline__ void pair_step_fn(uint32_t (&c)[NP], uint32_t &pprev, uint32_t e1, uint32_t l1, uint32_t e2,
                                             uint32_t l2, uint32_t &q2) {
    const uint32_t pn = lop3<0x30>(e2, l2, 0u);       // e2 & ~l2
    q2 = lop3<0x30>(l2, e2, 0u);                      // l2 & ~e2
    pair_core_fn<NP>(c, pprev, lop3<0x3c>(pprev, e1, 0u), lop3<0xc0>(pprev, e1, 0u), lop3<0x3c>(l1, q2, 0u),
                     lop3<0xc0>(l1, q2, 0u), pn);
}

constexpr int kDw = 2;                       // diagonal words per thread
constexpr int kDwWarps = kDiagWarps / kDw;   // 4 warps of 2048 diagonals
template <int NP>
__global__ void __launch_bounds__(kDwWarps * 32) diag_min2_kernel(const DiagParams prm) {
    if (prm.sel) {  // device-side choice between the narrow- and the full-counter instance
        const uint32_t tg = __ldg(prm.tmax_ptr);
        const bool narrow = tg <= prm.sel_limit && __ldg(prm.low_ptr) <= prm.low_max;
        if ((prm.sel == 1) != narrow) return;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t grp_local = blockIdx.x / prm.n_seg, seg = blockIdx.x - grp_local * prm.n_seg;
    const uint32_t l = prm.l_first + grp_local;
    const long long q = (long long)(l / prm.q_span) * prm.q_period + prm.q_lo + l % prm.q_span;
    const long long grp = (long long)prm.part + q * prm.nparts;
    const long long S0cta = prm.s_first + grp * kGroupDiags;
    const long long S1cta = S0cta + kGroupDiags - 1;
    const long long Mrow = prm.Mrow, Mcol = prm.Mcol;
    long long row_lo = S1cta < 0 ? -S1cta : 0;
    long long row_hi = Mcol - S0cta;
    if (prm.mode == kDiagCrick) row_hi >>= 1;  // first half of every mirror-symmetric diagonal
    if (row_hi > Mrow) row_hi = Mrow;
    const long long r_start = row_lo + (long long)seg * prm.rows_per_seg;
    if (r_start > row_hi) return;
    long long r_end = r_start + prm.rows_per_seg;
    if (r_end > row_hi + 1) r_end = row_hi + 1;
    constexpr int kBand = kSuperBand * kDw;  // diagonals per warp
    const long long S0w = S0cta + (long long)warp * kBand;
    const long long s0 = S0w + lane * (32 * kDw);
    const uint32_t K = prm.K;

    // ---- threshold: upper bound of every minimum this warp can still lower ----
    uint32_t tmax = prm.t_fixed;
    if (prm.mode != kDiagRect) {
        auto scan = [&](long long lo, long long hi, long long top) {
            if (lo < 0) lo = 0;
            if (hi > top) hi = top;
            if (lo > hi) return;
            for (long long bk = (lo >> prm.bm_shift) + lane; bk <= (hi >> prm.bm_shift); bk += 32)
                tmax = max(tmax, __ldg(prm.blockmax + bk));
        };
        scan(r_start, r_end - 1, Mrow);
        const long long c_lo = r_start + S0w, c_hi = r_end - 1 + S0w + kBand - 1;
        if (!prm.col_flip) scan(c_lo, c_hi, Mcol);
        else scan(Mcol - c_hi, Mcol - c_lo, Mcol);
        tmax = __reduce_max_sync(0xffffffffu, tmax);
    }
    if (tmax == 0) return;
    const uint32_t t_floor = (K + 1 > (1u << (NP - 1))) ? K + 1 - (1u << (NP - 1)) : 0u;
    const uint32_t T = tmax > t_floor ? tmax : t_floor;
    const uint32_t bias = (1u << (NP - 1)) - T;

    uint32_t c[kDw][NP];
#pragma unroll
    for (int w = 0; w < kDw; ++w)
#pragma unroll
        for (int b = 0; b < NP; ++b) c[w][b] = ((bias >> b) & 1u) ? 0xffffffffu : 0u;

    FlushCtx fc;
    fc.valid_row = prm.va.valid();
    fc.valid_col = prm.vb.valid();
    fc.best = prm.best;
    fc.Mrow = Mrow;
    fc.Mcol = Mcol;
    fc.row_flip = prm.row_flip;
    fc.col_flip = prm.col_flip;
    fc.update_cols = prm.update_cols;
    auto flush = [&](int w, uint32_t flags, long long row, uint32_t adj0, uint32_t adj1, int rows) {
        Counters<NP> cs;
#pragma unroll
        for (int b = 0; b < NP; ++b) cs.c[b] = c[w][b];
        diag_flush<NP>(cs, flags, bias, row, s0 + 32 * w, adj0, adj1, rows, fc);
    };

    // ---- warm-up: the first K bases of the window enter, nothing leaves (word by word) ----
    {
        long long e = r_start;
        uint32_t remaining = K;
        while (remaining) {
            const uint32_t n = remaining < 32 ? remaining : 32;
#pragma unroll
            for (int w = 0; w < kDw; ++w) {
                SideWords<2> we;
                load_side<2>(prm.a, prm.b, e, s0 + 32 * w, we);
#pragma unroll 1
                for (uint32_t t = 0; t < n; ++t) {
                    uint32_t act = mism_word<2, false>(we, t);
#pragma unroll
                    for (int b = 0; b < NP; ++b) {  // ripple increment
                        const uint32_t old = c[w][b];
                        c[w][b] = old ^ act;
                        act &= old;
                    }
                }
            }
            e += n;
            remaining -= n;
        }
    }
#pragma unroll
    for (int w = 0; w < kDw; ++w) {
        const uint32_t f = ~c[w][NP - 1];
        if (f) flush(w, f, r_start, 0u, 0u, 1);
    }

    // ---- main: row pairs, the counters of a word hold min(d(r), d(r+1)) (see diag_min_kernel) ----
    uint32_t pprev[kDw] = {0u, 0u};
    __shared__ uint4 mtab2[2][4][kDwWarps * 32];  // per thread and side: {Ma, Mb, Mc, -} for the four row bases
    constexpr uint32_t kBaseStride = kDwWarps * 32 * sizeof(uint4);  // 2048 B: bits 11, 12 select the base
    constexpr uint32_t kSideStride = 4 * kBaseStride;
    static_assert(kBaseStride == 2048, "offset extraction below assumes a 2 KB base stride");
    long long row = r_start + 1;
    while (row < r_end) {
        const long long left = r_end - row;
        if (left >= 32) {
            uint32_t ce_lo, ce_hi, cl_lo, cl_hi;  // 2-bit base codes of the 32 enter / leave rows (uniform)
            auto build = [&](long long idx, int side, uint32_t &c_lo, uint32_t &c_hi) {
                const uint32_t *cw = prm.a.code2() + (idx >> 4);
                const uint32_t csh = (uint32_t)(idx & 15) * 2u;
                const uint32_t k0 = __ldg(cw), k1 = __ldg(cw + 1), k2 = __ldg(cw + 2);
                c_lo = __reduce_or_sync(0xffffffffu, __funnelshift_r(k0, k1, csh));
                c_hi = __reduce_or_sync(0xffffffffu, __funnelshift_r(k1, k2, csh));
                const long long wb = idx + s0;
                const long long wj = wb >> 5;  // floor: may be slightly negative (front pad)
                const uint32_t ws = (uint32_t)(wb & 31);
                const uint32_t *b0 = prm.b.plane(0) + wj, *b1 = prm.b.plane(1) + wj;
                const uint32_t p0 = __ldg(b0), p1 = __ldg(b0 + 1), p2 = __ldg(b0 + 2), p3 = __ldg(b0 + 3);
                const uint32_t q0 = __ldg(b1), q1 = __ldg(b1 + 1), q2w = __ldg(b1 + 2), q3 = __ldg(b1 + 3);
                const uint32_t xa = __funnelshift_r(p0, p1, ws), xb = __funnelshift_r(p1, p2, ws), xc = __funnelshift_r(p2, p3, ws);
                const uint32_t ya = __funnelshift_r(q0, q1, ws), yb = __funnelshift_r(q1, q2w, ws), yc = __funnelshift_r(q2w, q3, ws);
                // mismatch windows (x ^ b0) | (y ^ b1) for the four row bases, one LOP3 per word
                mtab2[side][0][threadIdx.x] = make_uint4(lop3<0xfc>(xa, ya, 0u), lop3<0xfc>(xb, yb, 0u), lop3<0xfc>(xc, yc, 0u), 0u);
                mtab2[side][1][threadIdx.x] = make_uint4(lop3<0xcf>(xa, ya, 0u), lop3<0xcf>(xb, yb, 0u), lop3<0xcf>(xc, yc, 0u), 0u);
                mtab2[side][2][threadIdx.x] = make_uint4(lop3<0xf3>(xa, ya, 0u), lop3<0xf3>(xb, yb, 0u), lop3<0xf3>(xc, yc, 0u), 0u);
                mtab2[side][3][threadIdx.x] = make_uint4(lop3<0x3f>(xa, ya, 0u), lop3<0x3f>(xb, yb, 0u), lop3<0x3f>(xc, yc, 0u), 0u);
            };
            build(row + K - 1, 0, ce_lo, ce_hi);
            build(row - 1, 1, cl_lo, cl_hi);
            const char *slot = reinterpret_cast<const char *>(&mtab2[0][0][threadIdx.x]);
            auto off = [&](uint32_t c_lo, uint32_t c_hi, uint32_t t) -> uint32_t {
                const uint32_t wv = t < 16 ? c_lo : c_hi, b = (t & 15u) * 2u;  // the code sits at bits b, b+1 of wv
                return (b <= 11 ? (wv << (11 - b)) : (wv >> (b - 11))) & 0x1800u;
            };
#pragma unroll
            for (uint32_t t = 0; t < 32; t += 2) {
                const uint4 E1 = *reinterpret_cast<const uint4 *>(slot + off(ce_lo, ce_hi, t));
                const uint4 L1 = *reinterpret_cast<const uint4 *>(slot + kSideStride + off(cl_lo, cl_hi, t));
                const uint4 E2 = *reinterpret_cast<const uint4 *>(slot + off(ce_lo, ce_hi, t + 1));
                const uint4 L2 = *reinterpret_cast<const uint4 *>(slot + kSideStride + off(cl_lo, cl_hi, t + 1));
                uint32_t q2a, q2b;
                pair_step_fn<NP>(c[0], pprev[0], __funnelshift_r(E1.x, E1.y, t), __funnelshift_r(L1.x, L1.y, t),
                                 __funnelshift_r(E2.x, E2.y, t + 1), __funnelshift_r(L2.x, L2.y, t + 1), q2a);
                pair_step_fn<NP>(c[1], pprev[1], __funnelshift_r(E1.y, E1.z, t), __funnelshift_r(L1.y, L1.z, t),
                                 __funnelshift_r(E2.y, E2.z, t + 1), __funnelshift_r(L2.y, L2.z, t + 1), q2b);
                const uint32_t fa = ~c[0][NP - 1], fb = ~c[1][NP - 1];
                if (__builtin_expect(fa != 0, 0)) flush(0, fa, row + t, q2a, pprev[0], 2);
                if (__builtin_expect(fb != 0, 0)) flush(1, fb, row + t, q2b, pprev[1], 2);
            }
            row += 32;
            continue;
        }
#pragma unroll
        for (int w = 0; w < kDw; ++w) {  // fewer than 32 rows left: word by word on the plain path
            SideWords<2> we, wl;
            load_side<2>(prm.a, prm.b, row + K - 1, s0 + 32 * w, we);
            load_side<2>(prm.a, prm.b, row - 1, s0 + 32 * w, wl);
#pragma unroll 1
            for (uint32_t t = 0; t < (uint32_t)left; t += 2) {
                const bool two = t + 1 < (uint32_t)left;
                uint32_t q2;
                pair_step_fn<NP>(c[w], pprev[w], mism_word<2, false>(we, t), mism_word<2, false>(wl, t),
                                 two ? mism_word<2, false>(we, t + 1) : 0u, two ? mism_word<2, false>(wl, t + 1) : 0u, q2);
                const uint32_t f = ~c[w][NP - 1];
                if (f) flush(w, f, row + t, q2, pprev[w], two ? 2 : 1);
            }
        }
        row += left;
    }
}

cudaError_t launch_blockmax(const uint32_t *d_best, ImageView a, uint32_t n_pos, uint32_t shift,
                            uint32_t *d_blockmax, uint32_t n_blocks, uint32_t *d_tmax, uint32_t low_floor,
                            uint32_t *d_n_low, cudaStream_t st) {
    if (!n_blocks) return cudaSuccess;
    const uint32_t threads = n_blocks * 32;
    blockmax_kernel<<<(threads + 255) / 256, 256, 0, st>>>(d_best, a, n_pos, shift, d_blockmax, n_blocks, d_tmax,
                                                           low_floor, d_n_low);
    return cudaGetLastError();
}

int diag_planes_for_k(uint32_t K) {
    int b = 0;
    while ((1u << b) < K + 1) ++b;  // 2^b >= K+1 >= any threshold T
    return b + 1;
}

static cudaError_t launch_diag_dw(const DiagParams &p, int np, dim3 grid, cudaStream_t st) {
    switch (np) {
#define K4B_DIAG2_CASE(N) \
    case N: diag_min2_kernel<N><<<grid, kDwWarps * 32, 0, st>>>(p); break
        K4B_DIAG2_CASE(5);
        K4B_DIAG2_CASE(6);
        K4B_DIAG2_CASE(7);
        K4B_DIAG2_CASE(8);
        K4B_DIAG2_CASE(9);
        K4B_DIAG2_CASE(10);
        K4B_DIAG2_CASE(11);
        K4B_DIAG2_CASE(12);
        K4B_DIAG2_CASE(13);
        K4B_DIAG2_CASE(14);
#undef K4B_DIAG2_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int P, bool WILD, bool EWIN>
static cudaError_t launch_diag_p(const DiagParams &p, int np, dim3 grid, cudaStream_t st) {
    switch (np) {
#define K4B_DIAG_CASE(N) \
    case N: diag_min_kernel<N, P, WILD, EWIN><<<grid, kDiagWarps * 32, 0, st>>>(p); break
        K4B_DIAG_CASE(5);
        K4B_DIAG_CASE(6);
        K4B_DIAG_CASE(7);
        K4B_DIAG_CASE(8);
        K4B_DIAG_CASE(9);
        K4B_DIAG_CASE(10);
        K4B_DIAG_CASE(11);
        K4B_DIAG_CASE(12);
        K4B_DIAG_CASE(13);
        K4B_DIAG_CASE(14);
#undef K4B_DIAG_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_diag(const DiagParams &p, bool three_planes, int np, uint32_t n_groups, cudaStream_t st,
                        unsigned long long *n_ctas) {
    if (n_ctas) *n_ctas = 0;
    if (!n_groups || !p.n_seg) return cudaSuccess;
    const unsigned long long total = (unsigned long long)n_groups * p.n_seg;
    if (total > 0x7fffffffull) return cudaErrorInvalidValue;
    if (n_ctas) *n_ctas = total;
    if (np < 5) np = 5;
    if ((1u << np) < p.K + 1) return cudaErrorInvalidValue;  // K + bias would not fit
    dim3 grid((unsigned)total);
    if (!three_planes) {
        // K4B_DIAG_EWIN=0 selects the previous two-plane path (row table + IMAD broadcast XOR)
        const char *e = getenv("K4B_DIAG_EWIN");  // read per launch: tests toggle it
        const int ewin = e ? atoi(e) : 1;
        // two diagonal words per thread (diag_min2_kernel) when its registers still allow 7-8 CTAs per SM
        // (NP <= 6: 63-64 registers; at NP = 7, 70 registers, it measured 5 % slower) and the launch is large enough to fill the GPU with 128-thread
        // CTAs; K4B_DIAG_DW=1|2 forces a variant (tests, profiles)
        const char *d = getenv("K4B_DIAG_DW");
        const bool dw2 = d ? atoi(d) == 2 : (np <= 6 && total >= 4736ull);
        if (ewin && dw2) return launch_diag_dw(p, np, grid, st);
        return ewin ? launch_diag_p<2, false, true>(p, np, grid, st) : launch_diag_p<2, false, false>(p, np, grid, st);
    }
    return p.wild ? launch_diag_p<3, true, false>(p, np, grid, st) : launch_diag_p<3, false, false>(p, np, grid, st);
}

}  // namespace k4b
