// acs_hist.cuh -- add-compare-select with SURVIVOR-HISTORY TAGS (K <= 7), whole-frame batch decode.
//
// Replaces ViterbiDecoder_Scalar::update / bfly / renormalise (include/viterbi/viterbi_decoder_scalar.h:28-153) for a batch, like
// acs_pair.cuh, but with a different way of producing the decisions.  Every path metric register carries, BELOW the metric, the
// last decisions of the SURVIVOR PATH that ends in that state:
//
//     FMT 0 (uint8_t  error metrics): register q = [metric A << 8 | history A (8 bits)] | [metric B << 8 | history B] << 16
//                                     two frames per thread, packed 16x2 arithmetic (VIADD.16x2 / VIADDMNMX.U16x2), period 8 steps
//     FMT 1 (uint16_t error metrics): register q =  metric << 16 | history (16 bits)
//                                     one frame per thread, 32-bit arithmetic (IADD / VIADDMNMX.U32), period 16 steps
//
//   * At step k of a period the path-1 operand of each compare gets the tag 2^k added (the history fields of the operands hold only
//     bits < k, so the field never carries into the metric), and min() then
//       - compares the metrics exactly (metric fields differ -> the history fields are irrelevant),
//       - breaks a metric tie in favour of path 0, because hist(a) < 2^k <= hist(b): the reference's strict '>' (scalar.h:123-124),
//       - and keeps the winner's history: bit k of the new state's history IS the decision, bits < k are the decisions along the
//         winner's own survivor path.
//     So a butterfly is 2 adds + 2 fused add-min = FOUR instructions for its 4 add-compare-selects (x2 frames for FMT 0): no
//     predicates, no per-step bit collection, no tag clearing (acs_pair.cuh's predicate form needs 10, its tagged form 8).
//     Metric arithmetic wraps modulo 2^8 / 2^16 exactly like the reference's error_t.
//     TIE_SIMD tags path 0 instead (a tie selects path 1, x86/viterbi_decoder_avx_u16.h:112-115) and the bits are inverted on output.
//   * every period the 2^(K-1) history fields of a frame are written out as one "history record" (2^(K-1) bits per frame and step -
//     exactly the size of the reference's decision rows, core.h:49-83) and cleared.  Traceback (traceback_hist_kernel) then moves a
//     whole period per lookup: the history of the current state holds the decisions of its survivor path, which are the decoded
//     bits AND determine the state one period earlier (core.h:96-113 applied 8 / 16 times).
//   * the metrics ping-pong between two register sets in logical state order, so the step code does not depend on the step index:
//     the hot loop is two steps (about 5 KB of SASS) and there is no register renaming.
//   * renormalisation (scalar.h:48, 139-153) works on the metric fields; the history fields ride along.
//
// Record layout (both formats): NW = 2^(K-1)/2 32-bit words per thread and period;
//     FMT 0: word w = bytes [A:state 2w, B:state 2w, A:state 2w+1, B:state 2w+1];   FMT 1: word w = halfwords [state 2w, state 2w+1]
//     dec = uint32 [n_blocks][n_periods][NW / VW][32 lanes][VW], VW = min(4, NW)  (coalesced 16-byte stores);
//     a block is one warp: 64 frames (FMT 0, lane l = frames 2l, 2l+1) or 32 frames (FMT 1, lane l = frame l).
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"

namespace vitb {

template <int FMT>
struct HistLane;

template <>
struct HistLane<0> {
    static constexpr int HB = 8, FPT = 2, SBY = 1;
    static constexpr uint32_t TAG1 = 0x00010001u, METRIC_MASK = 0xff00ff00u;
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
    static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return __vsub2(a, b); }
    static __device__ __forceinline__ uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
};

template <>
struct HistLane<1> {
    static constexpr int HB = 16, FPT = 1, SBY = 2;
    static constexpr uint32_t TAG1 = 1u, METRIC_MASK = 0xffff0000u;
    static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { return a + b; }
    static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { return a - b; }
    static __device__ __forceinline__ uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_u32(a, b, c); }
    static __device__ __forceinline__ uint32_t min3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }
    static __device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return min(a, b); }
};

// kernel-uniform constants in the lane format (FMT 1 derives them from the packed 16x2 parameters the host fills for sh = 0)
struct HistConsts {
    uint32_t c_low, c_high, c_high_n, c_inv, thr, init_start, init_other;      // high - s = c_high - s = ~s + c_high_n
    uint32_t thr_m1;       // FMT 0: thr - 1 per half (trigger test with one packed max)
    bool always;           // FMT 0: a threshold of 0 renormalises every step
};

template <int FMT>
__device__ __forceinline__ HistConsts hist_consts(const AcsParams& p) {
    HistConsts c;
    if constexpr (FMT == 0) {
        c.c_low = p.c_low2; c.c_inv = p.c_inv2; c.thr = p.thr2; c.init_start = p.init_start2; c.init_other = p.init_other2;
        c.c_high_n = p.c_high2;                                    // (high << 8) + 1 per half
        c.c_high = p.c_high2 - 0x00010001u;                        // high << 8 per half
        c.always = (p.thr2 & 0xffffu) == 0u || (p.thr2 >> 16) == 0u;
        c.thr_m1 = c.always ? 0xffffffffu : p.thr2 - 0x00010001u;
    } else {
        c.c_low = p.c_low2 & 0xffff0000u;                          // (-low) << 16
        c.c_high = (p.c_high2 & 0xffff0000u) - 0x00010000u;        // high << 16  (c_high2 holds high + 1 per half)
        c.c_high_n = c.c_high + 1u;
        c.c_inv = p.c_inv2 & 0xffff0000u; c.thr = p.thr2 & 0xffff0000u;
        c.init_start = p.init_start2 & 0xffff0000u; c.init_other = p.init_other2 & 0xffff0000u;
        c.always = false; c.thr_m1 = 0u;
    }
    return c;
}

template <class C>
struct HistShape {
    static constexpr int NS = C::NS;
    static constexpr int NW = NS / 2;                    // record words per thread and period (both formats)
    static constexpr int VW = NW < 4 ? NW : 4;           // words per vector store
};

template <class C, int FMT, int TIE_SIMD, int J, bool FLIP = false>
__device__ __forceinline__ void hist_bfly_at(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t (&T)[C::NP],
                                             const uint32_t (&TT)[C::NP], const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP]) {
    using LN = HistLane<FMT>;
    constexpr int H = C::NS / 2;
    constexpr uint32_t pat = bfly_pattern<C>(uint32_t(J));
    constexpr uint32_t ipat = (~pat) & uint32_t(C::NP - 1);
    if constexpr (!TIE_SIMD && FLIP) {          // same butterfly, the (0|X)-side results first (source order only: VITB_HIST_FLIP)
        const uint32_t b1 = LN::add(x[J + H], TT[pat]);
        const uint32_t b0 = LN::add(x[J + H], VT[ipat]);
        y[2 * J + 1] = LN::addmin(x[J], V[ipat], b1);
        y[2 * J] = LN::addmin(x[J], T[pat], b0);
    } else if constexpr (!TIE_SIMD) {
        const uint32_t b0 = LN::add(x[J + H], VT[ipat]);                    // (1|X) + inverted, tagged    scalar.h:114
        const uint32_t b1 = LN::add(x[J + H], TT[pat]);                     // (1|X) + total,    tagged    scalar.h:116
        y[2 * J] = LN::addmin(x[J], T[pat], b0);                            // min((0|X) + total, b0)      scalar.h:113,127
        y[2 * J + 1] = LN::addmin(x[J], V[ipat], b1);                       // min((0|X) + inverted, b1)   scalar.h:115,128
    } else {
        const uint32_t a0 = LN::add(x[J], TT[pat]);                         // tag on path 0: a tie selects path 1
        const uint32_t a1 = LN::add(x[J], VT[ipat]);
        y[2 * J] = LN::addmin(x[J + H], V[ipat], a0);
        y[2 * J + 1] = LN::addmin(x[J + H], T[pat], a1);
    }
}

// Source order of the butterflies of a step (VITB_HIST_ORDER; profiles/microbench/hist_kernel_variants.cu times the config-2 kernel
// for each on the bench's own frames): 0 = ascending (the default), 1 = descending, 2 = the two halves of the state space interleaved,
// 3 = grouped by branch pattern pair {p, ~p}, 4 = grouped by pattern, 6 = 3 descending inside the groups.  The butterfly is bound by
// register-file reads (profiles/r02_summary.md) and ptxas keeps much of the source order, so grouping the butterflies of a pattern
// lets consecutive instructions take their table operand from the operand reuse cache - but the register allocation that comes with
// each order matters as much: config 2 on real frames, B200: 0.635 ms (0, 136 registers), 0.674 (1, 128 + spills), 0.649 (3, 133),
// 0.656 (4, 127 + spills), 0.646 (6, 134); 3 with __launch_bounds__(128, 1): 0.635 (149 registers).  No order beats the plain one.
#ifndef VITB_HIST_ORDER
#define VITB_HIST_ORDER 0
#endif
template <class C>
__host__ __device__ constexpr int hist_bfly_order(int i) {
    constexpr int H = C::NS / 2;
    if (VITB_HIST_ORDER == 1) return H - 1 - i;
    if (VITB_HIST_ORDER == 2) return (i & 1) ? (H / 2 + i / 2) : (i / 2);
    if (VITB_HIST_ORDER == 3 || VITB_HIST_ORDER == 5 || VITB_HIST_ORDER == 6) {      // stable sort by min(pattern, ~pattern)
        int n = 0;
        for (uint32_t g = 0; g < uint32_t(C::NP); g++)
            for (int jj = 0; jj < H; jj++) {
                const int j = (VITB_HIST_ORDER == 6) ? (H - 1 - jj) : jj;
                const uint32_t p = bfly_pattern<C>(uint32_t(j)), q = (~p) & uint32_t(C::NP - 1), key = p < q ? p : q;
                if (key == g) { if (n == i) return j; n++; }
            }
    }
    if (VITB_HIST_ORDER == 4) {      // stable sort by pattern
        int n = 0;
        for (uint32_t g = 0; g < uint32_t(C::NP); g++)
            for (int j = 0; j < H; j++)
                if (bfly_pattern<C>(uint32_t(j)) == g) { if (n == i) return j; n++; }
    }
    return i;
}
// VITB_HIST_FLIP: 1 = every second butterfly (in source order) computes its two results in the opposite order; 2 = butterflies whose
// pattern is the larger of {p, ~p} do (so that consecutive butterflies of a group start with the same table operand)
#ifndef VITB_HIST_FLIP
#define VITB_HIST_FLIP 0
#endif
template <class C>
__host__ __device__ constexpr bool hist_bfly_flip(int i) {
    if (VITB_HIST_FLIP == 1) return (i & 1) != 0;
    if (VITB_HIST_FLIP == 2) {
        const uint32_t p = bfly_pattern<C>(uint32_t(hist_bfly_order<C>(i))), q = (~p) & uint32_t(C::NP - 1);
        return p > q;
    }
    return false;
}
template <class C, int FMT, int TIE_SIMD, int... Js>
__device__ __forceinline__ void hist_bfly_all(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t (&T)[C::NP],
                                              const uint32_t (&TT)[C::NP], const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP],
                                              std::integer_sequence<int, Js...>) {
    (hist_bfly_at<C, FMT, TIE_SIMD, hist_bfly_order<C>(Js), hist_bfly_flip<C>(Js)>(x, y, T, TT, V, VT), ...);
}

// branch metric table in the lane format: T[p] = sum_i (bit_i(p) ? e_high_i : e_low_i)
template <int FMT, int R, int LEVEL>
struct HistTable {
    template <int NP>
    static __device__ __forceinline__ void run(uint32_t (&T)[NP], const uint32_t (&lo)[R], const uint32_t (&hi)[R]) {
        HistTable<FMT, R, LEVEL - 1>::run(T, lo, hi);
        constexpr int H = 1 << (LEVEL - 1);
#pragma unroll
        for (int p = 0; p < H; p++) {
            T[p + H] = HistLane<FMT>::add(T[p], hi[LEVEL - 1]);
            T[p] = HistLane<FMT>::add(T[p], lo[LEVEL - 1]);
        }
    }
};
template <int FMT, int R>
struct HistTable<FMT, R, 1> {
    template <int NP>
    static __device__ __forceinline__ void run(uint32_t (&T)[NP], const uint32_t (&lo)[R], const uint32_t (&hi)[R]) {
        T[0] = lo[0];
        T[1] = hi[0];
    }
};

// one trellis step x -> y; tag = 2^k in the history field(s), k = step index inside the period.
//
// LAZY RENORMALISATION: the rare branch only finds the minimum m of the triggered frame(s) and leaves pend = -m in the metric
// field(s); the NEXT step adds pend to both branch errors of symbol 0, i.e. to every entry of its branch metric table, so its sums
// are (x + T - m) = ((x - m) + T) modulo 2^8 / 2^16 - bit for bit what the reference computes from the renormalised metrics
// (scalar.h:139-153), wrap-around included - without the 2^(K-1) subtractions.  A value still pending after the last step is
// applied when the final metrics are written.  LAZY = false keeps the subtraction inside the branch: measured on the B200, the
// packed two-frame format behind the direct fetch (config 2) loses 7 % with the lazy form - the table of step k+1 then depends on
// the trigger of step k and can no longer be built under step k's butterflies - while the 32-bit format gains 0 - 3 %.
template <int FMT>
struct HistLazyRenorm { static constexpr bool value = (FMT == 1); };

template <class C, int FMT, int TIE_SIMD, bool CONSISTENT, bool NEG_BY_NOT>
__device__ __forceinline__ void hist_step(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t* sym, const HistConsts& c,
                                          const uint32_t tag, uint64_t& accA, uint64_t& accB, uint32_t& pend) {
    using LN = HistLane<FMT>;
    constexpr int R = C::R, NP = C::NP, NS = C::NS;
    uint32_t lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        constexpr bool LAZY = HistLazyRenorm<FMT>::value;
        const uint32_t cl = (LAZY && i == 0) ? LN::add(c.c_low, pend) : c.c_low;
        lo[i] = LN::add(sym[i], cl);               // s - low
        // high - s  (|branch - s| for s in [low, high], scalar.h:30-34, 96-105).  Two equivalent forms; which one ptxas schedules
        // better depends on where the symbols come from (measured on config 2 / 1 / 4: ~s + (c + 1) is 3 % faster behind the direct
        // fetch, where the NOT fuses into the unpacking LOP3; c - s is 5 % faster behind the packed stream)
        if (LAZY && i == 0) hi[i] = NEG_BY_NOT ? LN::add(~sym[i], LN::add(c.c_high_n, pend)) : LN::sub(LN::add(c.c_high, pend), sym[i]);
        else hi[i] = NEG_BY_NOT ? LN::add(~sym[i], c.c_high_n) : LN::sub(c.c_high, sym[i]);
    }
    if constexpr (HistLazyRenorm<FMT>::value) pend = 0u;
    uint32_t T[NP], TT[NP];
    HistTable<FMT, R, R>::run(T, lo, hi);
#pragma unroll
    for (int k = 0; k < NP; k++) TT[k] = LN::add(T[k], tag);
    if constexpr (CONSISTENT) {
        hist_bfly_all<C, FMT, TIE_SIMD>(x, y, T, TT, T, TT, std::make_integer_sequence<int, NS / 2>{});
    } else {
        uint32_t V[NP], VT[NP];       // inverted_error = max_error - total_error (scalar.h:107) = T[~pattern] + c_inv
#pragma unroll
        for (int k = 0; k < NP; k++) { V[k] = LN::add(T[k], c.c_inv); VT[k] = LN::add(TT[k], c.c_inv); }
        hist_bfly_all<C, FMT, TIE_SIMD>(x, y, T, TT, V, VT, std::make_integer_sequence<int, NS / 2>{});
    }
    // renormalisation: "if (new_metric[0] >= renormalisation_threshold)" (scalar.h:48).  The history field below the metric cannot
    // change the outcome: metric >= thr  <=>  y[0] (field) > (thr << HB) - 1  <=>  y[0] (field) >= thr << HB.
    bool any;
    if constexpr (FMT == 0) any = c.always || __vmaxu2(y[0], c.thr_m1) != c.thr_m1;      // either frame (exact test in the rare branch)
    else any = y[0] >= c.thr;
    if (any) {
        uint32_t m = y[0];                                        // scalar.h:140-146: the metric field of the minimum is the minimum metric
#pragma unroll
        for (int q = 1; q + 1 < NS; q += 2) m = LN::min3(m, y[q], y[q + 1]);
        m = LN::min2(m, y[NS - 1]);
        uint32_t sub;
        if constexpr (FMT == 0) {
            bool trigB, trigA;
            (void)__vibmin_u16x2(c.thr, y[0], &trigB, &trigA);     // pred = thr <= y0 per half
            const uint32_t mA = m & 0xff00u, mB = (m >> 16) & 0xff00u;
            sub = (trigA ? mA : 0u) | ((trigB ? mB : 0u) << 16);
            if (trigA) accA += uint64_t(mA >> 8);                   // scalar.h:49, 152
            if (trigB) accB += uint64_t(mB >> 8);
        } else {
            sub = m & 0xffff0000u;
            accA += uint64_t(sub >> 16);
        }
        const uint32_t neg = (FMT == 0) ? __vsub2(0u, sub) : (0u - sub);
        if constexpr (HistLazyRenorm<FMT>::value) {
            pend = neg;                                            // scalar.h:148-150, applied through the next step's table
        } else {
#pragma unroll
            for (int q = 0; q < NS; q++) y[q] = LN::add(y[q], neg);  // scalar.h:148-150
        }
    }
}

constexpr int HIST_WARPS = 4;

// grid = ceil(n_blocks / warps per CTA), at most HIST_WARPS warps per CTA; one warp per block of 64 (FMT 0) / 32 (FMT 1) frames.
// Whole frames only: p.resume must be 0 and p.dec_row0 must be 0 (the streaming API keeps using acs_pair_kernel / acs_group_kernel).
// FMT 1 reads the packed stream in the 16-pairs-per-warp-block layout (ingest with ppw = 16).
// VITB_HIST_MINB (experiment knob): CTAs per SM the register allocation is made for.  Not giving the second argument at all is NOT the
// same as giving 1: ptxas then allocates 137 instead of 150 registers for the config-2 kernel, and that build is 5 % faster on real
// frames (0.637 vs 0.669 ms; profiles/microbench/hist_kernel_variants.cu).
#ifdef VITB_HIST_MINB
#define VITB_HIST_BOUNDS __launch_bounds__(32 * HIST_WARPS, VITB_HIST_MINB)
#else
#define VITB_HIST_MINB 0
#define VITB_HIST_BOUNDS __launch_bounds__(32 * HIST_WARPS)
#endif
template <class C, int FMT, int TIE_SIMD, bool CONSISTENT, bool DIRECT>
__global__ void VITB_HIST_BOUNDS acs_hist_kernel(const AcsParams p) {
    using LN = HistLane<FMT>;
    constexpr int R = C::R, NS = C::NS, NW = HistShape<C>::NW, VW = HistShape<C>::VW, HB = LN::HB, FPT = LN::FPT;
    const uint32_t lane = threadIdx.x & 31, blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (blk >= p.n_blocks) return;
    const size_t fA = size_t(blk) * (32 * FPT) + FPT * lane, fB = fA + 1;      // fB only exists for FMT 0
    const HistConsts c = hist_consts<FMT>(p);

    uint32_t x[NS], y[NS], pend = 0u;
    uint64_t accA = 0, accB = 0;
    {
        const uint32_t s = p.start_state & uint32_t(NS - 1);       // core.h:209-210
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = (uint32_t(q) == s) ? c.init_start : c.init_other;
    }

    const uint32_t n_periods = (p.n_steps + HB - 1) / HB;
    uint32_t* rec = static_cast<uint32_t*>(p.dec) + size_t(blk) * n_periods * (32 * NW) + lane * VW;
    // packed stream: FMT 0 [blk][e][32 pair slots], FMT 1 [blk][e][16 pair slots] (lane l = half l & 1 of slot l >> 1)
    constexpr int SLOTS = (FMT == 0) ? 32 : 16;
    const uint32_t* pk = p.pk + size_t(blk) * p.n_steps * R * SLOTS + (FMT == 0 ? lane : (lane >> 1));
    const bool pk_hi = (lane & 1u) != 0u;

    // ---- symbol supply ------------------------------------------------------------------------------------------------------
    // DIRECT: each lane streams its own rows of the caller's [frame][step][R] soft_t array (the reference's layout, scalar.h:43-46):
    // two int8_t rows (FMT 0) or one int16_t row (FMT 1).  A fetch group is 8 steps = WPG whole words of a row; the words of group g+1
    // are loaded while group g runs (about 2700 clocks ahead, more than a DRAM miss), into registers that the step-pair loop consumes
    // from the bottom and shifts down per iteration (no dynamic register index).  A lane meets a new 32-byte sector of a row every
    // few steps, so nearly every warp-level load has a missing lane: one step pair of lookahead (tried first) left the kernel
    // waiting on memory for a third of its time.
    // Packed stream (ingest output: punctured or unaligned input): words of the next step pair are loaded one iteration ahead and the
    // lines of the group after next are prefetched with one prefetch instruction per group (a load would share a scoreboard with the
    // demand loads and make them wait for its DRAM latency).
    constexpr int NROW = FPT;                                      // rows per lane
    constexpr int WPG = 8 * R * LN::SBY / 4;                       // words per row and fetch group
    constexpr int BYTES = 2 * R * LN::SBY;                         // bytes of a row consumed per iteration (step pair)
    constexpr int PAD = BYTES / 4 + 1;
    uint32_t buf[DIRECT ? NROW : 1][DIRECT ? WPG + PAD : 1], nx[DIRECT ? NROW : 1][DIRECT ? WPG : 1], nxt[DIRECT ? 1 : 2 * R];
    const uint32_t* row[NROW];
    uint32_t maxw[NROW] = {};
    if constexpr (DIRECT) {
        const size_t lastf = size_t(p.n_frames) - 1;
#pragma unroll
        for (int a = 0; a < NROW; a++) {
            const size_t f = fA + a, ld = f < lastf ? f : lastf;                      // padding lanes re-read the last frame
            row[a] = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ld * p.sym_row_bytes);
            maxw[a] = clamp_words_left(p.sym_total_bytes, ld * p.sym_row_bytes);    // loads are clamped to stay inside the array
#pragma unroll
            for (int j = 0; j < WPG; j++) nx[a][j] = __ldg(row[a] + (uint32_t(j) < maxw[a] ? uint32_t(j) : maxw[a]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2 * R; k++) nxt[k] = (uint32_t(k / R) < p.n_steps) ? __ldg(pk + size_t(k) * SLOTS) : 0u;
    }

    // one step pair from the bottom of the group buffer, in the lane format, then shift the buffer down by the bytes consumed
    auto take_pair = [&](uint32_t (&out)[2 * R]) {
#pragma unroll
        for (int e = 0; e < 2 * R; e++) {
            if constexpr (FMT == 0) {
                const int w = e >> 2, o = e & 3;          // byte o of A -> byte 1, byte o of B -> byte 3, history bytes cleared
                out[e] = __byte_perm(buf[0][w], buf[NROW - 1][w], uint32_t(((4 + o) << 12) | (o << 4))) & 0xff00ff00u;
            } else {
                const int w = e >> 1;
                out[e] = (e & 1) ? (buf[0][w] & 0xffff0000u) : (buf[0][w] << 16);
            }
        }
#pragma unroll
        for (int a = 0; a < NROW; a++) {
            if constexpr (BYTES % 4 == 0) {
#pragma unroll
                for (int j = 0; j < WPG; j++) buf[a][j] = buf[a][j + BYTES / 4];
            } else {                                      // 4m + 2 bytes: shift by m words and 16 bits
#pragma unroll
                for (int j = 0; j < WPG; j++) buf[a][j] = __funnelshift_r(buf[a][j + BYTES / 4], buf[a][j + BYTES / 4 + 1], 16);
            }
        }
    };
    auto unpack_pk = [&](uint32_t w) -> uint32_t {        // packed pair word -> lane format
        if constexpr (FMT == 0) return w;
        else return pk_hi ? (w & 0xffff0000u) : (w << 16);
    };

    uint32_t t = 0;        // next step
    for (uint32_t r = 0; r < n_periods; r++) {
        const uint32_t pleft = p.n_steps - t, pst = pleft < uint32_t(HB) ? pleft : uint32_t(HB);     // steps in this period
        uint32_t tag = LN::TAG1;
#pragma unroll 1
        for (uint32_t g0 = 0; g0 < pst; g0 += 8) {          // fetch groups of the period (one for FMT 0, up to two for FMT 1)
            const uint32_t nst = (pst - g0) < 8u ? (pst - g0) : 8u;
            if constexpr (DIRECT) {
#pragma unroll
                for (int a = 0; a < NROW; a++) {
#pragma unroll
                    for (int j = 0; j < WPG; j++) buf[a][j] = nx[a][j];
#pragma unroll
                    for (int j = WPG; j < WPG + PAD; j++) buf[a][j] = 0u;
                }
                if (t + 8 < p.n_steps) {
                    const uint32_t w0 = ((t >> 3) + 1u) * uint32_t(WPG);
#pragma unroll
                    for (int a = 0; a < NROW; a++) {
                        // uint16 format only: in the packed two-frame format the same change costs 9 % (config 2 ACS 0.637 -> 0.693 ms,
                        // fewer instructions but a worse ptxas schedule); the 32-bit format gains 3 % (config 1 0.179 -> 0.174 ms)
                        if (FMT == 1 && w0 + uint32_t(WPG - 1) <= maxw[a]) {  // the whole group lies inside the array: one base, constant offsets
                            const uint32_t* q = row[a] + w0;
#pragma unroll
                            for (int j = 0; j < WPG; j++) nx[a][j] = __ldg(q + j);
                        } else {                                             // last rows of the array: clamp every word
#pragma unroll
                            for (int j = 0; j < WPG; j++) {
                                const uint32_t w = w0 + uint32_t(j);
                                nx[a][j] = __ldg(row[a] + (w < maxw[a] ? w : maxw[a]));
                            }
                        }
                    }
                }
            } else {
                const uint32_t line = (t + 16u) * uint32_t(R) + lane;          // lines of the group after next: 8 * R <= 32 of them
                if (lane < 8u * uint32_t(R) && line < p.n_steps * uint32_t(R))
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(p.pk + (size_t(blk) * p.n_steps * R + line) * SLOTS));
            }
            uint32_t k = 0;
#pragma unroll 1
            for (; k + 2 <= nst; k += 2) {
                uint32_t cur[2 * R];
                if constexpr (DIRECT) {
                    take_pair(cur);
                } else {
#pragma unroll
                    for (int i = 0; i < 2 * R; i++) cur[i] = unpack_pk(nxt[i]);
#pragma unroll
                    for (int i = 0; i < 2 * R; i++) nxt[i] = (t + 2 + uint32_t(i / R) < p.n_steps) ? __ldg(pk + (size_t(t + 2) * R + i) * SLOTS) : 0u;
                }
                hist_step<C, FMT, TIE_SIMD, CONSISTENT, DIRECT>(x, y, &cur[0], c, tag, accA, accB, pend);
                hist_step<C, FMT, TIE_SIMD, CONSISTENT, DIRECT>(y, x, &cur[R], c, tag << 1, accA, accB, pend);
                tag <<= 2;
                t += 2;
            }
            if (k < nst) {          // odd number of steps left (end of the frame): one more step, then back into x
                uint32_t cur[2 * R];
                if constexpr (DIRECT) {
                    take_pair(cur);
                } else {
#pragma unroll
                    for (int i = 0; i < 2 * R; i++) cur[i] = unpack_pk(nxt[i]);
                }
                hist_step<C, FMT, TIE_SIMD, CONSISTENT, DIRECT>(x, y, &cur[0], c, tag, accA, accB, pend);
#pragma unroll
                for (int q = 0; q < NS; q++) x[q] = y[q];
                t += 1;
            }
        }
        // history record of this period, then clear the history fields
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = __byte_perm(x[2 * i], x[2 * i + 1], FMT == 0 ? 0x6420u : 0x5410u);
            if constexpr (TIE_SIMD) {                                  // the tag marked path 0: decision = !tag
                const uint32_t vmask = (FMT == 0 ? 0x01010101u : 0x00010001u) * ((1u << pst) - 1u);
                w[i] = ~w[i] & vmask;
            }
        }
        uint32_t* dst = rec + size_t(r) * (32 * NW);
#pragma unroll
        for (int v = 0; v < NW / VW; v++) {
            if constexpr (VW == 4) *reinterpret_cast<uint4*>(dst + v * 128) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
            else if constexpr (VW == 2) *reinterpret_cast<uint2*>(dst + v * 64) = make_uint2(w[2 * v], w[2 * v + 1]);
            else dst[v * 32] = w[v];
        }
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] &= LN::METRIC_MASK;
    }

    // final metrics in logical order (core.h:195-199 reads old_metrics[end_state]); a renormalisation triggered by the last step is
    // still pending
    if constexpr (HistLazyRenorm<FMT>::value) {
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = LN::add(x[q], pend);
    }
    uint16_t* mA = p.metrics + fA * NS;
    if constexpr (FMT == 0) {
        uint16_t* mB = p.metrics + fB * NS;
#pragma unroll
        for (int q = 0; q < NS; q++) {
            mA[q] = uint16_t((x[q] & 0xffffu) >> 8);
            mB[q] = uint16_t(x[q] >> 24);
        }
        p.acc[fA] = accA;
        p.acc[fB] = accB;
    } else {
#pragma unroll
        for (int q = 0; q < NS; q++) mA[q] = uint16_t(x[q] >> 16);
        p.acc[fA] = accA;
    }
}

}  // namespace vitb
