// acs_hist.cuh -- add-compare-select with SURVIVOR-HISTORY TAGS for uint8_t error metrics (K <= 7), batch decode only.
//
// Replaces ViterbiDecoder_Scalar::update / bfly / renormalise (include/viterbi/viterbi_decoder_scalar.h:28-153) for a batch, like
// acs_pair.cuh, with the same mapping (one thread owns a frame pair, register q = metric of state q of frame A | frame B << 16,
// uint8_t metrics held as metric << 8) but a different way of producing the decisions:
//
//   * the free low byte of every 16-bit half carries the last <= 8 decisions of the SURVIVOR PATH that ends in that state.
//     At step k of an 8-step period the path-1 operand of each compare gets the tag 2^k added (low bytes of the operands hold
//     only bits < k, so the low byte never carries into the metric), and min() then
//       - compares the metrics exactly (high bytes differ -> the low bytes are irrelevant),
//       - breaks a metric tie in favour of path 0, because low(a) < 2^k <= low(b): the reference's strict '>' (scalar.h:123-124),
//       - and keeps the winner's low byte: bit k of the new state's low byte IS the decision, bits < k are the decisions along the
//         winner's own survivor path.
//     So a butterfly is 2 VIADD.16x2 + 2 VIADDMNMX.U16x2 = FOUR instructions for 4 add-compare-selects of 2 frames each: no
//     predicates, no per-step bit collection, no tag clearing (acs_pair.cuh's tagged butterfly needs 8, its predicate form 10).
//     TIE_SIMD tags path 0 instead (a tie selects path 1, x86/viterbi_decoder_avx_u16.h:112-115) and the bits are inverted on output.
//   * every 8 steps the 2^(K-1) low bytes of a frame are written out as one "history record" (2^(K-1) bytes per frame and period -
//     exactly the size of 8 reference decision rows, core.h:49-83) and cleared.  Traceback (traceback_hist_kernel) then moves 8
//     steps per lookup: the byte of the current state holds the 8 decisions of its survivor path, which are the decoded bits
//     AND determine the state 8 steps earlier (core.h:96-113 applied 8 times).
//   * the metrics ping-pong between two register sets in logical state order, so the step code does not depend on the step index:
//     the hot loop is two steps (about 5 KB of SASS) and there is no register renaming.
//   * renormalisation (scalar.h:48, 139-153) works on the high bytes exactly as in acs_pair.cuh; the low bytes ride along.
//
// Record layout: NW = 2^(K-1)/2 32-bit words per thread and period, word w = [A:state 2w, B:state 2w, A:state 2w+1, B:state 2w+1];
// dec = uint32 [n_blocks][n_periods][NW / VW][32 lanes][VW], VW = min(4, NW)  (coalesced 16-byte stores).
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"

namespace vitb {

template <class C>
struct HistShape {
    static constexpr int NS = C::NS;
    static constexpr int NW = NS / 2;                    // record words per thread (two frames) and period
    static constexpr int VW = NW < 4 ? NW : 4;           // words per vector store
    static constexpr int PERIOD = 8;                     // steps per record (bits of history per state)
    static constexpr size_t BLOCK_RECORD_BYTES = size_t(64) * NS;     // one 64-frame block, one period
};

template <class C, bool TIE_SIMD, int J>
__device__ __forceinline__ void hist_bfly_at(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t (&T)[C::NP],
                                             const uint32_t (&TT)[C::NP], const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP]) {
    constexpr int H = C::NS / 2;
    constexpr uint32_t pat = bfly_pattern<C>(uint32_t(J));
    constexpr uint32_t ipat = (~pat) & uint32_t(C::NP - 1);
    if constexpr (!TIE_SIMD) {
        const uint32_t b0 = __vadd2(x[J + H], VT[ipat]);                    // (1|X) + inverted, tagged    scalar.h:114
        const uint32_t b1 = __vadd2(x[J + H], TT[pat]);                     // (1|X) + total,    tagged    scalar.h:116
        y[2 * J] = __viaddmin_u16x2(x[J], T[pat], b0);                      // min((0|X) + total, b0)      scalar.h:113,127
        y[2 * J + 1] = __viaddmin_u16x2(x[J], V[ipat], b1);                 // min((0|X) + inverted, b1)   scalar.h:115,128
    } else {
        const uint32_t a0 = __vadd2(x[J], TT[pat]);                         // tag on path 0: a tie selects path 1
        const uint32_t a1 = __vadd2(x[J], VT[ipat]);
        y[2 * J] = __viaddmin_u16x2(x[J + H], V[ipat], a0);
        y[2 * J + 1] = __viaddmin_u16x2(x[J + H], T[pat], a1);
    }
}

template <class C, bool TIE_SIMD, int... Js>
__device__ __forceinline__ void hist_bfly_all(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t (&T)[C::NP],
                                              const uint32_t (&TT)[C::NP], const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP],
                                              std::integer_sequence<int, Js...>) {
    (hist_bfly_at<C, TIE_SIMD, Js>(x, y, T, TT, V, VT), ...);
}

// one trellis step x -> y; tag2 = 2^k in both halves, k = step index inside the period
template <class C, bool TIE_SIMD, bool CONSISTENT>
__device__ __forceinline__ void hist_step(const uint32_t (&x)[C::NS], uint32_t (&y)[C::NS], const uint32_t* sym, const AcsParams& p,
                                          const uint32_t tag2, const uint32_t thr_m1, const bool always, uint64_t& accA, uint64_t& accB) {
    constexpr int R = C::R, NP = C::NP, NS = C::NS;
    uint32_t lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        lo[i] = __vadd2(sym[i], p.c_low2);
        hi[i] = __vadd2(~sym[i], p.c_high2);
    }
    uint32_t T[NP], TT[NP];
    TableBuild<R, R>::run(T, lo, hi);
#pragma unroll
    for (int k = 0; k < NP; k++) TT[k] = __vadd2(T[k], tag2);
    if constexpr (CONSISTENT) {
        hist_bfly_all<C, TIE_SIMD>(x, y, T, TT, T, TT, std::make_integer_sequence<int, NS / 2>{});
    } else {
        uint32_t V[NP], VT[NP];       // inverted_error = max_error - total_error (scalar.h:107) = T[~pattern] + c_inv
#pragma unroll
        for (int k = 0; k < NP; k++) { V[k] = __vadd2(T[k], p.c_inv2); VT[k] = __vadd2(TT[k], p.c_inv2); }
        hist_bfly_all<C, TIE_SIMD>(x, y, T, TT, V, VT, std::make_integer_sequence<int, NS / 2>{});
    }
    // renormalisation: "if (new_metric[0] >= renormalisation_threshold)" (scalar.h:48).  The history byte below the metric cannot
    // change the outcome: metric >= thr  <=>  y[0] half > (thr << 8) - 1.  One packed max + one compare decide whether EITHER frame
    // triggers (thr == 0 means "always": thr_m1 is then unusable, `always` covers it); the exact per-frame work is in the rare branch.
    if (always || __vmaxu2(y[0], thr_m1) != thr_m1) {
        bool trigB, trigA;
        (void)__vibmin_u16x2(p.thr2, y[0], &trigB, &trigA);     // pred = thr <= y0 per half
        const uint32_t m = packed_min<NS>(y);                   // scalar.h:140-146: the high byte of the 16-bit minimum is the minimum metric
        const uint32_t mA = m & 0xff00u, mB = (m >> 16) & 0xff00u;
        const uint32_t sub = (trigA ? mA : 0u) | ((trigB ? mB : 0u) << 16);
        const uint32_t neg = __vsub2(0u, sub);
#pragma unroll
        for (int q = 0; q < NS; q++) y[q] = __vadd2(y[q], neg);  // scalar.h:148-150
        if (trigA) accA += uint64_t(mA >> 8);                   // scalar.h:49, 152
        if (trigB) accB += uint64_t(mB >> 8);
    }
}

constexpr int HIST_WARPS = 4;

// grid = ceil(n_blocks / HIST_WARPS); one warp per 64-frame block, lane l owns frames 64*blk + 2l and 64*blk + 2l + 1.
// Whole frames only: p.resume must be 0 and p.dec_row0 must be 0 (the streaming API keeps using acs_pair_kernel).
template <class C, bool TIE_SIMD, bool CONSISTENT, bool DIRECT>
__global__ void __launch_bounds__(32 * HIST_WARPS) acs_hist_kernel(const AcsParams p) {
    constexpr int R = C::R, NS = C::NS, NW = HistShape<C>::NW, VW = HistShape<C>::VW;
    const uint32_t lane = threadIdx.x & 31, blk = blockIdx.x * HIST_WARPS + (threadIdx.x >> 5);
    if (blk >= p.n_blocks) return;
    const size_t fA = size_t(blk) * 64 + 2 * lane, fB = fA + 1;

    uint32_t x[NS], y[NS];
    uint64_t accA = 0, accB = 0;
    {
        const uint32_t s = p.start_state & uint32_t(NS - 1);       // core.h:209-210
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = (uint32_t(q) == s) ? p.init_start2 : p.init_other2;
    }

    // either half of the threshold zero: every step renormalises (thr - 1 would wrap)
    const bool always = (p.thr2 & 0xffffu) == 0u || (p.thr2 >> 16) == 0u;
    const uint32_t thr_m1 = always ? 0xffffffffu : p.thr2 - 0x00010001u;

    const uint32_t n_periods = (p.n_steps + 7) / 8;
    uint32_t* rec = static_cast<uint32_t*>(p.dec) + size_t(blk) * n_periods * (32 * NW) + lane * VW;
    const uint32_t* pk = p.pk + size_t(blk) * p.n_steps * R * 32 + lane;

    // ---- symbol supply ------------------------------------------------------------------------------------------------------
    // DIRECT: each lane streams its own two rows of the caller's [frame][step][R] int8_t array (the reference's layout, scalar.h:43-46).
    // A period is 8 steps = 8R bytes = WPP whole words of a row; the words of period r+1 are loaded while period r runs (a full
    // period ahead: about 2700 clocks, more than a DRAM miss), into registers that the step-pair loop consumes from the bottom and
    // shifts down by 2R bytes per iteration (no dynamic register index).  A lane meets a new 32-byte sector of each row every 32/R
    // steps, so nearly every warp-level load has a missing lane: one step pair of lookahead (tried first) left the kernel waiting
    // on memory for a third of its time.
    // Packed stream (ingest output: punctured or unaligned input): words of the next step pair are loaded one iteration ahead and the
    // lines of the period after next are prefetched with one prefetch instruction per period (a load would share a scoreboard with the
    // demand loads and make them wait for its DRAM latency).
    constexpr int WPP = 2 * R;                               // words per row and period
    uint32_t bufA[DIRECT ? WPP + 2 : 1], bufB[DIRECT ? WPP + 2 : 1], nxA[DIRECT ? WPP : 1], nxB[DIRECT ? WPP : 1], nxt[DIRECT ? 1 : 2 * R];
    const uint32_t* rowA = nullptr;
    const uint32_t* rowB = nullptr;
    uint32_t maxwA = 0, maxwB = 0;
    if constexpr (DIRECT) {
        const size_t lastf = size_t(p.n_frames) - 1;
        const size_t ldA = fA < lastf ? fA : lastf, ldB = fB < lastf ? fB : lastf;     // padding lanes re-read the last frame
        rowA = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ldA * p.sym_row_bytes);
        rowB = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ldB * p.sym_row_bytes);
        maxwA = uint32_t((p.sym_total_bytes - ldA * p.sym_row_bytes - 4) >> 2);       // loads are clamped to stay inside the array
        maxwB = uint32_t((p.sym_total_bytes - ldB * p.sym_row_bytes - 4) >> 2);
#pragma unroll
        for (int j = 0; j < WPP; j++) {
            nxA[j] = __ldg(rowA + (uint32_t(j) < maxwA ? uint32_t(j) : maxwA));
            nxB[j] = __ldg(rowB + (uint32_t(j) < maxwB ? uint32_t(j) : maxwB));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 2 * R; k++) nxt[k] = (uint32_t(k / R) < p.n_steps) ? __ldg(pk + size_t(k) * 32) : 0u;
    }

    // one step pair from the bottom of the period buffer: out[n * R + i] = (sA << 8) & 0xffff | (sB << 8) << 16, then shift down
    auto take_pair = [&](uint32_t (&out)[2 * R]) {
#pragma unroll
        for (int n = 0; n < 2; n++) {
#pragma unroll
            for (int i = 0; i < R; i++) {
                const int bp = n * R + i, w = bp >> 2, o = bp & 3;
                out[n * R + i] = __byte_perm(bufA[w], bufB[w], uint32_t(((4 + o) << 12) | (o << 4))) & 0xff00ff00u;
            }
        }
        if constexpr ((2 * R) % 4 == 0) {
#pragma unroll
            for (int j = 0; j < WPP; j++) { bufA[j] = bufA[j + R / 2]; bufB[j] = bufB[j + R / 2]; }
        } else {                                  // 2R = 4m + 2 bytes: shift by m words and 16 bits
#pragma unroll
            for (int j = 0; j < WPP; j++) {
                bufA[j] = __funnelshift_r(bufA[j + R / 2], bufA[j + R / 2 + 1], 16);
                bufB[j] = __funnelshift_r(bufB[j + R / 2], bufB[j + R / 2 + 1], 16);
            }
        }
    };
    static_assert((R % 2 == 0) ? (R / 2 <= 2) : (R / 2 <= 1), "period buffer padding covers R <= 3 (odd) / R <= 4 (even)");

    uint32_t t = 0;        // next step
    for (uint32_t r = 0; r < n_periods; r++) {
        const uint32_t left = p.n_steps - t, nst = left < 8u ? left : 8u;
        if constexpr (DIRECT) {
#pragma unroll
            for (int j = 0; j < WPP; j++) { bufA[j] = nxA[j]; bufB[j] = nxB[j]; }
            bufA[WPP] = bufA[WPP + 1] = bufB[WPP] = bufB[WPP + 1] = 0u;
            if (r + 1 < n_periods) {
                const uint32_t w0 = (r + 1) * uint32_t(WPP);
#pragma unroll
                for (int j = 0; j < WPP; j++) {
                    const uint32_t w = w0 + uint32_t(j);
                    nxA[j] = __ldg(rowA + (w < maxwA ? w : maxwA));
                    nxB[j] = __ldg(rowB + (w < maxwB ? w : maxwB));
                }
            }
        } else {
            const uint32_t line = (t + 16u) * uint32_t(R) + lane;          // lines of the period after next: 8 * R <= 32 of them
            if (lane < 8u * uint32_t(R) && line < p.n_steps * uint32_t(R)) asm volatile("prefetch.global.L1 [%0];" :: "l"(pk + size_t(line) * 32 - lane));
        }
        uint32_t tag2 = 0x00010001u;
        uint32_t k = 0;
#pragma unroll 1
        for (; k + 2 <= nst; k += 2) {
            uint32_t cur[2 * R];
            if constexpr (DIRECT) {
                take_pair(cur);
            } else {
#pragma unroll
                for (int i = 0; i < 2 * R; i++) cur[i] = nxt[i];
#pragma unroll
                for (int i = 0; i < 2 * R; i++) nxt[i] = (t + 2 + uint32_t(i / R) < p.n_steps) ? __ldg(pk + (size_t(t + 2) * R + i) * 32) : 0u;
            }
            hist_step<C, TIE_SIMD, CONSISTENT>(x, y, &cur[0], p, tag2, thr_m1, always, accA, accB);
            hist_step<C, TIE_SIMD, CONSISTENT>(y, x, &cur[R], p, tag2 << 1, thr_m1, always, accA, accB);
            tag2 <<= 2;
            t += 2;
        }
        if (k < nst) {          // odd number of steps in the (last) period: one more step, then back into x
            uint32_t cur[2 * R];
            if constexpr (DIRECT) {
                take_pair(cur);
            } else {
#pragma unroll
                for (int i = 0; i < 2 * R; i++) cur[i] = nxt[i];
            }
            hist_step<C, TIE_SIMD, CONSISTENT>(x, y, &cur[0], p, tag2, thr_m1, always, accA, accB);
#pragma unroll
            for (int q = 0; q < NS; q++) x[q] = y[q];
            t += 1;
        }
        // history record of this period, then clear the low bytes
        const uint32_t vmask = 0x01010101u * ((1u << nst) - 1u);
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x6420);
            if constexpr (TIE_SIMD) w[i] = ~w[i] & vmask;           // the tag marked path 0: decision = !tag
        }
        uint32_t* dst = rec + size_t(r) * (32 * NW);
#pragma unroll
        for (int v = 0; v < NW / VW; v++) {
            if constexpr (VW == 4) *reinterpret_cast<uint4*>(dst + v * 128) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
            else if constexpr (VW == 2) *reinterpret_cast<uint2*>(dst + v * 64) = make_uint2(w[2 * v], w[2 * v + 1]);
            else dst[v * 32] = w[v];
        }
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] &= 0xff00ff00u;
    }

    // final metrics in logical order (core.h:195-199 reads old_metrics[end_state])
    uint16_t* mA = p.metrics + fA * NS;
    uint16_t* mB = p.metrics + fB * NS;
#pragma unroll
    for (int q = 0; q < NS; q++) {
        mA[q] = uint16_t((x[q] & 0xffffu) >> 8);
        mB[q] = uint16_t(x[q] >> 24);
    }
    p.acc[fA] = accA;
    p.acc[fB] = accB;
}

}  // namespace vitb
