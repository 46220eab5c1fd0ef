// acs_hist_group.cuh -- survivor-history add-compare-select with uint16_t error metrics, ONE FRAME OVER T = 4 LANES: K = 9 (the
// default for batches that fill the GPU) and K = 7 (small batches: a quarter of the registers per lane, four times the warps).
//
// The survivor-history idea of acs_hist.cuh (register = metric << 16 | last <= 16 decisions of the survivor path into that state;
// a butterfly is 2 adds + 2 fused add-min, no predicates, one record per 16 steps instead of a decision row per step) needs the
// 32-bit lane format for uint16_t metrics, i.e. one frame per register - and K = 9 has 256 states, too many for one thread.  So the
// states of a frame are divided among T = 4 lanes exactly like acs_group.cuh divides a frame pair:
//     position PHI = (q << 2) | t      q = register inside the lane (64 registers), t = lane inside the group
//     after n steps since the last exchange, logical state s sits at PHI = rotr^n(s)   (8-bit rotation)
// For n = 0 .. 5 the butterfly partner differs in a register bit (in-place butterflies, no communication); after 6 steps one exchange
// through shared memory rotates the positions back to PHI = s.  The branch pattern of a butterfly is (compile-time register part) XOR
// (per-lane part, one value per phase); the lane part is folded into the branch metric table by swapping e_low / e_high.
//
// Instruction budget per lane and step: 32 butterflies x 4 + table 16 + exchange 128 / 6 + record 104 / 16 + ~15 = ~185 for 64
// add-compare-selects of one frame = 2.9 per ACS, against 4.1 for acs_group_kernel<T16> (10 per butterfly of two frames plus its
// overheads); and traceback reads one halfword per 16 steps instead of streaming every decision row.
//
// Records are written in POSITION order (lane t, register q), same layout as acs_hist.cuh with 32 words per lane:
//     dec = uint32 [n_blocks][n_periods][8][32 lanes][4],  word w of a lane = halfwords [register 2w, register 2w + 1];
// a block is one warp = 8 frames.  traceback_hist_kernel finds state s of record r at PHI = rotr^m(s), m = (steps done at the end
// of the record) mod 6.  Whole frames only, symbols fetched directly from 4-byte aligned unpunctured rows (R even).
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"
#include "acs_hist.cuh"

namespace vitb {

template <class C>
struct HistGroupShape {
    static constexpr int LOGT = 2, T = 4, SB = C::SB, LB = SB - LOGT, NL = 1 << LB, NW = NL / 2, FPW = 32 / T, WARPS = 2;
    // K = 9: 64 registers per lane, exchange every 6 steps; K = 7 (small batches, see below): 16 registers, every 4 steps
    static_assert((SB == 8 || SB == 6) && ((LB * C::R) % 2) == 0, "built for K = 9 and K = 7; an exchange period must be whole 32-bit words of symbols");
    // can a 16-step record end behind phase PH of an exchange period (other than its last one)?
    static __host__ __device__ constexpr bool record_may_end_after(int PH) {
        int g = LB, b = 16;
        while (b) { const int r = g % b; g = b; b = r; }          // gcd(LB, 16)
        return PH != LB - 1 && ((PH + 1) % g) == 0;
    }
    // shared-memory word of position PHI' for frame fw of the warp: rows of 32 words = 8 groups of 4; position (qp, tp) sits in row
    // qp, group (fw + 2 * (qp >> (LB - 2))) mod 8, word tp.  A lane writes (t << LB) | q for consecutive q: four consecutive q are one
    // aligned group (STS.128), and the 8 lanes of a quarter warp - two frames x four t - hit 8 different groups; its reads
    // ((q << LOGT) | t, fixed q) touch 32 distinct banks per warp instruction.
    static __host__ __device__ constexpr uint32_t slot(uint32_t fw, uint32_t phi) {
        const uint32_t qp = phi >> LOGT, tp = phi & uint32_t(T - 1);
        return qp * 32u + ((fw + 2u * (qp >> (LB - LOGT))) & 7u) * 4u + tp;
    }
};

template <class F, int... PHs>
__device__ __forceinline__ void hg_for_each_phase(F&& f, std::integer_sequence<int, PHs...>) {
    (f(std::integral_constant<int, PHs>{}), ...);
}

// one in-place butterfly at compile-time phase PH on registers Q and Q | bit
template <class C, int PH, int TIE_SIMD, int Q>
__device__ __forceinline__ void hg_bfly_at(uint32_t (&x)[HistGroupShape<C>::NL], const uint32_t (&T)[C::NP], const uint32_t (&TT)[C::NP],
                                           const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP]) {
    using S = HistGroupShape<C>;
    constexpr int bit = 1 << (S::LB - 1 - PH);
    if constexpr ((Q & bit) == 0) {
        constexpr int q0 = Q, q1 = Q | bit;
        constexpr uint32_t jq = rotl_bits(uint32_t(q0) << S::LOGT, PH, S::SB);   // register-bit part of the old state index
        constexpr uint32_t pat = bfly_pattern<C>(jq), ipat = (~pat) & uint32_t(C::NP - 1);
        uint32_t m0, m1;
        if constexpr (!TIE_SIMD) {
            const uint32_t b0 = x[q1] + VT[ipat], b1 = x[q1] + TT[pat];        // path 1 tagged                 scalar.h:114,116
            m0 = __viaddmin_u32(x[q0], T[pat], b0);                             // new state 2j   stays at q0   scalar.h:113,127
            m1 = __viaddmin_u32(x[q0], V[ipat], b1);                            // new state 2j+1 goes to q1    scalar.h:115,128
        } else {
            const uint32_t a0 = x[q0] + TT[pat], a1 = x[q0] + VT[ipat];         // tag on path 0: a tie selects path 1
            m0 = __viaddmin_u32(x[q1], V[ipat], a0);
            m1 = __viaddmin_u32(x[q1], T[pat], a1);
        }
        x[q0] = m0;
        x[q1] = m1;
    }
}

template <class C, int PH, int TIE_SIMD, int... Qs>
__device__ __forceinline__ void hg_bfly_all(uint32_t (&x)[HistGroupShape<C>::NL], const uint32_t (&T)[C::NP], const uint32_t (&TT)[C::NP],
                                            const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP], std::integer_sequence<int, Qs...>) {
    (hg_bfly_at<C, PH, TIE_SIMD, Qs>(x, T, TT, V, VT), ...);
}

// one trellis step at phase PH.  fold[i] != 0: the lane part of the branch pattern flips symbol i (e_low and e_high trade places).
//
// LAZY RENORMALISATION.  The reference subtracts the minimum from every metric right after the step that triggered
// (scalar.h:139-153).  Here the rare branch only FINDS the minimum m and leaves `pend = -m` (metric field); the next step adds pend
// to the two branch errors of symbol 0, i.e. to every entry of its branch metric table, so its sums are (x + T - m) mod 2^16 =
// ((x - m) + T) mod 2^16: bit for bit what the reference computes from the renormalised metrics, wrap-around included.  The
// registers x are not written inside the branch, which (a) saves the 64 subtractions and (b) removes the control-flow merge that
// made ptxas shuffle half of the in-place butterfly results back into canonical registers after every step (34 IMAD.MOV per step
// in the round-1 build, profiles/r02_ncu_full_acs_hist_group_cfg3.txt).  A pending value left after the last step is applied when
// the final metrics are written.
template <class C, int PH, int TIE_SIMD, bool CONSISTENT>
__device__ __forceinline__ void hg_step(uint32_t (&x)[HistGroupShape<C>::NL], const uint32_t* sym, const uint32_t fold_bits, const HistConsts& c,
                                        const uint32_t tag, const uint32_t lane, uint64_t& acc, uint32_t& pend) {
    using S = HistGroupShape<C>;
    constexpr int R = C::R, NP = C::NP, NL = S::NL;
    uint32_t lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        uint32_t l = sym[i] + c.c_low, h = c.c_high - sym[i];              // s - low, high - s  (scalar.h:96-105 for s in [low, high])
        if (i == 0) { l += pend; h += pend; }                              // pending renormalisation: every table entry holds one of them
        const bool f = (fold_bits >> (PH * R + i)) & 1u;
        lo[i] = f ? h : l;
        hi[i] = f ? l : h;
    }
    pend = 0u;
    uint32_t T[NP], TT[NP];
    HistTable<1, R, R>::run(T, lo, hi);
#pragma unroll
    for (int k = 0; k < NP; k++) TT[k] = T[k] + tag;
    if constexpr (CONSISTENT) {
        hg_bfly_all<C, PH, TIE_SIMD>(x, T, TT, T, TT, std::make_integer_sequence<int, NL>{});
    } else {
        uint32_t V[NP], VT[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) { V[k] = T[k] + c.c_inv; VT[k] = TT[k] + c.c_inv; }
        hg_bfly_all<C, PH, TIE_SIMD>(x, T, TT, V, VT, std::make_integer_sequence<int, NL>{});
    }
    // renormalisation (scalar.h:48, 139-153): state 0 sits in register 0 of lane 0 of the group in every phase.  Only those lanes
    // vote (no shuffle on the per-step critical path); the branch is taken by the whole warp when any of its 8 frames triggers
    // (about every 12th step), so the shuffles inside are convergent.
    const bool leader_trig = ((lane & uint32_t(S::T - 1)) == 0u) && (x[0] >= c.thr);
    if (__any_sync(0xffffffffu, leader_trig)) {
        const bool trig = __shfl_sync(0xffffffffu, leader_trig ? 1u : 0u, int(lane & ~uint32_t(S::T - 1))) != 0u;
        uint32_t m = x[0];
#pragma unroll
        for (int q = 1; q + 1 < NL; q += 2) m = min(min(m, x[q]), x[q + 1]);
        m = min(m, x[NL - 1]);
        m = min(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = min(m, __shfl_xor_sync(0xffffffffu, m, 2));
        const uint32_t sub = trig ? (m & 0xffff0000u) : 0u;
        pend = 0u - sub;                                                    // applied by the next step's table (or the final write-back)
        acc += uint64_t(sub >> 16);
    }
}

// grid = ceil(n_blocks / WARPS); one warp = 8 frames, lane = 4 * frame + t
template <class C, int TIE_SIMD, bool CONSISTENT>
__global__ void __launch_bounds__(32 * HistGroupShape<C>::WARPS, 7) acs_hist_group_kernel(const AcsParams p) {
    using S = HistGroupShape<C>;
    constexpr int R = C::R, NL = S::NL, NW = S::NW, LB = S::LB, LOGT = S::LOGT, T = S::T, SB = S::SB, HB = 16;
    __shared__ __align__(16) uint32_t xch[S::WARPS][NL * 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, blk = blockIdx.x * (blockDim.x >> 5) + warp;
    if (blk >= p.n_blocks) return;
    const uint32_t fw = lane >> LOGT, t = lane & uint32_t(T - 1);
    const size_t f = size_t(blk) * S::FPW + fw;
    const HistConsts c = hist_consts<1>(p);
    uint32_t* my_xch = xch[warp];

    // lane part of the branch pattern, per phase and symbol: bit (n * R + i)
    uint32_t fold_bits = 0;
#pragma unroll
    for (int n = 0; n < LB; n++) fold_bits |= bfly_pattern_dyn<C>(rotl_bits(t, n, SB)) << (n * R);

    uint32_t x[NL];
    uint64_t acc = 0;
    {
        const uint32_t s0 = p.start_state & uint32_t(C::NS - 1);       // core.h:209-210; phase 0: position = state
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = (((uint32_t(q) << LOGT) | t) == s0) ? c.init_start : c.init_other;
    }

    const uint32_t n_periods = (p.n_steps + HB - 1) / HB;
    uint32_t* rec = static_cast<uint32_t*>(p.dec) + size_t(blk) * n_periods * (32 * NW) + lane * 4;

    // symbols: the 4 lanes of a frame read the same row (L1 broadcast).  One exchange period = 6 steps = WPP whole words of the row
    // (R is even), loaded a period ahead; inside the period every index is a compile-time constant.
    constexpr int WPP = LB * R / 2;                               // words per period
    uint32_t cur[WPP], nxt[WPP];
    const size_t lastf = size_t(p.n_frames) - 1, ld = f < lastf ? f : lastf;      // padding frames re-read the last frame
    const uint32_t* row = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ld * p.sym_row_bytes);
    const uint32_t maxw = clamp_words_left(p.sym_total_bytes, ld * p.sym_row_bytes);     // loads are clamped to stay inside the array
#pragma unroll
    for (int j = 0; j < WPP; j++) nxt[j] = __ldg(row + (uint32_t(j) < maxw ? uint32_t(j) : maxw));

    // exchange after LB phases: the value at (q, t) moves to PHI' = (t << LB) | q, read back as PHI' = (q << LOGT) | t.  With the
    // layout of HistGroupShape::slot the write side is one base address per lane and 16-byte stores of four registers, the read side
    // four base addresses per lane plus compile-time offsets:
    //   write (q, t): row (t << (LB-2)) | (q >> 2), group (fw + 2 t) & 7, word q & 3      read q: row q, group (fw + 2 (q >> (LB-2))) & 7, word t
    uint32_t* wr_base = my_xch + (t << (LB - LOGT)) * 32u + ((fw + 2u * t) & 7u) * 4u;
    const uint32_t* rd_base[4];
#pragma unroll
    for (int k = 0; k < 4; k++) rd_base[k] = my_xch + ((fw + 2u * uint32_t(k)) & 7u) * 4u + t;
    auto exchange = [&]() {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NL; q += 4) *reinterpret_cast<uint4*>(wr_base + (q >> 2) * 32) = make_uint4(x[q], x[q + 1], x[q + 2], x[q + 3]);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = rd_base[(q >> (LB - LOGT)) & 3][q * 32];
    };

    uint32_t tag = 1u, pst = 0, r = 0, pend = 0u;
    // history record of the period that just ended (position order), then clear the history fields
    auto emit_record = [&]() {
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x5410u);
            if constexpr (TIE_SIMD) w[i] = ~w[i] & (0x00010001u * ((1u << pst) - 1u));      // the tag marked path 0: decision = !tag
        }
        uint32_t* dst = rec + size_t(r) * (32 * NW);
#pragma unroll
        for (int v = 0; v < NW / 4; v++) *reinterpret_cast<uint4*>(dst + v * 128) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] &= 0xffff0000u;
        r++;
        tag = 1u;
        pst = 0;
    };
    // one step at compile-time phase PH of the period that starts at step n0; records fill up (16 steps) only after odd phases.
    // GUARD = false: the period is known to be complete (main loop: no per-step bounds test, no control-flow merge per step)
    auto do_phase = [&](auto ph_tag, auto guard_tag, uint32_t n0) {
        constexpr int PH = decltype(ph_tag)::value;
        constexpr bool GUARD = decltype(guard_tag)::value;
        if constexpr (GUARD) { if (n0 + uint32_t(PH) >= p.n_steps) return; }
        uint32_t sym[R];
#pragma unroll
        for (int i = 0; i < R; i++) {
            const int e = PH * R + i;
            sym[i] = (e & 1) ? (cur[e >> 1] & 0xffff0000u) : (cur[e >> 1] << 16);
        }
        hg_step<C, PH, TIE_SIMD, CONSISTENT>(x, sym, fold_bits, c, tag, lane, acc, pend);
        tag <<= 1;
        pst++;
        if constexpr (S::record_may_end_after(PH)) {     // (after the last phase the record is cut behind the exchange, see the loop)
            if (pst == uint32_t(HB)) emit_record();
        }
    };

    using NoGuard = std::false_type;
    using Guard = std::true_type;
    auto all_phases = [&](auto guard_tag, uint32_t n0_, auto seq) {
        hg_for_each_phase([&](auto phc) { do_phase(phc, guard_tag, n0_); }, seq);
    };
    uint32_t n0 = 0, w0 = uint32_t(WPP);
#pragma unroll 1
    for (; n0 + LB <= p.n_steps; n0 += LB, w0 += uint32_t(WPP)) {      // complete exchange periods
#pragma unroll
        for (int j = 0; j < WPP; j++) cur[j] = nxt[j];
#pragma unroll
        for (int j = 0; j < WPP; j++) {                                  // words of the next period (clamped: the last period reads ahead)
            const uint32_t w = w0 + uint32_t(j);
            nxt[j] = __ldg(row + (w < maxw ? w : maxw));
        }
        all_phases(NoGuard{}, n0, std::make_integer_sequence<int, LB>{});
        exchange();                                 // a full period ran: positions back to PHI = s
        if (pst == uint32_t(HB)) emit_record();     // records are cut in the layout the traceback expects: PHI = rotr^(steps mod 6)(s)
    }
    if (n0 < p.n_steps) {                           // the incomplete last period: guarded steps, no exchange
#pragma unroll
        for (int j = 0; j < WPP; j++) cur[j] = nxt[j];
        all_phases(Guard{}, n0, std::make_integer_sequence<int, LB - 1>{});
    }
    if (pst) emit_record();                         // last, partial record (its pst is still needed for the SIMD tie-break mask)
    const uint32_t ph = p.n_steps % uint32_t(LB);
#pragma unroll
    for (int q = 0; q < NL; q++) x[q] += pend;      // renormalisation still pending from the last step (metric field only)

    // final metrics in logical order: state s sits at PHI = rotr^ph(s)   (core.h:195-199 reads old_metrics[end_state])
    uint16_t* m = p.metrics + f * C::NS;
    hg_for_each_phase([&](auto phc) {
        constexpr int PH = decltype(phc)::value;
        if (ph == uint32_t(PH)) {
#pragma unroll
            for (int q = 0; q < NL; q++) m[rotl_bits((uint32_t(q) << LOGT) | t, PH, SB)] = uint16_t(x[q] >> 16);
        }
    }, std::make_integer_sequence<int, LB>{});
    if (t == 0) p.acc[f] = acc;
}

}  // namespace vitb
