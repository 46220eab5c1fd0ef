// acs_hist_cta.cuh -- survivor-history add-compare-select for K = 15 (Cassini: 16384 states) with uint16_t error metrics:
// ONE FRAME PER CTA of 512 threads x 32 registers.
//
// Replaces ViterbiDecoder_Scalar::update / bfly / renormalise (include/viterbi/viterbi_decoder_scalar.h:28-153) for whole-frame
// batch calls, like acs_cta.cuh, but with the survivor-history formulation of acs_hist.cuh / acs_hist_group.cuh:
//     register = metric << 16 | last <= 16 decisions of the survivor path into that state
//     butterfly = 1 LDS.64 (branch metric table entry {total, inverted}) + 2 three-input adds (x + metric + tag) + 2 fused add-min
// instead of acs_cta.cuh's 4 packed adds + 2 packed min + 4 predicated FADDs + decision-byte assembly + a 4 KB decision row per
// step.  One frame per register (no packing: uint16_t metrics leave no spare bits in a 16-bit half), so a CTA holds ONE frame:
//   * 1024 config-5 frames are 1024 CTAs = 6.92 rounds on 148 SMs (acs_cta.cuh: 512 CTAs = 3.46 rounds, the last one half empty),
//     and a 128-frame shard of an 8-GPU run still fills 128 SMs (acs_cta.cuh: 64);
//   * traceback reads one halfword per 16 steps and frame instead of one word per step.
//
// Position algebra of acs_cta.cuh with LOGT = 9: PHI = (q << 9) | t; n steps after an exchange state s sits at PHI = rotr^n(s);
// LB = 5 register-only steps, then the CTA exchanges through 64 KB of shared memory (skewed, conflict-free) under two
// __syncthreads.  The branch metric table of a whole group ({total, inverted} x 64 patterns x 5 steps, metric field format) is built
// cooperatively from the caller's int16_t row (read where it lies: unpunctured input only, any alignment) at each exchange.
//
// Renormalisation (scalar.h:48, 139-153) is the speculation / rollback / replay scheme of acs_cta.cuh: a group runs without it
// while thread 0 (state 0 sits in its register 0 in every phase) tracks the trigger; if it fired, the CTA restores the metrics AND
// the record bookkeeping of the group's start from the exchange buffer and replays the group one step at a time with a barrier
// after every step, subtracting the CTA-wide minimum at exactly the reference's step.
//
// Records (position order, like acs_hist_group.cuh): every 16 steps each thread writes its 32 history halfwords as 16 words,
//     dec = uint32 [frame][period][4][512 threads][4],  word w of a thread = halfwords [register 2w, register 2w + 1];
// traceback_hist_kernel finds state s of record r at PHI = rotr^m(s), m = (steps done at the end of the record) mod 5.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"
#include "acs_hist.cuh"
#include "acs_cta.cuh"

namespace vitb {

template <class C>
struct HistCtaShape {
    using S = CtaShape<C, 9>;
    static constexpr int LOGT = 9, T = 512, SB = C::SB, LB = S::LB, NL = S::NL, NW = NL / 2, NP = C::NP, HB = 16, WARPS = T / 32;
    static_assert(SB == 14 && LB == 5 && NL == 32, "built for K = 15");
    static constexpr size_t XCH_WORDS = S::XCH_WORDS, TBL_WORDS = S::TBL_WORDS;
    static constexpr size_t SMEM_BYTES = (XCH_WORDS + TBL_WORDS + 64) * 4;
};

__host__ __device__ constexpr uint32_t rotl_rt(uint32_t v, uint32_t r, uint32_t n) {
    return r == 0 ? v : (((v << r) | (v >> (n - r))) & ((1u << n) - 1u));
}

template <class C, int PH, int TIE_SIMD, int Q>
__device__ __forceinline__ void hc_bfly_at(uint32_t (&x)[HistCtaShape<C>::NL], const uint2* tbl_ph, const uint32_t pt, const uint32_t tag,
                                           const uint32_t one) {
    using H = HistCtaShape<C>;
    constexpr int bit = 1 << (H::LB - 1 - PH);
    if constexpr ((Q & bit) == 0) {
        constexpr int q0 = Q, q1 = Q | bit;
        constexpr uint32_t jq = rotl_bits(uint32_t(q0) << H::LOGT, PH, H::SB);   // register-bit part of the old state index
        constexpr uint32_t pq = bfly_pattern<C>(jq);
        const uint2 e = tbl_ph[pq ^ pt];                // {total_error, inverted_error} << 16 of pattern pq ^ pt (scalar.h:66-73, 107)
        uint32_t m0, m1;
        // The tag is added to the path-1 register ONCE (x1t), with a multiply-add by a run-time 1 so that ptxas cannot fold it back
        // into two three-input IADD3: IADD3 issues on the same pipe as VIADDMNMX and LOP3 (6 of 6 instructions per butterfly on one
        // pipe: the kernel ran at 1360 clocks per step); IMAD and the two-input adds (IMAD.IADD) go to the other pipe.
        if constexpr (!TIE_SIMD) {
            const uint32_t x1t = tag * one + x[q1];                             // path 1 tagged
            const uint32_t b0 = x1t + e.y, b1 = x1t + e.x;                      //                               scalar.h:114,116
            m0 = __viaddmin_u32(x[q0], e.x, b0);                                // new state 2j   stays at q0   scalar.h:113,127
            m1 = __viaddmin_u32(x[q0], e.y, b1);                                // new state 2j+1 goes to q1    scalar.h:115,128
        } else {
            const uint32_t x0t = tag * one + x[q0];                             // tag on path 0: a tie selects path 1
            const uint32_t a0 = x0t + e.x, a1 = x0t + e.y;
            m0 = __viaddmin_u32(x[q1], e.y, a0);
            m1 = __viaddmin_u32(x[q1], e.x, a1);
        }
        x[q0] = m0;
        x[q1] = m1;
    }
}

template <class C, int PH, int TIE_SIMD, int... Qs>
__device__ __forceinline__ void hc_bfly_all(uint32_t (&x)[HistCtaShape<C>::NL], const uint2* tbl_ph, const uint32_t pt, const uint32_t tag,
                                            const uint32_t one, std::integer_sequence<int, Qs...>) {
    (hc_bfly_at<C, PH, TIE_SIMD, Qs>(x, tbl_ph, pt, tag, one), ...);
}

// grid = number of frames, block = 512, dynamic shared memory = HistCtaShape::SMEM_BYTES.  Whole frames only (no resume).
// MINB: CTAs per SM the register allocation is made for (1: 128 registers per thread; 2: 64 registers - measured slower, it spills)
template <class C, int TIE_SIMD, int MINB>
__global__ void __launch_bounds__(HistCtaShape<C>::T, MINB) acs_hist_cta_kernel(const AcsParams p) {
    using H = HistCtaShape<C>;
    using S = typename H::S;
    constexpr int LB = H::LB, NL = H::NL, NW = H::NW, R = C::R, NP = H::NP, SB = H::SB, LOGT = H::LOGT, HB = H::HB;
    extern __shared__ uint32_t smem[];
    uint32_t* xch = smem;                                             // [NS] exchange buffer = registers at the start of the group
    uint2* tbl = reinterpret_cast<uint2*>(smem + H::XCH_WORDS);       // [LB][NP] {total, inverted}
    uint32_t* red = smem + H::XCH_WORDS + H::TBL_WORDS;               // [WARPS] reduction scratch
    uint32_t* flag = red + H::WARPS;                                  // [2] trigger flags

    const uint32_t t = threadIdx.x;
    const size_t f = blockIdx.x;
    const HistConsts c = hist_consts<1>(p);
    const uint32_t one = p.resume + 1u;                               // whole frames only: resume is 0.  A 1 the compiler cannot see (hc_bfly_at)

    uint32_t pt[LB];                                                  // thread part of the branch pattern per phase
#pragma unroll
    for (int n = 0; n < LB; n++) pt[n] = bfly_pattern_dyn<C>(rotl_bits(t, n, SB));

    uint32_t x[NL];
    uint64_t acc = 0;
    {
        const uint32_t s0 = p.start_state & uint32_t(C::NS - 1);      // core.h:209-210; phase 0: position = state
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = (((uint32_t(q) << LOGT) | t) == s0) ? c.init_start : c.init_other;
    }

    const uint32_t n_periods = (p.n_steps + HB - 1) / HB;
    uint32_t* rec = static_cast<uint32_t*>(p.dec) + f * size_t(n_periods) * (size_t(C::NS) / 2) + t * 4;
    const int16_t* row = reinterpret_cast<const int16_t*>(static_cast<const uint8_t*>(p.sym) + f * p.sym_row_bytes);

    // branch metric tables of the n steps of the group that starts at first_step, one entry per thread (320 of the 512)
    auto build_tables = [&](uint32_t first_step, uint32_t n) {
        if (t < uint32_t(LB * NP)) {
            const uint32_t tph = t / NP, pat = t % NP;
            if (tph < n) {
                const int16_t* sy = row + size_t(first_step + tph) * R;
                uint32_t tot = 0, inv = c.c_inv;
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const uint32_t sv = uint32_t(uint16_t(__ldg(sy + i))) << 16;
                    const uint32_t lo = sv + c.c_low, hi = c.c_high - sv;       // s - low, high - s  (scalar.h:96-105 for s in [low, high])
                    const bool bb = (pat >> i) & 1u;
                    tot += bb ? hi : lo;                    // viterbi_branch_table.h:52 + scalar.h:66-73
                    inv += bb ? lo : hi;                    // scalar.h:107 (max_error - total): complementary pattern + c_inv
                }
                tbl[tph * NP + pat] = make_uint2(tot, inv);
            }
        }
    };

    uint32_t tag = 1u, pst = 0, r = 0;
    // history record of the period that just ended (position order), then clear the history fields
    auto emit_record = [&]() {
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x5410u);
            if constexpr (TIE_SIMD) w[i] = ~w[i] & (0x00010001u * ((1u << pst) - 1u));      // the tag marked path 0: decision = !tag
        }
        uint32_t* dst = rec + size_t(r) * (size_t(C::NS) / 2);
#pragma unroll
        for (int v = 0; v < NW / 4; v++) *reinterpret_cast<uint4*>(dst + size_t(v) * (H::T * 4)) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] &= 0xffff0000u;
        r++;
        tag = 1u;
        pst = 0;
    };
    auto cta_min = [&]() -> uint32_t {                   // the metric field of the smallest register is the smallest metric
        uint32_t m = x[0];
#pragma unroll
        for (int q = 1; q < NL; q++) m = min(m, x[q]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((t & 31u) == 0u) red[t >> 5] = m;
        __syncthreads();
        uint32_t rr = red[0];
#pragma unroll
        for (int w = 1; w < H::WARPS; w++) rr = min(rr, red[w]);
        __syncthreads();
        return rr;
    };

    uint32_t done = 0;
    // prologue: the exchange buffer always holds the registers at the start of the current group (for the rollback)
#pragma unroll
    for (int q = 0; q < NL; q++) xch[S::slot((uint32_t(q) << LOGT) | t)] = x[q];
    if (p.n_steps) build_tables(0u, p.n_steps < uint32_t(LB) ? p.n_steps : uint32_t(LB));
    __syncthreads();

    while (done < p.n_steps) {
        const uint32_t left = p.n_steps - done, span = left < uint32_t(LB) ? left : uint32_t(LB);
        const uint32_t tag0 = tag, pst0 = pst, r0 = r;

        // ---- speculative run of the group (no renormalisation); thread 0 keeps the running maximum of state 0's register
        uint32_t mx = 0u;
        auto spec_phase = [&](auto PHc, auto guard_tag) {
            constexpr int PH = decltype(PHc)::value;
            constexpr bool GUARD = decltype(guard_tag)::value;
            if constexpr (GUARD) { if (uint32_t(PH) >= span) return; }
            hc_bfly_all<C, PH, TIE_SIMD>(x, tbl + PH * NP, pt[PH], tag, one, std::make_integer_sequence<int, NL>{});
            mx = max(mx, x[0]);
            tag <<= 1;
            pst++;
            if constexpr (PH != LB - 1) { if (pst == uint32_t(HB)) emit_record(); }     // after the last phase: behind the exchange
        };
        if (span == uint32_t(LB)) {
            spec_phase(std::integral_constant<int, 0>{}, std::false_type{});
            spec_phase(std::integral_constant<int, 1>{}, std::false_type{});
            spec_phase(std::integral_constant<int, 2>{}, std::false_type{});
            spec_phase(std::integral_constant<int, 3>{}, std::false_type{});
            spec_phase(std::integral_constant<int, 4>{}, std::false_type{});
        } else {
            spec_phase(std::integral_constant<int, 0>{}, std::true_type{});
            spec_phase(std::integral_constant<int, 1>{}, std::true_type{});
            spec_phase(std::integral_constant<int, 2>{}, std::true_type{});
            spec_phase(std::integral_constant<int, 3>{}, std::true_type{});
        }
        if (t == 0) flag[0] = (mx >= c.thr) ? 1u : 0u;       // metric >= threshold  <=>  register >= threshold << 16
        __syncthreads();                                     // B1: all steps done, tables and exchange buffer free again
        const uint32_t any_trig = flag[0];

        if (any_trig) {
            // ---- roll back (registers and record bookkeeping) and replay step by step with the reference's renormalisation
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xch[S::slot((uint32_t(q) << LOGT) | t)];
            tag = tag0; pst = pst0; r = r0;
            auto replay_phase = [&](auto PHc) {
                constexpr int PH = decltype(PHc)::value;
                if (uint32_t(PH) < span) {
                    hc_bfly_all<C, PH, TIE_SIMD>(x, tbl + PH * NP, pt[PH], tag, one, std::make_integer_sequence<int, NL>{});
                    tag <<= 1;
                    pst++;
                    if (t == 0) flag[1] = x[0];
                    __syncthreads();
                    const uint32_t x00 = flag[1];
                    if (x00 >= c.thr) {                                    // uniform across the CTA (scalar.h:48)
                        const uint32_t sub = cta_min() & 0xffff0000u;      // scalar.h:140-146
#pragma unroll
                        for (int q = 0; q < NL; q++) x[q] -= sub;          // scalar.h:148-150
                        acc += uint64_t(sub >> 16);                        // scalar.h:49, 152 (only thread 0's copy is stored)
                    } else {
                        __syncthreads();                                   // flag[1] may be rewritten by the next phase
                    }
                    if constexpr (PH != LB - 1) { if (pst == uint32_t(HB)) emit_record(); }
                }
            };
            replay_phase(std::integral_constant<int, 0>{});
            replay_phase(std::integral_constant<int, 1>{});
            replay_phase(std::integral_constant<int, 2>{});
            replay_phase(std::integral_constant<int, 3>{});
            replay_phase(std::integral_constant<int, 4>{});
            __syncthreads();
        }

        done += span;
        const bool full = span == uint32_t(LB);
        if (full) {
            // ---- exchange: value at (q, t) moves to PHI' = (t << LB) | q, bringing the layout back to PHI = s
#pragma unroll
            for (int q = 0; q < NL; q++) xch[S::slot((t << LB) | uint32_t(q))] = x[q];
        }
        if (done < p.n_steps) {
            const uint32_t nleft = p.n_steps - done;
            build_tables(done, nleft < uint32_t(LB) ? nleft : uint32_t(LB));
        }
        __syncthreads();                                     // B2: exchange data and the next group's tables are in place
        if (full) {
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xch[S::slot((uint32_t(q) << LOGT) | t)];
            // no barrier needed here: the buffer is next written after B1 of the next group, which every thread reaches only
            // after these loads; until then it doubles as the rollback copy of the group's starting registers
            if (pst == uint32_t(HB)) {
                emit_record();                               // records are cut in the layout the traceback expects: PHI = rotr^(steps mod 5)(s)
                // the rollback copy must match the registers the next group starts from (history fields cleared); every thread
                // rewrites exactly the words it has just read and will read back on a rollback: no barrier needed
#pragma unroll
                for (int q = 0; q < NL; q++) xch[S::slot((uint32_t(q) << LOGT) | t)] = x[q];
            }
        }
    }
    if (pst) emit_record();                                  // last, partial record (its pst is still needed for the SIMD tie-break mask)

    // final metrics in logical order: state s sits at PHI = rotr^ph(s)   (core.h:195-199 reads old_metrics[end_state])
    const uint32_t ph = p.n_steps % uint32_t(LB);
    uint16_t* m = p.metrics + f * C::NS;
#pragma unroll
    for (int q = 0; q < NL; q++) m[rotl_rt((uint32_t(q) << LOGT) | t, ph, SB)] = uint16_t(x[q] >> 16);
    if (t == 0) p.acc[f] = acc;
}

}  // namespace vitb
