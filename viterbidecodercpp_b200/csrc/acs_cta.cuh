// acs_cta.cuh -- add-compare-select for long constraint lengths (K = 15, Cassini: 16384 states): ONE CTA PER FRAME PAIR.
//
// 512 threads hold the 16384 packed path metrics of two frames in registers (32 each).  Same position algebra as acs_group.cuh
// with LOGT = 9: PHI = (q << 9) | t, after n steps state s sits at PHI = rotr^n(s), LB = 5 steps run with register-only
// butterflies, then the CTA exchanges through 64 KB of shared memory (XOR-swizzled, conflict-free) under two __syncthreads.
//
// Differences from the warp-group kernel, all forced by the size of the problem:
//   * branch metrics: R = 6 gives 64 patterns, too many for a per-thread register table.  The table of one whole group of steps
//     (LB x 64 entries of {total, inverted total}) is built cooperatively into shared memory at each exchange, 320 threads
//     computing one entry each; a butterfly fetches its entry with one LDS.64 at index (register part ^ lane part of the pattern).
//   * renormalisation: the trigger (state 0's metric >= threshold, scalar.h:48) is only visible to thread 0.  Steps run
//     SPECULATIVELY without it; thread 0 records a trigger, the CTA learns of it at the next exchange barrier, ROLLS BACK to the
//     metrics at the start of the group (still in the exchange buffer) and replays the group one step at a time with a barrier
//     after every step, renormalising at exactly the reference's step.  Rare (a handful of times per frame), so the common
//     path has no per-step CTA synchronisation, and the replay makes the result exact whatever the speculative steps did.
//   * decision rows: thread t writes 32 bits per frame and step (bit q = decision of the state in its register q after the
//     step), 4 KB per pair and step = the reference's 2048 B per frame and step, fully coalesced.  traceback_cta_kernel reads it.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"

namespace vitb {

template <class C, int LT>
struct CtaShape {
    static constexpr int LOGT = LT;                  // 9: 512 threads x 32 registers, 10: 1024 threads x 16 registers
    static constexpr int T = 1 << LOGT;              // threads per CTA = per frame pair
    static constexpr int SB = C::SB;
    static constexpr int LB = SB - LOGT;             // steps between exchanges
    static_assert(LB >= 1 && LB <= 5 && LOGT >= 5 && LOGT <= 10, "CtaShape is meant for K = 11..15");
    static constexpr int NL = 1 << LB;               // packed registers per thread
    static constexpr int NP = C::NP;                 // branch patterns
    static constexpr int WARPS = T / 32;
    static constexpr size_t XCH_WORDS = size_t(C::NS) + size_t(C::NS) / 8 + 32;    // skewed by four words per 32
    static constexpr size_t TBL_WORDS = size_t(LB) * NP * 4;          // [LB][NP] 16-byte slots {own entry, partner's entry}
    // Exchange buffer and tables are double-buffered.  The two table sets lie a power of two apart, so that "which set" is one bit of
    // the per-thread byte offset a fetch XORs its pattern into - the base stays a constant the load takes as immediate.
    static constexpr uint32_t TBL_SET_BYTES = 8192;
    static_assert(TBL_WORDS * 4 <= TBL_SET_BYTES, "a table set must fit its slot");
    static constexpr size_t SMEM_BYTES = (2 * XCH_WORDS + 64) * 4 + 2 * TBL_SET_BYTES;
    // Skew instead of XOR so that every access is (per-thread base) + (compile-time offset): the exchange writes position
    // (t << LB) | q -> word 36t + q (LB = 5) and reads (q << LOGT) | t -> word q * (T + T/8) + t + 4 (t >> 5); conflict-free on both
    // sides (tests/test_host_cpu.py::test_cta_exchange_skew_is_conflict_free).
    static __host__ __device__ constexpr uint32_t slot(uint32_t phi) { return phi + ((phi >> 5) << 2); }
    // the same, split into a per-thread offset and a compile-time one (the buffers alternate, so their base is a run-time value and
    // ptxas would otherwise recompute the whole expression for every register and group):
    //   slot((q << LOGT) | t) = rd_off(t) + q * RD_STRIDE,   slot((t << LB) | q) = wr_off(t) + q   (q < 2^LB <= 32)
    static constexpr uint32_t RD_STRIDE = uint32_t(T) + uint32_t(T) / 8;
    static __host__ __device__ constexpr uint32_t rd_off(uint32_t t) { return t + ((t >> 5) << 2); }
    static __host__ __device__ constexpr uint32_t wr_off(uint32_t t) { return slot(t << LB); }
    static_assert(LB <= 5 && LB >= 2 && LOGT >= 5, "split form of slot()");
    // Four words of skew per 32 (not one) keep a thread's 2^LB consecutive words 16-byte aligned: the write side of the exchange is
    // STS.128 - a quarter warp of them hits 8 different bank groups (36 t and 16 t + 4 (t >> 1) modulo 32 words, t = 8k .. 8k+7) -,
    // the read side stays 32 consecutive words per warp and register.
};

// TABLE FETCH BY BUTTERFLY PAIRS (OR QUADS).  The entry of a butterfly is T[pq ^ pt] (pq: pattern of the register bits, a compile-time constant;
// pt: pattern of the thread bits).  Two butterflies whose registers differ in one bit (the "pairing bit") have patterns that differ by
// a constant v, so in slot coordinates M with M(v) = 1 (M linear and invertible) their entries sit at idx and idx ^ 1.  The table
// stores every entry twice - slot idx = {T[idx], T[idx ^ 1]}, 16 bytes - so that ONE LDS.128 at slot M(pq) ^ M(pt) fetches the entries
// of both butterflies, in butterfly order whatever the thread's M(pt) is: 8 LDS.128 per step instead of 16 LDS.64 for the same bytes,
// and few enough addresses that ptxas keeps them in registers for the whole kernel (no XOR per fetch).  The three low rows of M are
// chosen so that the 8 lanes of a quarter warp read 8 different 16-byte bank groups (profiles/microbench/lds128_merge.cu: an LDS.128
// is served a quarter warp at a time, lanes of different quarters never share a wavefront).
// Used by acs_hist_cta.cuh (quads of 4-byte entries) and by the frame-pair kernel below (pairs of {total, inverted}).
// Slot coordinates of the table of one phase: index bit i of pattern p = parity(p & row[i]); pb = the pairing register bit.
struct PairMap {
    uint32_t row[8];
    int pb[2];
};
__host__ __device__ constexpr uint32_t pair_apply(const PairMap& m, int R, uint32_t p) {
    uint32_t idx = 0;
    for (int i = 0; i < R; i++) idx |= parity32(p & m.row[i]) << i;
    return idx;
}
// insert r into an XOR basis kept in echelon form by leading bit; false if r depends on the basis
__host__ __device__ constexpr bool pair_basis_add(uint32_t (&lead)[8], uint32_t r) {
    for (int b = 7; b >= 0; b--) {
        if (!((r >> b) & 1u)) continue;
        if (lead[b] == 0) { lead[b] = r; return true; }
        r ^= lead[b];
    }
    return false;
}
// NV pairing bits (1: pairs, 2: quads of butterflies per slot; slot idx then holds the entries idx, idx ^ 1, idx ^ 2, idx ^ 3).
// EXCL_NEXT: a pairing bit must not be the bit that says "path-1 operand of the next step" either (acs_hist_cta.cuh)
template <class C, int LOGT, bool EXCL_NEXT, int NV>
__host__ __device__ constexpr PairMap pair_map(int PH) {
    constexpr int SB = C::SB, LB = SB - LOGT, R = C::R;
    static_assert(R >= 3 && R <= 8 && NV >= 1 && NV <= 2, "the slot map is built for 8 <= 2^R <= 256 patterns and pairs or quads");
    PairMap m{};
    // pairing bits: not the butterfly bit (LB-1-PH), not the next step's (LB-2-PH), patterns independent
    uint32_t v[2] = {}, vlead[8] = {};
    int nv = 0;
    m.pb[0] = m.pb[1] = -1;
    for (int b = 0; b < LB && nv < NV; b++) {
        if (b == LB - 1 - PH || (EXCL_NEXT && b == LB - 2 - PH)) continue;
        const uint32_t pv = bfly_pattern<C>(rotl_bits((1u << b) << LOGT, PH, SB));
        if (pair_basis_add(vlead, pv)) { m.pb[nv] = b; v[nv++] = pv; }
    }
    // patterns of the three lane bits that vary inside a quarter warp
    uint32_t pl[3] = {};
    for (int k = 0; k < 3; k++) pl[k] = bfly_pattern<C>(rotl_bits(1u << k, PH, SB));
    auto rho = [&](uint32_t r) { return parity32(r & pl[0]) | (parity32(r & pl[1]) << 1) | (parity32(r & pl[2]) << 2); };
    uint32_t lead[8] = {}, lead_rho[8] = {};
    const uint32_t NPAT = 1u << R;
    // row i < NV is 1 on v[i] and 0 on the other pairing pattern; every other row vanishes on all of them (M(v[i]) = 2^i exactly).
    // The three low rows additionally get independent restrictions to the lane patterns where that is possible (pass 0); the rows
    // with the fewest candidates left are not the problem here, so the free ones go first.
    auto pick = [&](int want0, int want1, bool need_rho) -> uint32_t {
        for (int pass = need_rho ? 0 : 1; pass < 2; pass++)
            for (uint32_t r = 1; r < NPAT; r++) {
                if (int(parity32(r & v[0])) != want0) continue;
                if (NV > 1 && int(parity32(r & v[1])) != want1) continue;
                uint32_t t1[8] = {}, t2[8] = {};
                for (int i = 0; i < 8; i++) { t1[i] = lead[i]; t2[i] = lead_rho[i]; }
                if (!pair_basis_add(t1, r)) continue;
                if (pass == 0 && !pair_basis_add(t2, rho(r))) continue;
                for (int i = 0; i < 8; i++) { lead[i] = t1[i]; if (pass == 0) lead_rho[i] = t2[i]; }
                return r;
            }
        return 0u;
    };
    for (int i = NV; i < 3; i++) m.row[i] = pick(0, 0, true);
    for (int i = NV - 1; i >= 0; i--) m.row[i] = pick(i == 0 ? 1 : 0, i == 1 ? 1 : 0, true);
    for (int i = 3; i < R; i++) m.row[i] = pick(0, 0, false);
    return m;
}
// do the 8 lanes of a quarter warp read 8 different 16-byte bank groups (or the same slot)?
template <class C, int LOGT, bool EXCL_NEXT, int NV>
__host__ __device__ constexpr bool pair_map_conflict_free(int PH) {
    const PairMap m = pair_map<C, LOGT, EXCL_NEXT, NV>(PH);
    uint32_t idx[8] = {};
    for (uint32_t l = 0; l < 8; l++) idx[l] = pair_apply(m, C::R, bfly_pattern<C>(rotl_bits(l, PH, C::SB)));
    for (int a = 0; a < 8; a++)
        for (int b = a + 1; b < 8; b++)
            if (idx[a] != idx[b] && (idx[a] & 7u) == (idx[b] & 7u)) return false;
    return true;
}
template <class C, int LOGT, bool EXCL_NEXT, int NV, int PH, int... Is>
__device__ __forceinline__ uint32_t pair_apply_dyn_impl(uint32_t p, std::integer_sequence<int, Is...>) {
    constexpr PairMap M = pair_map<C, LOGT, EXCL_NEXT, NV>(PH);
    uint32_t idx = 0;
    ((idx |= (uint32_t(__popc(p & std::integral_constant<uint32_t, M.row[Is]>::value)) & 1u) << Is), ...);
    return idx;
}
template <class C, int LOGT, bool EXCL_NEXT, int NV, int PH>
__device__ __forceinline__ uint32_t pair_apply_dyn(uint32_t p) {
    return pair_apply_dyn_impl<C, LOGT, EXCL_NEXT, NV, PH>(p, std::make_integer_sequence<int, C::R>{});
}
// same for a run-time phase (table build)
template <class C, int LOGT, bool EXCL_NEXT, int NV, int... PHs>
__device__ __forceinline__ uint32_t pair_apply_phase(uint32_t ph, uint32_t p, std::integer_sequence<int, PHs...>) {
    uint32_t idx = 0;
    ((idx = (ph == uint32_t(PHs)) ? pair_apply_dyn<C, LOGT, EXCL_NEXT, NV, PHs>(p) : idx), ...);
    return idx;
}
// the pattern that sits at slot index idx (inverse of pair_apply): XOR of the patterns of the index bits
template <class C, int LOGT, bool EXCL_NEXT, int NV>
__host__ __device__ constexpr uint32_t pair_column(int PH, int k) {
    const PairMap m = pair_map<C, LOGT, EXCL_NEXT, NV>(PH);
    for (uint32_t p = 0; p < (1u << C::R); p++)
        if (pair_apply(m, C::R, p) == (1u << k)) return p;
    return 0u;
}
template <class C, int LOGT, bool EXCL_NEXT, int NV, int PH, int... Ks>
__device__ __forceinline__ uint32_t pair_unapply_dyn_impl(uint32_t idx, std::integer_sequence<int, Ks...>) {
    uint32_t p = 0;
    ((p ^= ((idx >> Ks) & 1u) ? std::integral_constant<uint32_t, pair_column<C, LOGT, EXCL_NEXT, NV>(PH, Ks)>::value : 0u), ...);
    return p;
}
template <class C, int LOGT, bool EXCL_NEXT, int NV, int... PHs>
__device__ __forceinline__ uint32_t pair_unapply_phase(uint32_t ph, uint32_t idx, std::integer_sequence<int, PHs...>) {
    uint32_t p = 0;
    ((p = (ph == uint32_t(PHs)) ? pair_unapply_dyn_impl<C, LOGT, EXCL_NEXT, NV, PHs>(idx, std::make_integer_sequence<int, C::R>{}) : p), ...);
    return p;
}
// slot part of thread t for every phase
template <class C, int LOGT, bool EXCL_NEXT, int NV, int... PHs>
__device__ __forceinline__ void pair_thread_slots(uint32_t (&mpt)[sizeof...(PHs)], uint32_t t, std::integer_sequence<int, PHs...>) {
    ((mpt[PHs] = pair_apply_dyn<C, LOGT, EXCL_NEXT, NV, PHs>(bfly_pattern_dyn<C>(rotl_bits(t, PHs, C::SB)))), ...);
}


// float accumulators per frame: one per 8 decision bits, each split into two chains of 4 predicated FADDs (low nibble starts at 2^23,
// high nibble at 0, summed exactly at the end).  Short chains let ptxas consume the VIMNMX predicates quickly: with chains of 8 it
// ran out of predicate registers and spilled them through pairs of LOP3 bit-mask updates (36 LOP3 of 171 instructions per step in
// profiles/r01_ncu_full_acs_cta_cfg5.txt; with 2 x 16 bits per frame it was +81 instructions).
template <int NL> struct CtaAcc { static constexpr int value = NL >= 32 ? 4 : (NL >= 16 ? 2 : 1); };
#ifndef VITB_CTA_CHAINS
#define VITB_CTA_CHAINS 4
#endif
constexpr int CTA_CHAINS = VITB_CTA_CHAINS;          // chains per decision byte (1, 2 or 4)

// store registers x[Q0 .. Q0+3] of a thread into the exchange buffer with one STS.128 (xout + Q0 is 16-byte aligned: wr_off is a
// multiple of 4 words and so is Q0)
template <int Q0, int NL>
__device__ __forceinline__ void cta_store4(uint32_t* xout, const uint32_t (&x)[NL]) {
    static_assert(Q0 % 4 == 0 && Q0 + 3 < NL, "aligned group of four registers");
    *reinterpret_cast<uint4*>(xout + Q0) = make_uint4(x[Q0], x[Q0 + 1], x[Q0 + 2], x[Q0 + 3]);
}

// ARITHMETIC DECISIONS (the wrapping flavours).  The predicate form below pays ~1.4 instructions per decision bit on top of its
// 4 adds + 2 min per butterfly (predicated FADD / FSEL + FADD, 89 of 290 instructions per step, and the kernel is issue bound).
// Here the min is fused with one of the adds and the decision is read off the result:
//     a = x0 + ex;  m = min(x1 + ey, a)  [VIADDMNMX.U16x2];  a - m != 0  <=>  path 1 strictly better: the reference's decision
//     d = min(a - m, 1) per half (a plain 32-bit subtraction: m <= a in both halves);  word += d << q                          (scalar.h:113-128; A bits in the low half, B in the high)
// 5 instructions per new state for both frames instead of 5.9.  The SIMD tie-break (a tie selects path 1) uses the path-1 sum as
// the explicit one and collects the complement, inverted when the word is stored.  ia[h][c]: registers 16h .. 16h+15, chain c.
template <class C, int LT, int TIE_SIMD, bool STORE, int q0, int q1>
__device__ __forceinline__ void cta_bfly_arith_one(uint32_t (&x)[CtaShape<C, LT>::NL], const uint32_t ex, const uint32_t ey, uint32_t (&ia)[2][2], uint32_t* xout) {
    static_assert(TIE_SIMD == 0 || TIE_SIMD == 1, "the saturating flavour keeps the predicate form");
    uint32_t m0, m1, t0, t1;
    if constexpr (TIE_SIMD == 0) {
        const uint32_t a0 = __vadd2(x[q0], ex), a1 = __vadd2(x[q0], ey);            // scalar.h:113, 115
        m0 = __viaddmin_u16x2(x[q1], ey, a0);                                       // scalar.h:114, 127
        m1 = __viaddmin_u16x2(x[q1], ex, a1);                                       // scalar.h:116, 128
        t0 = a0 - m0;                                                               // != 0 per half  <=>  path 1 strictly better
        t1 = a1 - m1;                                                               // (m <= a in both halves: no borrow crosses)
    } else {
        const uint32_t b0 = __vadd2(x[q1], ey), b1 = __vadd2(x[q1], ex);
        m0 = __viaddmin_u16x2(x[q0], ex, b0);
        m1 = __viaddmin_u16x2(x[q0], ey, b1);
        t0 = b0 - m0;                                                               // != 0 per half  <=>  path 0 strictly better
        t1 = b1 - m1;
    }
    x[q0] = m0;
    x[q1] = m1;
    const uint32_t d0 = __vminu2(t0, 0x00010001u), d1 = __vminu2(t1, 0x00010001u);
    ia[q0 >> 4][(q0 >> 3) & 1] += d0 << (q0 & 15);
    ia[q1 >> 4][(q1 >> 3) & 1] += d1 << (q1 & 15);
}

// STORE (last phase of a full group): both results go straight to the other exchange buffer, xout = buffer + slot(t << LB), so that
// the stores of the exchange run under the butterflies of the phase instead of behind a barrier
template <class C, int LT, int TIE_SIMD, bool STORE, int q0, int q1>
__device__ __forceinline__ void cta_bfly_pred_one(uint32_t (&x)[CtaShape<C, LT>::NL], const uint32_t ex, const uint32_t ey,
                                                  float (&fa)[2][CtaAcc<CtaShape<C, LT>::NL>::value][CTA_CHAINS], uint32_t* xout) {
    using S = CtaShape<C, LT>;
    constexpr bool SAT = Sat<TIE_SIMD>::value;  // saturating flavour, see acs_pair.cuh
    const uint32_t a0 = metric_add<SAT>(x[q0], ex), b0 = metric_add<SAT>(x[q1], ey);     // scalar.h:113-114
    const uint32_t a1 = metric_add<SAT>(x[q0], ey), b1 = metric_add<SAT>(x[q1], ex);     // scalar.h:115-116
    bool h0, l0, h1, l1, dA0, dB0, dA1, dB1;
    if constexpr (!TIE_SIMD) {
        x[q0] = __vibmin_u16x2(a0, b0, &h0, &l0);
        x[q1] = __vibmin_u16x2(a1, b1, &h1, &l1);
        dA0 = !l0; dB0 = !h0; dA1 = !l1; dB1 = !h1;
    } else {
        x[q0] = __vibmin_u16x2(b0, a0, &h0, &l0);
        x[q1] = __vibmin_u16x2(b1, a1, &h1, &l1);
        dA0 = l0; dB0 = h0; dA1 = l1; dB1 = h1;
    }
    constexpr int NACC = CtaAcc<S::NL>::value;
    constexpr int acc0 = (q0 >> 3) % NACC, acc1 = (q1 >> 3) % NACC, n0 = ((q0 & 7) * CTA_CHAINS) >> 3, n1 = ((q1 & 7) * CTA_CHAINS) >> 3;
    constexpr float w0 = float(1u << (q0 & 7)), w1 = float(1u << (q1 & 7));
    if (dA0) fa[0][acc0][n0] += w0;
    if (dB0) fa[1][acc0][n0] += w0;
    if (dA1) fa[0][acc1][n1] += w1;
    if (dB1) fa[1][acc1][n1] += w1;
}

// the two butterflies (Q, Q | bit) and (Q | pbit, Q | pbit | bit) of a pair: one 16-byte table fetch (PairMap above); pt is the
// thread's byte offset into the table of the phase and also selects the table set (CtaShape::TBL_SET_BYTES)
template <class C, int LT, int PH, int TIE_SIMD, bool STORE, int Q, class Acc>
__device__ __forceinline__ void cta_bfly_pair_at(uint32_t (&x)[CtaShape<C, LT>::NL], const uint4* tbl_ph, const uint32_t pt, Acc& acc, uint32_t* xout) {
    using S = CtaShape<C, LT>;
    constexpr PairMap M = pair_map<C, LT, false, 1>(PH);
    static_assert(M.pb[0] >= 0, "no pairing bit with a non-zero pattern");
    static_assert(pair_map_conflict_free<C, LT, false, 1>(PH), "table fetch with shared-memory bank conflicts inside a quarter warp");
    constexpr int bit = 1 << (S::LB - 1 - PH), pbit = 1 << M.pb[0];
    if constexpr ((Q & (bit | pbit)) == 0) {
        constexpr uint32_t ia = pair_apply(M, C::R, bfly_pattern<C>(rotl_bits(uint32_t(Q) << S::LOGT, PH, S::SB)));
        static_assert(pair_apply(M, C::R, bfly_pattern<C>(rotl_bits(uint32_t(Q | pbit) << S::LOGT, PH, S::SB))) == (ia ^ 1u),
                      "paired butterflies must sit in one table slot");
        // {total_error, inverted_error} of both butterflies   (scalar.h:66-73, 107)
        const uint4 e = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(tbl_ph) + ((ia << 4) ^ pt));
        if constexpr (Sat<TIE_SIMD>::value) {
            cta_bfly_pred_one<C, LT, TIE_SIMD, STORE, Q, Q | bit>(x, e.x, e.y, acc, xout);
            cta_bfly_pred_one<C, LT, TIE_SIMD, STORE, Q | pbit, Q | pbit | bit>(x, e.z, e.w, acc, xout);
        } else {
            cta_bfly_arith_one<C, LT, TIE_SIMD, STORE, Q, Q | bit>(x, e.x, e.y, acc, xout);
            cta_bfly_arith_one<C, LT, TIE_SIMD, STORE, Q | pbit, Q | pbit | bit>(x, e.z, e.w, acc, xout);
        }
        if constexpr (STORE) {                      // slot((t << LB) | q) = slot(t << LB) + q
            if constexpr (bit == 1 && pbit == 2) {
                cta_store4<Q>(xout, x);             // the pair's four registers are Q .. Q+3
            } else {
                xout[Q] = x[Q]; xout[Q | bit] = x[Q | bit]; xout[Q | pbit] = x[Q | pbit]; xout[Q | pbit | bit] = x[Q | pbit | bit];
            }
        }
    }
}
template <class C, int LT, int PH, int TIE_SIMD, bool STORE, class Acc, int... Qs>
__device__ __forceinline__ void cta_bfly_all(uint32_t (&x)[CtaShape<C, LT>::NL], const uint4* tbl_ph, const uint32_t pt, Acc& acc, uint32_t* xout,
                                             std::integer_sequence<int, Qs...>) {
    (cta_bfly_pair_at<C, LT, PH, TIE_SIMD, STORE, Qs>(x, tbl_ph, pt, acc, xout), ...);
}

template <class C, int LT, int SH, int TIE_SIMD>
struct CtaKernel {
    using S = CtaShape<C, LT>;
    static constexpr int LB = S::LB, NL = S::NL, R = C::R, NP = C::NP, T = S::T, SB = S::SB;

    // one trellis step at compile-time phase PH; decisions to dec_row (this thread's uint2 of the row)
    static constexpr int W = NL > 16 ? 2 : 1;     // 32-bit decision words per thread and step
    // dec_row: this thread's W words of the row.  NL = 32: {A word, B word}; NL = 16: one word, A bits | B bits << 16
    template <int PH, bool STORE = false>
    static __device__ __forceinline__ void step(uint32_t (&x)[NL], const uint4* tbl, const uint32_t (&pt)[LB], uint32_t* dec_row, uint32_t* xout = nullptr) {
        static_assert(NL == 32 || NL == 16, "decision word packing assumes 32 or 16 registers per thread");
        if constexpr (!Sat<TIE_SIMD>::value) {
            uint32_t ia[2][2] = {{0u, 0u}, {0u, 0u}};
            cta_bfly_all<C, LT, PH, TIE_SIMD, STORE>(x, tbl + PH * NP, pt[PH], ia, xout, std::make_integer_sequence<int, NL>{});
            constexpr uint32_t inv = TIE_SIMD ? 0xffffffffu : 0u;      // the SIMD tie-break collected the complement
            const uint32_t lo = ia[0][0] + ia[0][1];                   // A bits 0..15 | B bits 0..15 << 16
            if constexpr (NL == 32) {
                const uint32_t hi = ia[1][0] + ia[1][1];               // A bits 16..31 | B bits 16..31 << 16
                *reinterpret_cast<uint2*>(dec_row) = make_uint2(__byte_perm(lo, hi, 0x5410) ^ inv, __byte_perm(lo, hi, 0x7632) ^ inv);
            } else {
                dec_row[0] = lo ^ inv;
            }
            return;
        }
        constexpr int NACC = CtaAcc<NL>::value;
        float fn[2][NACC][CTA_CHAINS];
#pragma unroll
        for (int a = 0; a < NACC; a++) {
#pragma unroll
            for (int c = 0; c < CTA_CHAINS; c++) { fn[0][a][c] = c ? 0.f : 8388608.f; fn[1][a][c] = c ? 0.f : 8388608.f; }
        }
        if constexpr (Sat<TIE_SIMD>::value) cta_bfly_all<C, LT, PH, TIE_SIMD, STORE>(x, tbl + PH * NP, pt[PH], fn, xout, std::make_integer_sequence<int, NL>{});
        float fa[2][NACC];
#pragma unroll
        for (int a = 0; a < NACC; a++) {
#pragma unroll
            for (int f = 0; f < 2; f++) {
                if constexpr (CTA_CHAINS == 8) fa[f][a] = ((fn[f][a][0] + fn[f][a][1]) + (fn[f][a][2] + fn[f][a][3])) + ((fn[f][a][4] + fn[f][a][5]) + (fn[f][a][6] + fn[f][a][7]));
                else if constexpr (CTA_CHAINS == 4) fa[f][a] = (fn[f][a][0] + fn[f][a][1]) + (fn[f][a][2] + fn[f][a][3]);
                else if constexpr (CTA_CHAINS == 2) fa[f][a] = fn[f][a][0] + fn[f][a][1];
                else fa[f][a] = fn[f][a][0];
            }
        }
        // byte k of a frame's bits = mantissa byte 0 of accumulator k (registers 8k .. 8k+7)
        if constexpr (NL == 32) {
            const uint32_t a01 = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[0][1]), 0x0040);
            const uint32_t a23 = __byte_perm(__float_as_uint(fa[0][2]), __float_as_uint(fa[0][3]), 0x0040);
            const uint32_t b01 = __byte_perm(__float_as_uint(fa[1][0]), __float_as_uint(fa[1][1]), 0x0040);
            const uint32_t b23 = __byte_perm(__float_as_uint(fa[1][2]), __float_as_uint(fa[1][3]), 0x0040);
            *reinterpret_cast<uint2*>(dec_row) = make_uint2(__byte_perm(a01, a23, 0x5410), __byte_perm(b01, b23, 0x5410));
        } else {
            const uint32_t a01 = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[0][1]), 0x0040);
            const uint32_t b01 = __byte_perm(__float_as_uint(fa[1][0]), __float_as_uint(fa[1][1]), 0x0040);
            dec_row[0] = __byte_perm(a01, b01, 0x5410);
        }
    }

    // CTA-wide packed minimum of all path metrics (both halves); result in every thread
    static __device__ __forceinline__ uint32_t cta_min(const uint32_t (&x)[NL], uint32_t* red /* WARPS words of smem */) {
        uint32_t m = packed_min<NL>(x);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) m = __vminu2(m, __shfl_xor_sync(0xffffffffu, m, d));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        uint32_t r = red[0];
#pragma unroll
        for (int w = 1; w < S::WARPS; w++) r = __vminu2(r, red[w]);
        __syncthreads();
        return r;
    }
};

// grid = number of frame pairs, block = 512, dynamic shared memory = CtaShape::SMEM_BYTES
template <class C, int LT, int SH, int TIE_SIMD>
__global__ void __launch_bounds__(CtaShape<C, LT>::T, 1) acs_cta_kernel(const AcsParams p) {
    using S = CtaShape<C, LT>;
    using Kn = CtaKernel<C, LT, SH, TIE_SIMD>;
    constexpr int LB = S::LB, NL = S::NL, R = C::R, NP = C::NP, SB = S::SB, LOGT = S::LOGT;
    extern __shared__ uint32_t smem[];
    // Two exchange buffers and two table sets, used in turn (cur): buffer cur = the metrics at the start of the current group (read
    // back by the exchange, kept for the rollback), the other one takes the results of the group while it still runs.
    uint32_t* xch0 = smem;                                            // [2][NS (skewed)]
    uint4* tbl0 = reinterpret_cast<uint4*>(smem + 2 * S::XCH_WORDS);  // [2 sets, TBL_SET_BYTES apart][LB][NP] {total, inverted} of slot idx, of slot idx ^ 1
    static_assert((2 * S::XCH_WORDS) % 4 == 0, "16-byte table slots");
    uint32_t* red = smem + 2 * S::XCH_WORDS + 2 * (S::TBL_SET_BYTES / 4);   // [WARPS] reduction scratch
    uint32_t* flag = red + S::WARPS;                                  // [0..1] trigger flag of the groups in turn, [2] replay scratch
    static_assert(S::WARPS + 3 <= 64, "scratch words");

    const uint32_t t = threadIdx.x, pair = blockIdx.x;
    const uint32_t rd_off = S::rd_off(t), wr_off = S::wr_off(t);
    const size_t fA = size_t(pair) * 2, fB = fA + 1;
    // the grid covers whole 64-frame blocks of the packed stream: pairs behind the last frame have nothing to decode (they used to
    // decode padding: 888 frames ran 448 CTAs, one more wave than 444 on 148 SMs)
    if (fA >= size_t(p.n_frames)) return;
    uint16_t* mA = p.metrics + fA * C::NS;
    uint16_t* mB = p.metrics + fB * C::NS;

    // thread part of the table slot per phase, as byte offset into the table of the phase (+ the table set, see the loop)
    uint32_t pt[LB];
    pair_thread_slots<C, LT, false, 1>(pt, t, std::make_integer_sequence<int, LB>{});
#pragma unroll
    for (int n = 0; n < LB; n++) pt[n] <<= 4;

    int ph = int(p.dec_row0 % uint32_t(LB));
    uint32_t x[NL];
    uint64_t accA = 0, accB = 0;
    if (p.resume) {
#pragma unroll
        for (int q = 0; q < NL; q++) {
            const uint32_t s = rotl_bits((uint32_t(q) << LOGT) | t, ph, SB);
            x[q] = ((uint32_t(mA[s]) << SH) & 0xffffu) | (uint32_t(mB[s]) << (16 + SH));
        }
        if (t == 0) { accA = p.acc[fA]; accB = p.acc[fB]; }
    } else {
        const uint32_t s0 = p.start_state & uint32_t(C::NS - 1);
#pragma unroll
        for (int q = 0; q < NL; q++) {
            const uint32_t s = rotl_bits((uint32_t(q) << LOGT) | t, ph, SB);
            x[q] = (s == s0) ? p.init_start2 : p.init_other2;
        }
    }

#pragma unroll
    for (int q = 0; q < NL; q++) x[q] = metric_ones<Sat<TIE_SIMD>::value, SH>(x[q]);       // saturating flavour: low byte 0xFF (acs_pair.cuh)

    const uint32_t* pk = p.pk + (size_t(pair) * (p.pk_steps ? p.pk_steps : p.n_steps) + p.pk_step0) * R;   // [steps][R], this launch's first step
    constexpr int W = Kn::W;
    uint32_t* dec = static_cast<uint32_t*>(p.dec) + ((size_t(pair) * p.dec_rows + p.dec_row0) * S::T + t) * W;

    // branch metric tables {total, inverted} of the steps of one group, built by the first LB*NP threads (one entry each)
    auto build_tables = [&](uint4* tbl, uint32_t first_step, int first_phase, uint32_t n) {
        if (t < uint32_t(LB * NP)) {
            const int tph = int(t) / NP;
            const uint32_t pat = t % NP;
            const int k = tph - first_phase;
            if (k >= 0 && uint32_t(k) < n) {
                const uint32_t* sy = pk + size_t(first_step + uint32_t(k)) * R;
                uint32_t tot = 0, inv = p.c_inv2;
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const uint32_t sv = __ldg(sy + i);
                    const uint32_t lo = __vadd2(sv, p.c_low2), hi = __vadd2(~sv, p.c_high2);
                    const bool bb = (pat >> i) & 1u;
                    if constexpr (Sat<TIE_SIMD>::value) {
                        tot = metric_field<true, SH>(__vaddus2(tot, bb ? hi : lo));       // avx_u16.h:95-97: saturating sum of the errors
                    } else {
                        tot = __vadd2(tot, bb ? hi : lo);      // viterbi_branch_table.h:52 + scalar.h:66-73
                        inv = __vadd2(inv, bb ? lo : hi);      // scalar.h:107 (max_error - total), complementary pattern + c_inv
                    }
                }
                if constexpr (Sat<TIE_SIMD>::value) inv = __vsubus2(p.max_err2, tot);     // avx_u16.h:106: max_error -sat total
                const uint32_t idx = pair_apply_phase<C, LT, false, 1>(uint32_t(tph), pat, std::make_integer_sequence<int, LB>{});
                uint2* t2 = reinterpret_cast<uint2*>(tbl + size_t(tph) * NP);
                t2[2 * idx] = make_uint2(tot, inv);                    // own member of its slot,
                t2[2 * (idx ^ 1u) + 1] = make_uint2(tot, inv);         // partner member of the neighbouring one
            }
        }
    };
    auto group_span = [&](uint32_t done_, int ph_) {
        const uint32_t left = p.n_steps - done_;
        return uint32_t(LB - ph_) < left ? uint32_t(LB - ph_) : left;
    };

    // the packed symbols of a group are pulled into L1 one group ahead of the table build that reads them
    auto prefetch_symbols = [&](uint32_t first_step) {
        if (t < uint32_t(LB)) {
            const uint32_t st = first_step + t;
            if (st < p.n_steps) asm volatile("prefetch.global.L1 [%0];" ::"l"(pk + size_t(st) * R));
        }
    };

    uint32_t done = 0, cur = 0;
    // prologue: buffer 0 holds the metrics at the start of the first group, table set 0 its tables
#pragma unroll
    for (int q = 0; q < NL; q++) xch0[rd_off + uint32_t(q) * S::RD_STRIDE] = x[q];
    if (p.n_steps) {
        build_tables(tbl0, 0u, ph, group_span(0u, ph));
        prefetch_symbols(group_span(0u, ph));
    }
    __syncthreads();

    while (done < p.n_steps) {
        const uint32_t span = group_span(done, ph);        // steps in this group: phases ph .. ph+span-1
        const bool full = uint32_t(ph) + span == uint32_t(LB);
        uint32_t* drow = dec + size_t(done) * S::T * W;
        uint32_t* xcur = xch0 + size_t(cur) * S::XCH_WORDS;
        uint32_t* xnew = xch0 + size_t(cur ^ 1u) * S::XCH_WORDS;
        const uint4* tcur = tbl0;                          // the set is selected by a bit of pt[]
        uint4* tnew = tbl0 + size_t(cur ^ 1u) * (S::TBL_SET_BYTES / 16);

        // ---- tables of the NEXT group (it starts at phase 0) into the other set - every thread left that set at the barrier of the
        //      previous group, or at the one behind its replay -, symbols of the group after that on their way
        if (full && done + span < p.n_steps) {
            build_tables(tnew, done + span, 0, group_span(done + span, 0));
            prefetch_symbols(done + span + uint32_t(LB));
        }

        // ---- speculative run of the group (no renormalisation).  The running maximum of register 0 is all the trigger test
        //      needs (thread 0 holds state 0 there): if it never reached the threshold, no step of the group renormalised.
        //      The last phase of a complete group stores its results into the other exchange buffer as they are produced.
        uint32_t mx = 0u;
        if (ph == 0 && span == uint32_t(LB)) {
            auto fast_phase = [&](auto PHc) {
                constexpr int PH = decltype(PHc)::value;
                Kn::template step<PH, PH == LB - 1>(x, tcur, pt, drow + size_t(PH) * S::T * W, xnew + wr_off);
                mx = __vmaxu2(mx, x[0]);
            };
            fast_phase(std::integral_constant<int, 0>{});
            if constexpr (LB > 1) fast_phase(std::integral_constant<int, 1>{});
            if constexpr (LB > 2) fast_phase(std::integral_constant<int, 2>{});
            if constexpr (LB > 3) fast_phase(std::integral_constant<int, 3>{});
            if constexpr (LB > 4) fast_phase(std::integral_constant<int, 4>{});
        } else {
            auto run_phase = [&](auto PHc) {
                constexpr int PH = decltype(PHc)::value;
                if (PH >= ph && uint32_t(PH - ph) < span) {
                    Kn::template step<PH>(x, tcur, pt, drow + size_t(PH - ph) * S::T * W);
                    mx = __vmaxu2(mx, x[0]);
                }
            };
            run_phase(std::integral_constant<int, 0>{});
            if constexpr (LB > 1) run_phase(std::integral_constant<int, 1>{});
            if constexpr (LB > 2) run_phase(std::integral_constant<int, 2>{});
            if constexpr (LB > 3) run_phase(std::integral_constant<int, 3>{});
            if constexpr (LB > 4) run_phase(std::integral_constant<int, 4>{});
            if (full) {                                      // a group that started inside an exchange period (streaming API)
#pragma unroll
                for (int q = 0; q < NL; q += 4) *reinterpret_cast<uint4*>(xnew + wr_off + q) = make_uint4(x[q], x[q + 1], x[q + 2], x[q + 3]);
            }
        }
        if (t == 0) {
            bool tb, ta;
            (void)__vibmin_u16x2(p.thr2, mx, &tb, &ta);     // thr <= max over the group of state 0's metric
            flag[cur] = (ta || tb) ? 1u : 0u;
        }
        __syncthreads();                                     // the ONE barrier of a group: steps done, results and next tables in place
        const uint32_t any_trig = flag[cur];                 // (the flag word of the next group is the other one)

        if (any_trig) {
            // ---- roll back and replay step by step with the reference's renormalisation (scalar.h:48-50, 139-153)
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xcur[rd_off + uint32_t(q) * S::RD_STRIDE];
            auto replay_phase = [&](auto PHc) {
                constexpr int PH = decltype(PHc)::value;
                if (PH >= ph && uint32_t(PH - ph) < span) {
                    Kn::template step<PH>(x, tcur, pt, drow + size_t(PH - ph) * S::T * W);
                    if (t == 0) flag[2] = x[0];
                    __syncthreads();
                    const uint32_t x00 = flag[2];
                    bool tb, ta;
                    (void)__vibmin_u16x2(p.thr2, x00, &tb, &ta);
                    if (ta || tb) {                                    // uniform across the CTA
                        const uint32_t m = metric_field<Sat<TIE_SIMD>::value, SH>(Kn::cta_min(x, red));
                        const uint32_t mAv = m & 0xffffu, mBv = m >> 16;
                        const uint32_t sub = (ta ? mAv : 0u) | ((tb ? mBv : 0u) << 16);
                        const uint32_t neg = __vsub2(0u, sub);
#pragma unroll
                        for (int q = 0; q < NL; q++) x[q] = __vadd2(x[q], neg);
                        if (ta) accA += uint64_t(mAv >> SH);
                        if (tb) accB += uint64_t(mBv >> SH);
                    } else {
                        __syncthreads();                               // flag[2] may be rewritten by the next phase
                    }
                }
            };
            replay_phase(std::integral_constant<int, 0>{});
            if constexpr (LB > 1) replay_phase(std::integral_constant<int, 1>{});
            if constexpr (LB > 2) replay_phase(std::integral_constant<int, 2>{});
            if constexpr (LB > 3) replay_phase(std::integral_constant<int, 3>{});
            if constexpr (LB > 4) replay_phase(std::integral_constant<int, 4>{});
            if (full) {
                // the exchange of the replayed group: value at (q, t) moves to PHI' = (t << LB) | q
#pragma unroll
                for (int q = 0; q < NL; q += 4) *reinterpret_cast<uint4*>(xnew + wr_off + q) = make_uint4(x[q], x[q + 1], x[q + 2], x[q + 3]);
            }
            __syncthreads();
        }

        done += span;
        if (full) {
            // ---- exchange, read side: the layout is back to PHI = s.  The buffer just read is the rollback copy of the next group;
            //      the other one is written again only at the end of the next group, behind every thread's loads.
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xnew[rd_off + uint32_t(q) * S::RD_STRIDE];
            cur ^= 1u;
#pragma unroll
            for (int n = 0; n < LB; n++) pt[n] ^= S::TBL_SET_BYTES;
            ph = 0;
        } else {
            ph += int(span);                                 // the call ends inside an exchange period (streaming API)
        }
    }

#pragma unroll
    for (int q = 0; q < NL; q++) {
        const uint32_t s = rotl_bits((uint32_t(q) << LOGT) | t, ph, SB);
        mA[s] = uint16_t((x[q] & 0xffffu) >> SH);
        mB[s] = uint16_t(x[q] >> (16 + SH));
    }
    if (t == 0) { p.acc[fA] = accA; p.acc[fB] = accB; }
}

}  // namespace vitb
