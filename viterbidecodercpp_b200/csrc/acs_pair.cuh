// acs_pair.cuh -- add-compare-select for short constraint lengths (K <= 7): ONE THREAD OWNS A PAIR OF FRAMES.
//
// Replaces ViterbiDecoder_Scalar::update / bfly / renormalise (include/viterbi/viterbi_decoder_scalar.h:28-153) for a batch.
//
// Mapping (B200-first, not a port of the SIMD lanes of x86/viterbi_decoder_avx_u16.h):
//   * a thread holds ALL 2^(K-1) path metrics of two frames in registers: register q = (metric of frame A | metric of frame
//     B << 16).  Every packed instruction (VIADD.16x2, VIMNMX.U16x2) therefore works on two independent frames and the
//     butterfly needs no lane permutation at all - no __shfl, no PRMT.
//   * the state permutation new[2j], new[2j+1] <- old[j], old[j+N/2] is done IN PLACE: after n steps logical state s lives
//     in register rotr^n(s) (rotation of the (K-1)-bit index).  The step loop is unrolled over the K-1 phases so every
//     register index, branch pattern and decision bit position is a compile-time constant.
//   * error_t = uint16_t runs natively in the 16-bit halves (wrapping adds == the scalar reference's modular error_t).
//     error_t = uint8_t is held as metric << 8 in the same halves, so 16-bit wrap-around is exactly 8-bit wrap-around
//     (sm_100a has no native u8x4 add/min: __vaddus4/__vminu4 expand to 7-8 instructions).
//   * compare-select is one VIMNMX.U16x2 with two predicate outputs (a <= b per half).  decision = !(a <= b) = (a > b):
//     the scalar reference's strict '>' tie-break (scalar.h:123-124).  TIE_SIMD swaps the operands to get the SSE/AVX/NEON
//     rule decision = (min == path1) (x86/viterbi_decoder_avx_u16.h:112-115).
//   * decision bits: each predicate feeds one predicated FADD into a float accumulator (2^23 + sum of 2^k): the FP32 pipe is
//     otherwise idle, while VIADD.16x2 (FMA-heavy pipe) and VIMNMX (ALU pipe) are the busy ones
//     (profiles/microbench/r01_pipe_rates_v2.txt).  The mantissa is the decision bit mask.
//   * renormalisation: trigger is state 0's metric >= threshold (scalar.h:48), evaluated per frame half; minimum over the
//     64 registers with VIMNMX3.U16x2; only the triggered half is shifted; the minimum is accumulated in a uint64.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"

namespace vitb {

struct AcsParams {
    const uint32_t* pk;     // packed symbols [n_blocks][n_steps][R][32 lanes]; word = (sA << SH) & 0xffff | (sB << SH) << 16
    void* dec;              // decision rows.  pair kernels: uint64 [n_blocks][dec_rows][64 frames], bit s = decision of state s
                            // (reference bit order); group kernels: uint32 [n_blocks][dec_rows][32 lanes][W] (acs_group.cuh)
    uint16_t* metrics;      // [n_blocks*64][NS] path metrics in logical state order (raw error_t values); in (resume) / out
    uint64_t* acc;          // [n_blocks*64] sum of renormalisation minima; in (resume) / out
    uint32_t n_blocks;      // warp blocks: 64 frames each (pair kernels) or 64/T frames each (group kernels, T lanes per pair)
    uint32_t n_steps;       // trellis steps to run in this launch
    uint32_t dec_rows;      // rows allocated per frame in `dec`
    uint32_t dec_row0;      // row index of this launch's first step (= m_current_decoded_bit)
    uint32_t resume;        // 0: metrics := initial errors (core.h:202-211); 1: continue from `metrics`/`acc`
    uint32_t start_state;
    uint32_t c_low2;        // per half: (-low) << SH                  -> e_low  = s - low   = S + c_low2
    uint32_t c_high2;       // per half: (high << SH) + 1              -> e_high = high - s  = ~S + c_high2
    uint32_t c_inv2;        // per half: (max_error - R*(high-low)) << SH  (0 for every reference preset)
    uint32_t thr2;          // per half: renormalisation_threshold << SH
    uint32_t init_start2;   // per half: initial_start_error << SH
    uint32_t init_other2;   // per half: initial_non_start_error << SH
    // direct-symbol variant of the pair kernel (no ingest pass): the caller's rows, read where they lie
    const void* sym;        // [n_frames][sym_row_bytes] soft_t, 4-byte aligned rows
    uint64_t sym_row_bytes;
    uint64_t sym_total_bytes;   // n_frames * sym_row_bytes (loads are clamped to stay inside)
    uint32_t n_frames;
    // acs_cta.cuh, sliding-window calls: the launch covers steps pk_step0 .. pk_step0 + n_steps - 1 of frames whose packed stream
    // holds pk_steps steps per pair (0 = the launch covers the whole stream: n_steps steps from step 0)
    uint32_t pk_steps, pk_step0;
    uint32_t max_err2;      // per half: soft_decision_max_error << SH (the saturating flavour forms inverted = max_error -sat total)
};

template <class C>
struct PairShape {
    static constexpr int NS = C::NS;
    static constexpr int NACC = NS >= 16 ? NS / 16 : 1;   // float accumulators per frame (16 decision bits each)
};

// ---- branch metric table for one step ------------------------------------------------------------------------------
// T[p] = sum_i (bit_i(p) ? e_high_i : e_low_i), all in packed wrapping u16 arithmetic.  e_low_i = s_i - low,
// e_high_i = high - s_i: the value of |branch - s| for s in [low, high], which is the reference's documented input range
// (scalar.h:30-34).
template <int R, int LEVEL>
struct TableBuild {
    // fills T[0 .. 2^LEVEL) using symbols 0..LEVEL-1
    template <int NP>
    static __device__ __forceinline__ void run(uint32_t (&T)[NP], const uint32_t (&lo)[R], const uint32_t (&hi)[R]) {
        TableBuild<R, LEVEL - 1>::run(T, lo, hi);
        constexpr int H = 1 << (LEVEL - 1);
#pragma unroll
        for (int p = 0; p < H; p++) {
            T[p + H] = __vadd2(T[p], hi[LEVEL - 1]);
            T[p] = __vadd2(T[p], lo[LEVEL - 1]);
        }
    }
};
template <int R>
struct TableBuild<R, 1> {
    template <int NP>
    static __device__ __forceinline__ void run(uint32_t (&T)[NP], const uint32_t (&lo)[R], const uint32_t (&hi)[R]) {
        T[0] = lo[0];
        T[1] = hi[0];
    }
};

// ---- saturating flavour (VITB_TIE_SIMD_SAT, template value TIE_SIMD == 2) ------------------------------------------------------------
// The reference's SSE / AVX / NEON decoders add path metrics with saturating instructions (x86/viterbi_decoder_avx_u16.h:107-110
// adds_epu16, avx_u8.h:107-110 adds_epu8), form inverted_error = max_error -sat total_error (:106) and accumulate the branch errors with
// saturating adds too (:97); a tie - including two saturated candidates - selects path 1 (:112-115).  This flavour reproduces all of
// that in the decision-row kernels.  uint16_t metrics: __vaddus2 on the native halves.  uint8_t metrics are held as metric << 8; in
// this flavour the LOW BYTE OF EVERY METRIC IS 0xFF, so that a 16-bit saturating add of (m << 8 | 0xFF) + (e << 8) saturates to
// 0xFFFF = (0xFF << 8 | 0xFF) exactly when m + e >= 256: the representation is closed under add, min and the subtraction of a
// minimum whose low byte was cleared.  Branch errors keep a zero low byte.
template <int TIE_SIMD>
struct Sat { static constexpr bool value = (TIE_SIMD == 2); };
template <bool SAT>
__device__ __forceinline__ uint32_t metric_add(uint32_t a, uint32_t b) { return SAT ? __vaddus2(a, b) : __vadd2(a, b); }
template <bool SAT, int SH>
__device__ __forceinline__ uint32_t metric_ones(uint32_t x) { return (SAT && SH == 8) ? (x | 0x00ff00ffu) : x; }      // metric representation
template <bool SAT, int SH>
__device__ __forceinline__ uint32_t metric_field(uint32_t x) { return (SAT && SH == 8) ? (x & 0xff00ff00u) : x; }     // value with a zero low byte

// branch metric table with saturating accumulation of the R errors (avx_u16.h:95-97); entries keep a zero low byte
template <int R, int SH, int NP>
__device__ __forceinline__ void table_build_sat(uint32_t (&T)[NP], const uint32_t (&lo)[R], const uint32_t (&hi)[R]) {
#pragma unroll
    for (int pat = 0; pat < NP; pat++) {
        uint32_t acc = 0u;
#pragma unroll
        for (int i = 0; i < R; i++) {
            acc = __vaddus2(acc, ((pat >> i) & 1) ? hi[i] : lo[i]);
            if (SH == 8) acc &= 0xff00ff00u;
        }
        T[pat] = acc;
    }
}

// ---- one butterfly, everything about its position known at compile time ----------------------------------------------
template <class C, int PH, int TIE_SIMD, int Q>
__device__ __forceinline__ void bfly_at(uint32_t (&x)[C::NS], const uint32_t (&T)[C::NP], float (&fa)[2][PairShape<C>::NACC],
                                        const uint32_t c_inv2, const bool consistent, const uint32_t max_err2) {
    constexpr int SB = C::SB;
    constexpr bool SAT = Sat<TIE_SIMD>::value;
    constexpr int bit = 1 << (SB - 1 - PH);
    if constexpr ((Q & bit) == 0) {
        constexpr int q0 = Q, q1 = Q | bit;
        constexpr uint32_t j = rotl_bits(uint32_t(q0), PH, SB);     // logical old state (leading bit 0)
        static_assert(j < uint32_t(C::NS / 2), "butterfly index must have leading bit 0");
        constexpr uint32_t pat = bfly_pattern<C>(j);
        constexpr uint32_t ipat = (~pat) & uint32_t(C::NP - 1);
        constexpr uint32_t s0 = 2 * j, s1 = 2 * j + 1;              // logical new states -> decision bit positions
        const uint32_t tot = T[pat];
        // inverted_error = max_error - total_error (scalar.h:107) = T[~pat] + c_inv2
        const uint32_t inv = consistent ? T[ipat] : (SAT ? __vsubus2(max_err2, tot) : __vadd2(T[ipat], c_inv2));
        const uint32_t a0 = metric_add<SAT>(x[q0], tot);   // (0|X) -> (X|0)   scalar.h:113
        const uint32_t b0 = metric_add<SAT>(x[q1], inv);   // (1|X) -> (X|0)   scalar.h:114
        const uint32_t a1 = metric_add<SAT>(x[q0], inv);   // (0|X) -> (X|1)   scalar.h:115
        const uint32_t b1 = metric_add<SAT>(x[q1], tot);   // (1|X) -> (X|1)   scalar.h:116
        bool h0, l0, h1, l1;
        bool dA0, dB0, dA1, dB1;
        if constexpr (!TIE_SIMD) {
            x[q0] = __vibmin_u16x2(a0, b0, &h0, &l0);   // pred = (a <= b); decision = a > b
            x[q1] = __vibmin_u16x2(a1, b1, &h1, &l1);
            dA0 = !l0; dB0 = !h0; dA1 = !l1; dB1 = !h1;
        } else {
            x[q0] = __vibmin_u16x2(b0, a0, &h0, &l0);   // pred = (b <= a) = (min == b)
            x[q1] = __vibmin_u16x2(b1, a1, &h1, &l1);
            dA0 = l0; dB0 = h0; dA1 = l1; dB1 = h1;
        }
        constexpr int acc0 = int(s0 >> 4) % PairShape<C>::NACC, acc1 = int(s1 >> 4) % PairShape<C>::NACC;
        constexpr float w0 = float(1u << (s0 & 15)), w1 = float(1u << (s1 & 15));
        if (dA0) fa[0][acc0] += w0;
        if (dB0) fa[1][acc0] += w0;
        if (dA1) fa[0][acc1] += w1;
        if (dB1) fa[1][acc1] += w1;
    }
}

template <class C, int PH, int TIE_SIMD, int... Qs>
__device__ __forceinline__ void bfly_all(uint32_t (&x)[C::NS], const uint32_t (&T)[C::NP], float (&fa)[2][PairShape<C>::NACC],
                                         const uint32_t c_inv2, const bool consistent, const uint32_t max_err2,
                                         std::integer_sequence<int, Qs...>) {
    (bfly_at<C, PH, TIE_SIMD, Qs>(x, T, fa, c_inv2, consistent, max_err2), ...);
}

// ---- tagged butterfly (uint8_t metrics only) --------------------------------------------------------------------------
// With metric << 8 in each half the low byte of every register is free.  Give the inverted-error operand (path 1) a low byte
// of 1 and the total-error operand (path 0) a low byte of 0: min() then breaks a metric tie in favour of path 0 - exactly the
// reference's strict '>' (scalar.h:123-124) - and the low byte of the minimum IS the decision bit.  No predicates, so add and min
// fuse into VIADDMNMX.U16x2 and the decision bits of two registers (4 frames x states) are collected with one PRMT and shifted
// into a byte-lane accumulator with one 3-input add: 8 instructions per butterfly instead of 10, and spread over both integer
// pipes (profiles/microbench/r01_acs_decision_mix.txt: the predicated FADDs cost more than the ACS arithmetic itself).
// TIE_SIMD puts the tag on path 0 instead and the collected bits are inverted at the end.
// T = total error by pattern, V = inverted error by pattern's complement index (V[k] = T[k] + c_inv), TT / VT = the same with the
// tag byte set.  For consistent configurations V aliases T and VT aliases TT.
template <class C, int PH, int TIE_SIMD, int J>
__device__ __forceinline__ void tag_bfly_at(uint32_t (&x)[C::NS], const uint32_t (&T)[C::NP], const uint32_t (&TT)[C::NP],
                                            const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP],
                                            uint32_t (&dacc)[(C::NS / 2 + 7) / 8]) {
    constexpr int SB = C::SB;
    constexpr int q0 = int(rotr_bits(uint32_t(J), PH, SB));                 // register of logical old state J (leading bit 0)
    constexpr int q1 = q0 | (1 << (SB - 1 - PH));
    static_assert((q0 & (1 << (SB - 1 - PH))) == 0, "butterfly register must have the phase bit clear");
    constexpr uint32_t pat = bfly_pattern<C>(uint32_t(J));
    constexpr uint32_t ipat = (~pat) & uint32_t(C::NP - 1);
    uint32_t m0, m1;
    if constexpr (!TIE_SIMD) {
        const uint32_t b0 = __vadd2(x[q1], VT[ipat]);                       // (1|X) + inverted, tag 1     scalar.h:114
        const uint32_t b1 = __vadd2(x[q1], TT[pat]);                        // (1|X) + total,    tag 1     scalar.h:116
        m0 = __viaddmin_u16x2(x[q0], T[pat], b0);                           // min((0|X) + total, b0)      scalar.h:113,127
        m1 = __viaddmin_u16x2(x[q0], V[ipat], b1);                          // min((0|X) + inverted, b1)   scalar.h:115,128
    } else {
        const uint32_t a0 = __vadd2(x[q0], TT[pat]);                        // tag on path 0: a tie selects path 1 (avx_u16.h:112-115)
        const uint32_t a1 = __vadd2(x[q0], VT[ipat]);
        m0 = __viaddmin_u16x2(x[q1], V[ipat], a0);
        m1 = __viaddmin_u16x2(x[q1], T[pat], a1);
    }
    // bytes: [A: state 2J, B: state 2J, A: state 2J+1, B: state 2J+1]; bit 7-(J%8) of each byte lane after 8 butterflies
    dacc[J / 8] = dacc[J / 8] + dacc[J / 8] + __byte_perm(m0, m1, 0x6420);
    x[q0] = m0 & 0xff00ff00u;
    x[q1] = m1 & 0xff00ff00u;
}

template <class C, int PH, int TIE_SIMD, int... Js>
__device__ __forceinline__ void tag_bfly_all(uint32_t (&x)[C::NS], const uint32_t (&T)[C::NP], const uint32_t (&TT)[C::NP],
                                             const uint32_t (&V)[C::NP], const uint32_t (&VT)[C::NP],
                                             uint32_t (&dacc)[(C::NS / 2 + 7) / 8], std::integer_sequence<int, Js...>) {
    (tag_bfly_at<C, PH, TIE_SIMD, Js>(x, T, TT, V, VT, dacc), ...);
}

// minimum over all registers, per half
template <int NS>
__device__ __forceinline__ uint32_t packed_min(const uint32_t (&x)[NS]) {
    uint32_t m = x[0];
#pragma unroll
    for (int q = 1; q + 1 < NS; q += 2) m = __vimin3_u16x2(m, x[q], x[q + 1]);
    if (NS % 2 == 0) m = __vminu2(m, x[NS - 1]);
    return m;
}

// ---- one trellis step at compile-time phase PH ---------------------------------------------------------------------
template <class C, int SH, int TIE_SIMD, bool CONSISTENT, int PH, bool TAG = (SH == 8 && TIE_SIMD != 2)>
__device__ __forceinline__ void acs_pair_step(uint32_t (&x)[C::NS], const uint32_t* sym /* R packed words */, const AcsParams& p,
                                              uint64_t* dec_row /* this lane's 16 bytes of the row */, uint64_t& accA, uint64_t& accB) {
    constexpr int R = C::R, NP = C::NP, NS = C::NS, NACC = PairShape<C>::NACC;
    static_assert(!TAG || SH == 8, "the tagged butterfly needs the free low byte of uint8_t metrics held as metric << 8");
    // per-symbol errors against low / high
    uint32_t lo[R], hi[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        lo[i] = __vadd2(sym[i], p.c_low2);
        hi[i] = __vadd2(~sym[i], p.c_high2);
    }
    constexpr bool SAT = Sat<TIE_SIMD>::value;
    uint32_t T[NP];
    if constexpr (SAT) table_build_sat<R, SH>(T, lo, hi);
    else TableBuild<R, R>::run(T, lo, hi);

    uint32_t wA0, wA1 = 0, wB0, wB1 = 0;
    if constexpr (TAG) {
        uint32_t TT[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) TT[k] = __vadd2(T[k], 0x00010001u);
        constexpr int NDA = (NS / 2 + 7) / 8;
        uint32_t dacc[NDA];
#pragma unroll
        for (int k = 0; k < NDA; k++) dacc[k] = 0u;
        if constexpr (CONSISTENT) {
            tag_bfly_all<C, PH, TIE_SIMD>(x, T, TT, T, TT, dacc, std::make_integer_sequence<int, NS / 2>{});
        } else {
            uint32_t V[NP], VT[NP];       // inverted_error = max_error - total_error (scalar.h:107) = T[~pattern] + c_inv
#pragma unroll
            for (int k = 0; k < NP; k++) { V[k] = __vadd2(T[k], p.c_inv2); VT[k] = __vadd2(TT[k], p.c_inv2); }
            tag_bfly_all<C, PH, TIE_SIMD>(x, T, TT, V, VT, dacc, std::make_integer_sequence<int, NS / 2>{});
        }
        if constexpr (TIE_SIMD) {
#pragma unroll
            for (int k = 0; k < NDA; k++) dacc[k] = ~dacc[k];         // the tag marked path 0: decision = !tag
        }
        // tagged row layout (LAYOUT_PAIR_TAG): byte 2k+h of a frame's 64-bit word holds the decisions of states 2J+h, J = 8k..8k+7,
        // first butterfly in the top bit.  Butterfly groups that do not exist (K < 7) leave zero bytes.
        if constexpr (NDA == 4) {
            wA0 = __byte_perm(dacc[0], dacc[1], 0x6420); wA1 = __byte_perm(dacc[2], dacc[3], 0x6420);
            wB0 = __byte_perm(dacc[0], dacc[1], 0x7531); wB1 = __byte_perm(dacc[2], dacc[3], 0x7531);
        } else if constexpr (NDA == 2) {
            wA0 = __byte_perm(dacc[0], dacc[1], 0x6420); wB0 = __byte_perm(dacc[0], dacc[1], 0x7531);
        } else {
            wA0 = __byte_perm(dacc[0], 0u, 0x4420); wB0 = __byte_perm(dacc[0], 0u, 0x4431);
        }
        if constexpr (NS / 2 < 8) {    // fewer than 8 butterflies: the bits sit at the bottom of the byte, move them to the top
            wA0 <<= (8 - NS / 2); wB0 <<= (8 - NS / 2);
        }
    } else {
    float fa[2][NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) { fa[0][a] = 8388608.f; fa[1][a] = 8388608.f; }

    bfly_all<C, PH, TIE_SIMD>(x, T, fa, p.c_inv2, CONSISTENT, p.max_err2, std::make_integer_sequence<int, NS>{});

    // decision row: bit s of the 64-bit word = decision of logical state s (same bit order as core.h:49-83 / scalar.h:131-134)
    if constexpr (NACC == 4) {
        wA0 = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[0][1]), 0x5410);
        wA1 = __byte_perm(__float_as_uint(fa[0][2]), __float_as_uint(fa[0][3]), 0x5410);
        wB0 = __byte_perm(__float_as_uint(fa[1][0]), __float_as_uint(fa[1][1]), 0x5410);
        wB1 = __byte_perm(__float_as_uint(fa[1][2]), __float_as_uint(fa[1][3]), 0x5410);
    } else if constexpr (NACC == 2) {
        wA0 = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[0][1]), 0x5410);
        wB0 = __byte_perm(__float_as_uint(fa[1][0]), __float_as_uint(fa[1][1]), 0x5410);
    } else {
        wA0 = __float_as_uint(fa[0][0]) & 0xffffu;
        wB0 = __float_as_uint(fa[1][0]) & 0xffffu;
    }
    }
    *reinterpret_cast<uint4*>(dec_row) = make_uint4(wA0, wA1, wB0, wB1);

    // renormalisation: "if (new_metric[0] >= renormalisation_threshold)" (scalar.h:48) - state 0 always sits in register 0
    bool trigB, trigA;
    (void)__vibmin_u16x2(p.thr2, x[0], &trigB, &trigA);      // pred = thr <= x0
    if (trigA || trigB) {
        const uint32_t m = metric_field<SAT, SH>(packed_min<NS>(x));   // scalar.h:140-146
        const uint32_t mA = m & 0xffffu, mB = m >> 16;
        const uint32_t sub = (trigA ? mA : 0u) | ((trigB ? mB : 0u) << 16);
        const uint32_t neg = __vsub2(0u, sub);
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = __vadd2(x[q], neg);  // scalar.h:148-150
        if (trigA) accA += uint64_t(mA >> SH);                  // scalar.h:49, 152
        if (trigB) accB += uint64_t(mB >> SH);
    }
}

// Number of in-place phases unrolled before the registers are renamed back to logical order (PHI = s).  SB phases need no
// renaming at all but make the hot loop SB steps long (36 KB of SASS for K=7: instruction-fetch bound, "no_instruction" is the top
// stall in profiles/r01_ncu_full_acs_pair_cfg2.txt); a shorter period trades 2^(K-1) register moves per period for a loop that
// fits the instruction cache.
template <class C>
struct PairPeriod { static constexpr int value = (C::SB == 6) ? 3 : C::SB; };

template <class C, int SH, int TIE_SIMD, bool CONSISTENT, int PERIOD = PairPeriod<C>::value>
struct PairRunner {
    static constexpr int P = PERIOD, R = C::R, NS = C::NS;

    template <int PH>
    static __device__ __forceinline__ bool phase(uint32_t (&x)[NS], const uint32_t (&cur)[P * R], const AcsParams& p, uint32_t t0,
                                                 uint64_t* dec_lane, uint64_t& accA, uint64_t& accB) {
        if (t0 + PH >= p.n_steps) return false;
        acs_pair_step<C, SH, TIE_SIMD, CONSISTENT, PH>(x, &cur[PH * R], p, dec_lane + size_t(t0 + PH) * 64, accA, accB);
        return true;
    }

    template <int... PHs>
    static __device__ __forceinline__ void group(uint32_t (&x)[NS], const uint32_t (&cur)[P * R], const AcsParams& p, uint32_t t0,
                                                 uint64_t* dec_lane, uint64_t& accA, uint64_t& accB, std::integer_sequence<int, PHs...>) {
        (void)(phase<PHs>(x, cur, p, t0, dec_lane, accA, accB) && ...);
    }
};

// One warp per 64-frame block (lane l owns frames 64*blk + 2l and 64*blk + 2l + 1); PAIR_WARPS warps per CTA so that the warps of
// a CTA land on all four SM sub-partitions evenly.  grid = ceil(n_blocks / PAIR_WARPS).
constexpr int PAIR_WARPS = 4;

// ---- direct symbol fetch ----------------------------------------------------------------------------------------------
// Each lane streams its own two rows of the caller's [frame][step][R] array (the reference's layout, scalar.h:43-46) with 32-bit
// loads, one period of steps ahead; a 128-byte line of a row serves ~20 periods from L1/L2, so DRAM traffic stays at the size of
// the input.  Saves the ingest pass (one full read of the symbols plus a 2x larger write and re-read of the packed stream).
// Requires 4-byte aligned rows and an even number of bytes per period; anything else (and punctured input) goes through ingest.
template <class C, int SH, int P>
struct DirectFetch {
    static constexpr int SBY = (SH == 8) ? 1 : 2;       // bytes per soft symbol (int8_t with uint8_t metrics, else int16_t)
    static constexpr int BS = C::R * SBY;               // bytes per trellis step
    static constexpr int PB = P * BS;                   // bytes per period
    static constexpr bool supported = (PB % 2) == 0;
    static constexpr int NWA = (PB + 3) / 4;            // aligned words holding one period
    static constexpr int NW = NWA + ((PB % 4) ? 1 : 0); // words loaded: one more when periods alternate between offsets 0 and 2

    static __device__ __forceinline__ void load(uint32_t (&raw)[NW], const uint32_t* row_words, uint32_t period, uint32_t max_word) {
        const uint32_t w0 = (period * uint32_t(PB)) >> 2;
#pragma unroll
        for (int j = 0; j < NW; j++) {
            const uint32_t w = w0 + uint32_t(j);
            raw[j] = __ldg(row_words + (w < max_word ? w : max_word));
        }
    }
    // packed symbol words of the period: out[n * R + i] = (sA << SH) & 0xffff | (sB << SH) << 16
    static __device__ __forceinline__ void convert(uint32_t (&out)[P * C::R], const uint32_t (&rawA)[NW], const uint32_t (&rawB)[NW],
                                                   uint32_t period) {
        uint32_t a[NWA], b[NWA];
        if constexpr ((PB % 4) != 0) {
            const uint32_t sh = ((period * uint32_t(PB)) & 2u) * 8u;      // 0 or 16
#pragma unroll
            for (int j = 0; j < NWA; j++) { a[j] = __funnelshift_r(rawA[j], rawA[j + 1], sh); b[j] = __funnelshift_r(rawB[j], rawB[j + 1], sh); }
        } else {
#pragma unroll
            for (int j = 0; j < NWA; j++) { a[j] = rawA[j]; b[j] = rawB[j]; }
        }
#pragma unroll
        for (int n = 0; n < P; n++) {
#pragma unroll
            for (int i = 0; i < C::R; i++) {
                const int bp = n * BS + i * SBY, w = bp >> 2, o = bp & 3;
                if constexpr (SBY == 1) {
                    // byte o of A -> byte 1, byte o of B -> byte 3, low bytes cleared (the metrics are held as value << 8)
                    out[n * C::R + i] = __byte_perm(a[w], b[w], uint32_t(((4 + o) << 12) | (o << 4))) & 0xff00ff00u;
                } else {
                    out[n * C::R + i] = __byte_perm(a[w], b[w], o ? 0x7632u : 0x5410u);
                }
            }
        }
    }
};

template <class C, int SH, int TIE_SIMD, bool CONSISTENT, int PERIOD = PairPeriod<C>::value, bool DIRECT = false>
__global__ void __launch_bounds__(32 * PAIR_WARPS) acs_pair_kernel(const AcsParams p) {
    constexpr int P = PERIOD, R = C::R, NS = C::NS, SB = C::SB;
    using Run = PairRunner<C, SH, TIE_SIMD, CONSISTENT, PERIOD>;
    using DF = DirectFetch<C, SH, P>;
    const uint32_t lane = threadIdx.x & 31, blk = blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
    if (blk >= p.n_blocks) return;
    const size_t fA = size_t(blk) * 64 + 2 * lane, fB = fA + 1;

    uint32_t x[NS];
    uint64_t accA = 0, accB = 0;
    if (p.resume) {
        const uint16_t* mA = p.metrics + fA * NS;
        const uint16_t* mB = p.metrics + fB * NS;
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = ((uint32_t(mA[q]) << SH) & 0xffffu) | (uint32_t(mB[q]) << (16 + SH));
        accA = p.acc[fA];
        accB = p.acc[fB];
    } else {
        const uint32_t s = p.start_state & uint32_t(NS - 1);       // core.h:209-210
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = (uint32_t(q) == s) ? p.init_start2 : p.init_other2;
    }
#pragma unroll
    for (int q = 0; q < NS; q++) x[q] = metric_ones<Sat<TIE_SIMD>::value, SH>(x[q]);

    const uint32_t* pk = p.pk + size_t(blk) * p.n_steps * R * 32 + lane;
    uint64_t* dec_lane = static_cast<uint64_t*>(p.dec) + (size_t(blk) * p.dec_rows + p.dec_row0) * 64 + 2 * lane;

    uint32_t cur[P * R], nxt[DIRECT ? 1 : P * R];
    uint32_t rawA[DIRECT ? DF::NW : 1], rawB[DIRECT ? DF::NW : 1];
    const uint32_t* rowA = nullptr;
    const uint32_t* rowB = nullptr;
    uint32_t maxwA = 0, maxwB = 0;
    if constexpr (DIRECT) {
        static_assert(!DIRECT || DF::supported, "direct symbol fetch needs an even number of bytes per period");
        const size_t lastf = size_t(p.n_frames) - 1;
        const size_t ldA = fA < lastf ? fA : lastf, ldB = fB < lastf ? fB : lastf;     // padding lanes re-read the last frame
        rowA = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ldA * p.sym_row_bytes);
        rowB = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.sym) + ldB * p.sym_row_bytes);
        maxwA = clamp_words_left(p.sym_total_bytes, ldA * p.sym_row_bytes);
        maxwB = clamp_words_left(p.sym_total_bytes, ldB * p.sym_row_bytes);
        DF::load(rawA, rowA, 0u, maxwA);
        DF::load(rawB, rowB, 0u, maxwB);
    } else {
#pragma unroll
        for (int k = 0; k < P * R; k++) nxt[k] = (uint32_t(k / R) < p.n_steps) ? __ldg(pk + size_t(k) * 32) : 0u;
    }

    uint32_t period = 0;
    for (uint32_t t0 = 0; t0 < p.n_steps; t0 += P, period++) {
        const uint32_t tn = t0 + P;
        if constexpr (DIRECT) {
            DF::convert(cur, rawA, rawB, period);
            if (tn < p.n_steps) { DF::load(rawA, rowA, period + 1, maxwA); DF::load(rawB, rowB, period + 1, maxwB); }
        } else {
#pragma unroll
            for (int k = 0; k < P * R; k++) cur[k] = nxt[k];
#pragma unroll
            for (int k = 0; k < P * R; k++) nxt[k] = (tn + uint32_t(k / R) < p.n_steps) ? __ldg(pk + (size_t(tn) * R + k) * 32) : 0u;
        }
        Run::group(x, cur, p, t0, dec_lane, accA, accB, std::make_integer_sequence<int, P>{});
        if constexpr (P < SB) {
            if (t0 + P <= p.n_steps) {          // a full period ran: state s sits in register rotr^P(s); rename back to register s
                uint32_t y[NS];
#pragma unroll
                for (int q = 0; q < NS; q++) y[q] = x[rotr_bits(uint32_t(q), P, SB)];
#pragma unroll
                for (int q = 0; q < NS; q++) x[q] = y[q];
            }
        }
    }

    // after n steps (since the last renaming) logical state s sits in register rotr^n(s); write back in logical order
    // (core.h:195-199 reads old_metrics[end_state])
    const int ph = int(p.n_steps % uint32_t(P));
    uint16_t* mA = p.metrics + fA * NS;
    uint16_t* mB = p.metrics + fB * NS;
#pragma unroll
    for (int q = 0; q < NS; q++) {
        const uint32_t s = rotl_bits(uint32_t(q), ph, SB);
        mA[s] = uint16_t((x[q] & 0xffffu) >> SH);
        mB[s] = uint16_t(x[q] >> (16 + SH));
    }
    p.acc[fA] = accA;
    p.acc[fB] = accB;
}

}  // namespace vitb
