// acs_generic.cuh -- add-compare-select for ANY generator polynomials (3 <= K <= 7, R <= 6): the fallback for codes outside the
// compiled catalogue (examples/helpers/common_codes.h), so that ViterbiBranchTable<K,R,soft_t>(G, high, low) with arbitrary G
// (include/viterbi/viterbi_branch_table.h:33-55) has a CUDA backend too.
//
// Same arithmetic, semantics and decision-row layout as acs_pair.cuh (one thread owns a frame pair, metrics packed 16x2, predicated
// FADD decision accumulators, bit s of a frame's 64-bit row word = decision of state s), so ingest, traceback, the streaming calls
// and the result gather are shared.  What differs: G is not a template parameter, so
//   * the branch pattern of butterfly j comes from a 32-entry table in the kernel parameters (filled by the host from G with
//     bfly_pattern_rt, the same parity((j << 1) & G[i]) rule as viterbi_branch_table.h:45-54); after unrolling, j is a compile-time
//     index into that table: a constant-bank operand, no dynamic register indexing;
//   * the per-step branch metric table T[2^R] cannot live in registers (its index is only known at run time): every thread keeps
//     its column in shared memory ([pattern][thread], conflict-free) and a butterfly fetches total / inverted error with two LDS;
//   * R is a run-time value (the table is built with a doubling loop), so one instantiation per (K, metric type, tie-break) serves
//     every rate;
//   * metrics ping-pong between two register sets in logical order (no in-place rotation: its register names depend on K only).
// About 14 instructions per butterfly instead of 10 (catalogue kernels) or 4 (survivor-history kernels): a correct, reasonably fast
// fallback, not a tuned path.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"

namespace vitb {

constexpr int GENERIC_MAX_R = 6;
constexpr int GENERIC_THREADS = 64;                  // 2 warps per CTA: 64 patterns x 64 threads x 4 B = 16 KB of shared memory

struct GenericCode {
    uint32_t R;
    uint8_t pat[32];          // branch pattern of butterfly j (old state with leading bit 0), j < 2^(K-2)
};

template <int K, int TIE_SIMD, int J>
__device__ __forceinline__ void gen_bfly_at(const uint32_t (&x)[1 << (K - 1)], uint32_t (&y)[1 << (K - 1)], const uint32_t* Tcol,
                                            const GenericCode& gc, const uint32_t np_mask, const uint32_t c_inv2,
                                            float (&fa)[2][(1 << (K - 1)) >= 16 ? (1 << (K - 1)) / 16 : 1]) {
    constexpr int NS = 1 << (K - 1), H = NS / 2, NACC = NS >= 16 ? NS / 16 : 1;
    const uint32_t pat = gc.pat[J], ipat = (~pat) & np_mask;
    const uint32_t tot = Tcol[pat * GENERIC_THREADS];
    const uint32_t inv = __vadd2(Tcol[ipat * GENERIC_THREADS], c_inv2);      // inverted_error = max_error - total_error (scalar.h:107)
    const uint32_t a0 = __vadd2(x[J], tot), b0 = __vadd2(x[J + H], inv);     // scalar.h:113-114
    const uint32_t a1 = __vadd2(x[J], inv), b1 = __vadd2(x[J + H], tot);     // scalar.h:115-116
    constexpr int s0 = 2 * J, s1 = 2 * J + 1;
    bool h0, l0, h1, l1, dA0, dB0, dA1, dB1;
    if constexpr (!TIE_SIMD) {
        y[s0] = __vibmin_u16x2(a0, b0, &h0, &l0);       // pred = (a <= b); decision = a > b (scalar.h:123-124)
        y[s1] = __vibmin_u16x2(a1, b1, &h1, &l1);
        dA0 = !l0; dB0 = !h0; dA1 = !l1; dB1 = !h1;
    } else {
        y[s0] = __vibmin_u16x2(b0, a0, &h0, &l0);       // pred = (b <= a) = (min == path 1), avx_u16.h:112-115
        y[s1] = __vibmin_u16x2(b1, a1, &h1, &l1);
        dA0 = l0; dB0 = h0; dA1 = l1; dB1 = h1;
    }
    constexpr int acc0 = (s0 >> 4) % NACC, acc1 = (s1 >> 4) % NACC;
    constexpr float w0 = float(1u << (s0 & 15)), w1 = float(1u << (s1 & 15));
    if (dA0) fa[0][acc0] += w0;
    if (dB0) fa[1][acc0] += w0;
    if (dA1) fa[0][acc1] += w1;
    if (dB1) fa[1][acc1] += w1;
}

template <int K, int TIE_SIMD, int... Js>
__device__ __forceinline__ void gen_bfly_all(const uint32_t (&x)[1 << (K - 1)], uint32_t (&y)[1 << (K - 1)], const uint32_t* Tcol,
                                             const GenericCode& gc, const uint32_t np_mask, const uint32_t c_inv2,
                                             float (&fa)[2][(1 << (K - 1)) >= 16 ? (1 << (K - 1)) / 16 : 1], std::integer_sequence<int, Js...>) {
    (gen_bfly_at<K, TIE_SIMD, Js>(x, y, Tcol, gc, np_mask, c_inv2, fa), ...);
}

// one trellis step x -> y; sym = this step's R packed symbol words
template <int K, int SH, int TIE_SIMD>
__device__ __forceinline__ void gen_step(const uint32_t (&x)[1 << (K - 1)], uint32_t (&y)[1 << (K - 1)], const uint32_t* sym, uint32_t* Tcol,
                                         const AcsParams& p, const GenericCode& gc, uint64_t* dec_row, uint64_t& accA, uint64_t& accB) {
    constexpr int NS = 1 << (K - 1), NACC = NS >= 16 ? NS / 16 : 1;
    const uint32_t R = gc.R;
    // T[pattern] = sum_i (bit_i ? high - s_i : s_i - low), built by doubling (this thread's column only: no synchronisation needed)
    Tcol[0] = 0u;
#pragma unroll
    for (int i = 0; i < GENERIC_MAX_R; i++) {          // unrolled so that sym[i] stays a register; the inner loop runs on shared memory
        if (uint32_t(i) < R) {
            const uint32_t lo = __vadd2(sym[i], p.c_low2), hi = __vadd2(~sym[i], p.c_high2);
            const uint32_t half = 1u << i;
            for (uint32_t q = 0; q < half; q++) {
                const uint32_t t = Tcol[q * GENERIC_THREADS];
                Tcol[(q + half) * GENERIC_THREADS] = __vadd2(t, hi);
                Tcol[q * GENERIC_THREADS] = __vadd2(t, lo);
            }
        }
    }
    float fa[2][NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) { fa[0][a] = 8388608.f; fa[1][a] = 8388608.f; }
    gen_bfly_all<K, TIE_SIMD>(x, y, Tcol, gc, (1u << R) - 1u, p.c_inv2, fa, std::make_integer_sequence<int, NS / 2>{});

    // decision row: bit s of the 64-bit word = decision of logical state s (core.h:49-83 / scalar.h:131-134)
    uint32_t wA0, wA1 = 0, wB0, wB1 = 0;
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
    *reinterpret_cast<uint4*>(dec_row) = make_uint4(wA0, wA1, wB0, wB1);

    bool trigB, trigA;
    (void)__vibmin_u16x2(p.thr2, y[0], &trigB, &trigA);      // scalar.h:48
    if (trigA || trigB) {
        const uint32_t m = packed_min<NS>(y);                 // scalar.h:140-146
        const uint32_t mA = m & 0xffffu, mB = m >> 16;
        const uint32_t sub = (trigA ? mA : 0u) | ((trigB ? mB : 0u) << 16);
        const uint32_t neg = __vsub2(0u, sub);
#pragma unroll
        for (int q = 0; q < NS; q++) y[q] = __vadd2(y[q], neg);   // scalar.h:148-150
        if (trigA) accA += uint64_t(mA >> SH);                // scalar.h:49, 152
        if (trigB) accB += uint64_t(mB >> SH);
    }
}

// grid = ceil(n_blocks / 2), block = 64 threads (2 warps, one 64-frame block each); packed symbol stream in the 32-pairs-per-warp layout
template <int K, int SH, int TIE_SIMD>
__global__ void __launch_bounds__(GENERIC_THREADS) acs_generic_kernel(const AcsParams p, const GenericCode gc) {
    constexpr int NS = 1 << (K - 1);
    __shared__ uint32_t Tsm[(1 << GENERIC_MAX_R) * GENERIC_THREADS];
    uint32_t* Tcol = Tsm + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, blk = blockIdx.x * (GENERIC_THREADS / 32) + (threadIdx.x >> 5);
    if (blk >= p.n_blocks) return;
    const size_t fA = size_t(blk) * 64 + 2 * lane, fB = fA + 1;
    const uint32_t R = gc.R;

    uint32_t x[NS], y[NS];
    uint64_t accA = 0, accB = 0;
    uint16_t* mA = p.metrics + fA * NS;
    uint16_t* mB = p.metrics + fB * NS;
    if (p.resume) {
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = ((uint32_t(mA[q]) << SH) & 0xffffu) | (uint32_t(mB[q]) << (16 + SH));
        accA = p.acc[fA];
        accB = p.acc[fB];
    } else {
        const uint32_t s = p.start_state & uint32_t(NS - 1);       // core.h:209-210
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = (uint32_t(q) == s) ? p.init_start2 : p.init_other2;
    }

    const uint32_t* pk = p.pk + size_t(blk) * p.n_steps * R * 32 + lane;
    uint64_t* dec_lane = static_cast<uint64_t*>(p.dec) + (size_t(blk) * p.dec_rows + p.dec_row0) * 64 + 2 * lane;

    uint32_t cur[2 * GENERIC_MAX_R], nxt[2 * GENERIC_MAX_R];
#pragma unroll
    for (int k = 0; k < 2 * GENERIC_MAX_R; k++) nxt[k] = 0u;
#pragma unroll
    for (int n = 0; n < 2; n++) {
#pragma unroll
        for (int i = 0; i < GENERIC_MAX_R; i++)
            if (uint32_t(i) < R && uint32_t(n) < p.n_steps) nxt[n * GENERIC_MAX_R + i] = __ldg(pk + (size_t(n) * R + i) * 32);
    }

    uint32_t t = 0;
#pragma unroll 1
    for (; t + 2 <= p.n_steps; t += 2) {
#pragma unroll
        for (int k = 0; k < 2 * GENERIC_MAX_R; k++) cur[k] = nxt[k];
#pragma unroll
        for (int n = 0; n < 2; n++) {
#pragma unroll
            for (int i = 0; i < GENERIC_MAX_R; i++)
                if (uint32_t(i) < R && t + 2 + uint32_t(n) < p.n_steps) nxt[n * GENERIC_MAX_R + i] = __ldg(pk + (size_t(t + 2 + n) * R + i) * 32);
        }
        gen_step<K, SH, TIE_SIMD>(x, y, &cur[0], Tcol, p, gc, dec_lane + size_t(t) * 64, accA, accB);
        gen_step<K, SH, TIE_SIMD>(y, x, &cur[GENERIC_MAX_R], Tcol, p, gc, dec_lane + size_t(t + 1) * 64, accA, accB);
    }
    if (t < p.n_steps) {     // odd tail: one more step, result ends in y
        gen_step<K, SH, TIE_SIMD>(x, y, &nxt[0], Tcol, p, gc, dec_lane + size_t(t) * 64, accA, accB);
#pragma unroll
        for (int q = 0; q < NS; q++) x[q] = y[q];
    }
#pragma unroll
    for (int q = 0; q < NS; q++) {
        mA[q] = uint16_t((x[q] & 0xffffu) >> SH);
        mB[q] = uint16_t(x[q] >> (16 + SH));
    }
    p.acc[fA] = accA;
    p.acc[fB] = accB;
}

}  // namespace vitb
