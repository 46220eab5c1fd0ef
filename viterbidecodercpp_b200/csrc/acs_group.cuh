// acs_group.cuh -- add-compare-select with a frame pair spread over T = 2^LOGT lanes of a warp (K <= 9 today).
//
// Same arithmetic as acs_pair.cuh (two frames packed in the 16-bit halves of every register, VIADD.16x2 + VIMNMX.U16x2 with
// predicate outputs, predicated FADDs collecting decision bits), but the 2^(K-1) states of the pair are divided among T lanes:
//     position PHI = (q << LOGT) | t      q = register index inside the lane (LB = K-1-LOGT bits), t = lane inside the group
//     after n steps since the last exchange, logical state s sits at PHI = rotr^n(s)   ((K-1)-bit rotation)
// For n = 0 .. LB-1 the butterfly partner differs in a REGISTER bit, so LB trellis steps run with no communication at all;
// then one exchange through shared memory (a fixed bit rotation of PHI, conflict-free with an XOR swizzle, __syncwarp only)
// brings the layout back to PHI = s.  This gives (a) 2^LOGT times more warps for the same batch - the B200 needs >= 4 warps
// per SM sub-partition to overlap the 2-cycle issue of VIADD.16x2 - and (b) a hot loop of a few KB that stays in the
// instruction cache, where the fully unrolled one-thread-per-pair kernel is instruction-fetch bound.
//
// The branch pattern of a butterfly splits into a compile-time part (register bits) XOR a per-lane part (lane bits, one value
// per phase): the per-lane part is folded into the branch metric table by swapping e_low/e_high of the affected symbols.
//
// Decision rows are stored in this kernel family's private layout: lane t writes, per step, bit q = decision of the state
// that sits in its register q after the step, i.e. logical state rotl^(n+1)((q << LOGT) | t) with n = row % LB.
// traceback_group_kernel (traceback.cuh) reads that layout; vitb_get_decisions converts it to the reference layout on request.
#pragma once
#include <cstdint>
#include <utility>
#include <cuda_runtime.h>
#include "vitb_code.cuh"
#include "acs_pair.cuh"

namespace vitb {

template <class C, int LOGT>
struct GroupShape {
    static_assert(LOGT >= 1 && LOGT <= 5, "frame pair must span 2..32 lanes");
    static constexpr int T = 1 << LOGT;
    static constexpr int SB = C::SB;
    static constexpr int LB = SB - LOGT;            // local index bits = steps between exchanges
    static_assert(LB >= 1, "too many lanes for this constraint length");
    static constexpr int NL = 1 << LB;              // packed registers per lane
    static_assert(NL <= 32, "at most 32 registers of path metrics per lane");
    static constexpr int PPW = 32 / T;              // frame pairs per warp
    static constexpr int NACC = NL > 16 ? 2 : 1;    // float accumulators per frame
    static constexpr int W = NL > 16 ? 2 : 1;       // 32-bit decision words per lane per step
    static constexpr int WARPS = 4;                 // warps per CTA
    static constexpr int MIN_CTAS = NL <= 16 ? 4 : 3;   // register budget: 128 regs when a lane holds <= 16 metrics, else 168 (measured best, profiles/r01_summary.md)
    // exchange swizzle (verified conflict-free for reads and writes by tests/test_host_cpu.py::test_exchange_swizzle)
    static __host__ __device__ constexpr uint32_t swz(uint32_t qp) {
        return LB >= LOGT ? ((qp >> (LB - LOGT)) & uint32_t(T - 1)) : (qp & 31u);
    }
    // shared-memory word of position PHI' for pair p of the warp
    static __host__ __device__ constexpr uint32_t slot(uint32_t p, uint32_t phi) {
        const uint32_t qp = phi >> LOGT, tp = phi & uint32_t(T - 1);
        return qp * 32u + ((p * uint32_t(T) + tp) ^ swz(qp));
    }
};

template <class C, int LOGT, int PH, int TIE_SIMD, int Q>
__device__ __forceinline__ void group_bfly_at(uint32_t (&x)[GroupShape<C, LOGT>::NL], const uint32_t (&T)[C::NP],
                                              float (&fa)[2][GroupShape<C, LOGT>::NACC], const uint32_t c_inv2, const bool consistent,
                                              const uint32_t max_err2) {
    using S = GroupShape<C, LOGT>;
    constexpr bool SAT = Sat<TIE_SIMD>::value;          // saturating flavour, see acs_pair.cuh
    constexpr int bit = 1 << (S::LB - 1 - PH);
    if constexpr ((Q & bit) == 0) {
        constexpr int q0 = Q, q1 = Q | bit;
        constexpr uint32_t jq = rotl_bits(uint32_t(q0) << LOGT, PH, S::SB);   // register-bit part of the old state index
        constexpr uint32_t pat = bfly_pattern<C>(jq);                          // ... and of its branch pattern
        constexpr uint32_t ipat = (~pat) & uint32_t(C::NP - 1);
        const uint32_t tot = T[pat];
        const uint32_t inv = consistent ? T[ipat] : (SAT ? __vsubus2(max_err2, tot) : __vadd2(T[ipat], c_inv2));
        const uint32_t a0 = metric_add<SAT>(x[q0], tot), b0 = metric_add<SAT>(x[q1], inv);     // scalar.h:113-114
        const uint32_t a1 = metric_add<SAT>(x[q0], inv), b1 = metric_add<SAT>(x[q1], tot);     // scalar.h:115-116
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
        constexpr int acc0 = (q0 >> 4) % S::NACC, acc1 = (q1 >> 4) % S::NACC;
        constexpr float w0 = float(1u << (q0 & 15)), w1 = float(1u << (q1 & 15));
        if (dA0) fa[0][acc0] += w0;
        if (dB0) fa[1][acc0] += w0;
        if (dA1) fa[0][acc1] += w1;
        if (dB1) fa[1][acc1] += w1;
    }
}

template <class C, int LOGT, int PH, int TIE_SIMD, int... Qs>
__device__ __forceinline__ void group_bfly_all(uint32_t (&x)[GroupShape<C, LOGT>::NL], const uint32_t (&T)[C::NP],
                                               float (&fa)[2][GroupShape<C, LOGT>::NACC], const uint32_t c_inv2, const bool consistent,
                                               const uint32_t max_err2, std::integer_sequence<int, Qs...>) {
    (group_bfly_at<C, LOGT, PH, TIE_SIMD, Qs>(x, T, fa, c_inv2, consistent, max_err2), ...);
}

// per-lane, per-phase constants that fold the lane part of the branch pattern into the table
template <class C, int LOGT>
struct LaneConsts {
    using S = GroupShape<C, LOGT>;
    uint32_t m[S::LB][C::R];      // 0 or ~0: x = sym ^ m;  e(folded bit 0) = x + (m ? c_high : c_low),  e(folded bit 1) = ~x + (m ? c_low : c_high)
};

template <class C, int LOGT, int SH, int TIE_SIMD, bool CONSISTENT>
struct GroupKernel {
    using S = GroupShape<C, LOGT>;
    static constexpr int LB = S::LB, NL = S::NL, R = C::R, NP = C::NP, T = S::T, SB = S::SB;

    template <int PH>
    static __device__ __forceinline__ void step(uint32_t (&x)[NL], const uint32_t* sym, const LaneConsts<C, LOGT>& lc, const AcsParams& p,
                                                uint32_t* dec_lane_row, uint64_t& accA, uint64_t& accB, const uint32_t lane, const uint32_t pair_mask) {
        uint32_t lo[R], hi[R];
#pragma unroll
        for (int i = 0; i < R; i++) {
            const uint32_t mk = lc.m[PH][i], cd = p.c_low2 ^ p.c_high2;
            const uint32_t xs = sym[i] ^ mk;
            lo[i] = __vadd2(xs, p.c_low2 ^ (mk & cd));      // constants swap when the lane part of the pattern flips this symbol
            hi[i] = __vadd2(~xs, p.c_high2 ^ (mk & cd));
        }
        constexpr bool SAT = Sat<TIE_SIMD>::value;
        uint32_t Tt[NP];
        if constexpr (SAT) table_build_sat<R, SH>(Tt, lo, hi);
        else TableBuild<R, R>::run(Tt, lo, hi);
        float fa[2][S::NACC];
#pragma unroll
        for (int a = 0; a < S::NACC; a++) { fa[0][a] = 8388608.f; fa[1][a] = 8388608.f; }
        group_bfly_all<C, LOGT, PH, TIE_SIMD>(x, Tt, fa, p.c_inv2, CONSISTENT, p.max_err2, std::make_integer_sequence<int, NL>{});

        if constexpr (S::W == 1) {
            dec_lane_row[0] = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[1][0]), 0x5410);   // A bits | B bits << 16
        } else {
            const uint32_t wA = __byte_perm(__float_as_uint(fa[0][0]), __float_as_uint(fa[0][1]), 0x5410);
            const uint32_t wB = __byte_perm(__float_as_uint(fa[1][0]), __float_as_uint(fa[1][1]), 0x5410);
            *reinterpret_cast<uint2*>(dec_lane_row) = make_uint2(wA, wB);
        }

        // renormalisation trigger: state 0 sits in register 0 of lane 0 of the group in every phase (scalar.h:48)
        const uint32_t x00 = __shfl_sync(0xffffffffu, x[0], int(lane & ~uint32_t(T - 1)));
        bool trigB, trigA;
        (void)__vibmin_u16x2(p.thr2, x00, &trigB, &trigA);
        if (trigA || trigB) {
            uint32_t m = packed_min<NL>(x);
#pragma unroll
            for (int d = 1; d < T; d <<= 1) m = __vminu2(m, __shfl_xor_sync(pair_mask, m, d));
            m = metric_field<SAT, SH>(m);
            const uint32_t mA = m & 0xffffu, mB = m >> 16;
            const uint32_t sub = (trigA ? mA : 0u) | ((trigB ? mB : 0u) << 16);
            const uint32_t neg = __vsub2(0u, sub);
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = __vadd2(x[q], neg);
            if (trigA) accA += uint64_t(mA >> SH);
            if (trigB) accB += uint64_t(mB >> SH);
        }
    }

    template <int PH>
    static __device__ __forceinline__ void phase(uint32_t (&x)[NL], const uint32_t (&cur)[LB * R], const LaneConsts<C, LOGT>& lc,
                                                 const AcsParams& p, uint32_t* dec_lane, const uint32_t done, const int ph0,
                                                 uint64_t& accA, uint64_t& accB, const uint32_t lane, const uint32_t pair_mask) {
        if (PH < ph0) return;
        const uint32_t tt = done + uint32_t(PH - ph0);
        if (tt >= p.n_steps) return;
        step<PH>(x, &cur[PH * R], lc, p, dec_lane + size_t(tt) * 32 * S::W, accA, accB, lane, pair_mask);
    }

    // a whole exchange period, no per-step bounds checks: the common path of batch decoding
    template <int... PHs>
    static __device__ __forceinline__ void fast_group(uint32_t (&x)[NL], const uint32_t (&cur)[LB * R], const LaneConsts<C, LOGT>& lc,
                                                      const AcsParams& p, uint32_t* dec_rows, uint64_t& accA, uint64_t& accB,
                                                      const uint32_t lane, const uint32_t pair_mask, std::integer_sequence<int, PHs...>) {
        (step<PHs>(x, &cur[PHs * R], lc, p, dec_rows + PHs * 32 * S::W, accA, accB, lane, pair_mask), ...);
    }

    template <int... PHs>
    static __device__ __forceinline__ void group(uint32_t (&x)[NL], const uint32_t (&cur)[LB * R], const LaneConsts<C, LOGT>& lc,
                                                 const AcsParams& p, uint32_t* dec_lane, const uint32_t done, const int ph0,
                                                 uint64_t& accA, uint64_t& accB, const uint32_t lane, const uint32_t pair_mask,
                                                 std::integer_sequence<int, PHs...>) {
        (phase<PHs>(x, cur, lc, p, dec_lane, done, ph0, accA, accB, lane, pair_mask), ...);
    }
};

// grid = ceil(n_wblocks / WARPS), block = 32 * WARPS
template <class C, int LOGT, int SH, int TIE_SIMD, bool CONSISTENT>
__global__ void __launch_bounds__(32 * GroupShape<C, LOGT>::WARPS, GroupShape<C, LOGT>::MIN_CTAS) acs_group_kernel(const AcsParams p) {
    using S = GroupShape<C, LOGT>;
    using Kn = GroupKernel<C, LOGT, SH, TIE_SIMD, CONSISTENT>;
    constexpr int LB = S::LB, NL = S::NL, R = C::R, T = S::T, SB = S::SB, PPW = S::PPW;
    __shared__ uint32_t xch[S::WARPS][NL * 32];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wblk = blockIdx.x * S::WARPS + warp;
    if (wblk >= p.n_blocks) return;
    const uint32_t pw = lane / T, t = lane % T;                      // pair inside the warp, lane inside the pair
    const uint32_t pair_mask = (T == 32) ? 0xffffffffu : (((1u << T) - 1u) << (lane & ~uint32_t(T - 1)));
    const size_t fA = (size_t(wblk) * PPW + pw) * 2, fB = fA + 1;
    uint32_t* my_xch = xch[warp];

    // lane part of the branch pattern for every phase -> table folding constants
    LaneConsts<C, LOGT> lc;
#pragma unroll
    for (int n = 0; n < LB; n++) {
        const uint32_t pt = bfly_pattern_dyn<C>(rotl_bits(t, n, SB));
#pragma unroll
        for (int i = 0; i < R; i++) {
            const bool b = (pt >> i) & 1u;
            lc.m[n][i] = b ? 0xffffffffu : 0u;
        }
    }

    int ph = int(p.dec_row0 % uint32_t(LB));
    uint32_t x[NL];
    uint64_t accA = 0, accB = 0;
    uint16_t* mA = p.metrics + fA * C::NS;
    uint16_t* mB = p.metrics + fB * C::NS;
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

    const uint32_t* pk = p.pk + size_t(wblk) * p.n_steps * R * PPW + pw;
    uint32_t* dec_lane = static_cast<uint32_t*>(p.dec) + ((size_t(wblk) * p.dec_rows + p.dec_row0) * 32 + lane) * S::W;

    auto exchange = [&]() {
        // all LB phases done: rotate positions back to PHI = s.  Value at (q, t) moves to PHI' = (t << LB) | q.
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NL; q++) my_xch[S::slot(pw, (t << LB) | uint32_t(q))] = x[q];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = my_xch[S::slot(pw, (uint32_t(q) << LOGT) | t)];
    };

    uint32_t done = 0;
    // ---- head: a call that starts in the middle of an exchange period (streaming API) or is shorter than one period
    if (ph != 0 || p.n_steps < uint32_t(LB)) {
        uint32_t cur[LB * R];
#pragma unroll
        for (int k = 0; k < LB * R; k++) {
            const int PH = k / R;
            const uint32_t tt = uint32_t(PH - ph);
            cur[k] = (PH >= ph && tt < p.n_steps) ? __ldg(pk + (size_t(tt) * R + (k % R)) * PPW) : 0u;
        }
        Kn::group(x, cur, lc, p, dec_lane, 0u, ph, accA, accB, lane, pair_mask, std::make_integer_sequence<int, LB>{});
        const uint32_t span = uint32_t(LB - ph);
        if (p.n_steps >= span) { done = span; exchange(); ph = 0; }
        else { done = p.n_steps; ph += int(p.n_steps); }
    }
    // ---- body: whole periods, symbols of the next period prefetched while this one is computed
    if (done + LB <= p.n_steps) {
        uint32_t nxt[LB * R];
#pragma unroll
        for (int k = 0; k < LB * R; k++) nxt[k] = __ldg(pk + (size_t(done) * R + k) * PPW);
        uint32_t* drow = dec_lane + size_t(done) * 32 * S::W;
#pragma unroll 1
        while (done + LB <= p.n_steps) {
            uint32_t cur[LB * R];
#pragma unroll
            for (int k = 0; k < LB * R; k++) cur[k] = nxt[k];
            const uint32_t nb = done + LB;
#pragma unroll
            for (int k = 0; k < LB * R; k++) nxt[k] = (nb + uint32_t(k / R) < p.n_steps) ? __ldg(pk + (size_t(nb) * R + k) * PPW) : 0u;
            Kn::fast_group(x, cur, lc, p, drow, accA, accB, lane, pair_mask, std::make_integer_sequence<int, LB>{});
            exchange();
            drow += LB * 32 * S::W;
            done = nb;
        }
    }
    // ---- tail: fewer than LB steps left, phase 0
    if (done < p.n_steps) {
        uint32_t cur[LB * R];
#pragma unroll
        for (int k = 0; k < LB * R; k++) {
            const uint32_t tt = done + uint32_t(k / R);
            cur[k] = (tt < p.n_steps) ? __ldg(pk + (size_t(tt) * R + (k % R)) * PPW) : 0u;
        }
        Kn::group(x, cur, lc, p, dec_lane, done, 0, accA, accB, lane, pair_mask, std::make_integer_sequence<int, LB>{});
        ph = int(p.n_steps - done);
        done = p.n_steps;
    }

    // write back in logical order; state s sits at PHI = rotr^ph(s)
#pragma unroll
    for (int q = 0; q < NL; q++) {
        const uint32_t s = rotl_bits((uint32_t(q) << LOGT) | t, ph, SB);
        mA[s] = uint16_t((x[q] & 0xffffu) >> SH);
        mB[s] = uint16_t(x[q] >> (16 + SH));
    }
    if (t == 0) { p.acc[fA] = accA; p.acc[fB] = accB; }
}

}  // namespace vitb
