// acs_hist_cta.cuh -- survivor-history add-compare-select for K = 15 (Cassini: 16384 states) with uint16_t error metrics:
// ONE FRAME PER CTA of 512 threads x 32 registers.
//
// Replaces ViterbiDecoder_Scalar::update / bfly / renormalise (include/viterbi/viterbi_decoder_scalar.h:28-153) for whole-frame
// batch calls, like acs_cta.cuh, but with the survivor-history formulation of acs_hist.cuh / acs_hist_group.cuh:
//     register = metric << 16 | last <= 16 decisions of the survivor path into that state
//     butterfly = a quarter of an LDS.128 (the total errors of FOUR butterflies) + 3 adds + 2 fused add-min
// instead of acs_cta.cuh's 4 packed adds + 2 packed min + 4 predicated FADDs + decision-byte assembly + a 4 KB decision row per
// step.  One frame per register (no packing: uint16_t metrics leave no spare bits in a 16-bit half), so a CTA holds ONE frame:
//   * 1024 config-5 frames are 1024 CTAs = 6.92 rounds on 148 SMs (acs_cta.cuh: 512 CTAs = 3.46 rounds, the last one half empty),
//     and a 128-frame shard of an 8-GPU run still fills 128 SMs (acs_cta.cuh: 64);
//   * traceback reads one halfword per 16 steps and frame instead of one word per step.
//
// Position algebra of acs_cta.cuh with LOGT = 9: PHI = (q << 9) | t; n steps after an exchange state s sits at PHI = rotr^n(s);
// LB = 5 register-only steps, then the CTA exchanges through 64 KB of shared memory (skewed, conflict-free).  Exchange buffer
// and tables are double-buffered: the last step of a group stores its results into the other buffer as they are produced, the
// tables of the next group are built (from symbols prefetched a group earlier) while the current one runs, and ONE __syncthreads
// per group separates the stores from the loads.  The branch metric table of a whole group ({total, inverted} x 64 patterns x 5 steps, metric field format) is built
// cooperatively from the caller's int16_t row (read where it lies: unpunctured input only, any alignment) at each exchange.
//
// TABLE FETCH: the kernel is bound by shared-memory wavefronts (table fetch + exchange), not by issue slots, so the table holds
// only the total error of a pattern (4 bytes; the inverted error is max_error - total, one more add per butterfly) and is fetched
// by butterfly QUADS: one LDS.128 for the four butterflies that differ in two register bits (acs_cta.cuh, PairMap with NV = 2).
// 16 bytes per butterfly and lane and step became 4: 33 -> 16 wavefronts per warp and step.
//
// TAGS.  The path-1 register of a butterfly must carry the tag 2^k of step k (acs_hist.cuh).  Instead of adding it to 16 registers
// per step, the butterflies of step k-1 whose results become path-1 operands of step k (a register bit, or for the step before an
// exchange a thread bit: either way a constant slot offset) fetch their entries from a second copy of the table that has 2^k added
// (and use max_error + 2 * 2^k, so that the inverted error carries it too) - both paths of the compare move by the same amount, the
// result carries the tag.  The first step of a period finds
// its tag set by the record cut that clears the history fields (and the first step of the frame by the initial metrics).
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
    static constexpr size_t XCH_WORDS = (S::XCH_WORDS + 3) / 4 * 4;                 // the table behind it holds 16-byte slots
    static constexpr size_t TBL_WORDS = size_t(LB) * 2 * NP * 4;                    // [LB][plain, next tag added][NP] x 4 total errors
    static constexpr size_t SMEM_BYTES = (2 * XCH_WORDS + 2 * TBL_WORDS + 64) * 4;  // both double-buffered (156 KB: one CTA per SM)
};

__host__ __device__ constexpr uint32_t rotl_rt(uint32_t v, uint32_t r, uint32_t n) {
    return r == 0 ? v : (((v << r) | (v >> (n - r))) & ((1u << n) - 1u));
}

// one butterfly on registers (x0, x1) = states (j, j + NS/2) with the branch error ex = total_error << 16 of its pattern and
// cq = max_error << 16: the inverted error is cq - ex (scalar.h:107).  The tag of this step is already in the path-1 (TIE_SIMD:
// path-0) register - which is also all the two tie-break flavours differ in -, the tag of the next step comes with ex AND cq (twice,
// so that cq - ex carries it once) where it is due.
__device__ __forceinline__ void hc_bfly(uint32_t& x0, uint32_t& x1, const uint32_t ex, const uint32_t cq) {
    const uint32_t ey = cq - ex;
    const uint32_t b0 = x1 + ey, a1 = x0 + ey;                                  // scalar.h:114,115
    const uint32_t m0 = __viaddmin_u32(x0, ex, b0);                             // new state 2j   stays at q0   scalar.h:113,127
    const uint32_t m1 = __viaddmin_u32(x1, ex, a1);                             // new state 2j+1 goes to q1    scalar.h:116,128
    x0 = m0;
    x1 = m1;
}

// does register q hold a tagged operand in the step whose "path-1 half" is register bit NEXTBIT?
template <int TIE_SIMD>
__host__ __device__ constexpr bool hc_tagged(int q, int nextbit) { return (((q >> nextbit) & 1) != 0) != (TIE_SIMD != 0); }

// the four butterflies whose lower registers are Q, Q | p1, Q | p2, Q | p1 | p2 (p1, p2: the pairing bits): one table fetch
// xout != nullptr (last phase of a full group): the eight results go straight to the other exchange buffer, position (t << LB) | q,
// so that the stores of the exchange run under the butterflies of the phase instead of behind a barrier
template <class C, int PH, int TIE_SIMD, bool STORE, int Q>
__device__ __forceinline__ void hc_bfly_quad_at(uint32_t (&x)[HistCtaShape<C>::NL], const uint4* tbl_ph, const uint32_t mpt, const uint32_t cplain,
                                                const uint32_t cpre, uint32_t* xout) {
    using H = HistCtaShape<C>;
    constexpr PairMap M = pair_map<C, H::LOGT, true, 2>(PH);
    static_assert(M.pb[0] >= 0 && M.pb[1] >= 0, "no two pairing bits with independent patterns");
    static_assert(pair_map_conflict_free<C, H::LOGT, true, 2>(PH), "table fetch with shared-memory bank conflicts inside a quarter warp");
    constexpr int bit = 1 << (H::LB - 1 - PH), p1 = 1 << M.pb[0], p2 = 1 << M.pb[1];
    if constexpr ((Q & (bit | p1 | p2)) == 0) {
        constexpr uint32_t ia = pair_apply(M, C::R, bfly_pattern<C>(rotl_bits(uint32_t(Q) << H::LOGT, PH, H::SB)));   // register-bit part of the old state index
        static_assert(pair_apply(M, C::R, bfly_pattern<C>(rotl_bits(uint32_t(Q | p1) << H::LOGT, PH, H::SB))) == (ia ^ 1u) &&
                      pair_apply(M, C::R, bfly_pattern<C>(rotl_bits(uint32_t(Q | p2) << H::LOGT, PH, H::SB))) == (ia ^ 2u),
                      "the butterflies of a quad must sit in one table slot");
        // entries with the next step's tag added: both results of a butterfly land in the same half of the next step's butterflies.
        // Before an exchange that half is a thread bit: mpt and cplain = cpre carry the choice then.
        constexpr bool pre = (PH < H::LB - 1) && hc_tagged<TIE_SIMD>(Q, H::LB - 2 - PH);
        const uint4 e = tbl_ph[(ia ^ mpt) + (pre ? uint32_t(H::NP) : 0u)];   // total_error of the four patterns   (scalar.h:66-73)
        const uint32_t cq = pre ? cpre : cplain;
        hc_bfly(x[Q], x[Q | bit], e.x, cq);
        hc_bfly(x[Q | p1], x[Q | p1 | bit], e.y, cq);
        hc_bfly(x[Q | p2], x[Q | p2 | bit], e.z, cq);
        hc_bfly(x[Q | p1 | p2], x[Q | p1 | p2 | bit], e.w, cq);
        if constexpr (STORE) {                               // xout = buffer + slot(t << LB): slot((t << LB) | q) = slot(t << LB) + q
            if constexpr (bit == 1 && p1 == 2 && p2 == 4) {
                cta_store4<Q>(xout, x);                      // the quad's eight registers are Q .. Q+7: two STS.128
                cta_store4<Q + 4>(xout, x);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    constexpr int qs[8] = {Q, Q | bit, Q | p1, Q | p1 | bit, Q | p2, Q | p2 | bit, Q | p1 | p2, Q | p1 | p2 | bit};
                    xout[qs[k]] = x[qs[k]];
                }
            }
        }
    }
}

template <class C, int PH, int TIE_SIMD, bool STORE, int... Qs>
__device__ __forceinline__ void hc_bfly_all(uint32_t (&x)[HistCtaShape<C>::NL], const uint4* tbl_ph, const uint32_t mpt, const uint32_t cplain,
                                            const uint32_t cpre, uint32_t* xout, std::integer_sequence<int, Qs...>) {
    (hc_bfly_quad_at<C, PH, TIE_SIMD, STORE, Qs>(x, tbl_ph, mpt, cplain, cpre, xout), ...);
}

// slot part of thread t per phase; the phase before an exchange adds the offset of the tagged copy for the threads whose registers
// become path-1 (TIE_SIMD: path-0) operands of the step behind the exchange: position bit LOGT - 1 -> register bit LB - 1
template <class C, int TIE_SIMD, int... PHs>
__device__ __forceinline__ void hc_thread_slots(uint32_t (&mpt)[sizeof...(PHs)], uint32_t t, std::integer_sequence<int, PHs...>) {
    using H = HistCtaShape<C>;
    pair_thread_slots<C, H::LOGT, true, 2>(mpt, t, std::integer_sequence<int, PHs...>{});
    const bool top = ((t >> (H::LOGT - 1)) & 1u) != 0u;
    if (top != (TIE_SIMD != 0)) mpt[H::LB - 1] += uint32_t(H::NP);
}

// grid = number of frames, block = 512, dynamic shared memory = HistCtaShape::SMEM_BYTES.  Whole frames only (no resume).
// MINB: CTAs per SM the register allocation is made for.  1: 128 registers per thread.  (2 = 64 registers was measured slower when
// the kernel still fitted twice per SM - it spilled -; with the double-buffered exchange, 168 KB of shared memory, it no longer does.)
template <class C, int TIE_SIMD, int MINB>
__global__ void __launch_bounds__(HistCtaShape<C>::T, MINB) acs_hist_cta_kernel(const AcsParams p) {
    using H = HistCtaShape<C>;
    using S = typename H::S;
    constexpr int LB = H::LB, NL = H::NL, NW = H::NW, R = C::R, NP = H::NP, SB = H::SB, LOGT = H::LOGT, HB = H::HB;
    extern __shared__ uint32_t smem[];
    // Two exchange buffers and two table sets, used in turn (cur): buffer cur = the registers at the start of the current group
    // (read back by the exchange, kept for the rollback), the other one takes the results of the group while it still runs.
    uint32_t* xch0 = smem;                                            // [2][NS (skewed)]
    uint4* tbl0 = reinterpret_cast<uint4*>(smem + 2 * H::XCH_WORDS);  // [2][LB][plain, tagged][NP] total errors of slots idx, idx ^ 1, idx ^ 2, idx ^ 3
    uint32_t* red = smem + 2 * H::XCH_WORDS + 2 * H::TBL_WORDS;       // [WARPS] reduction scratch
    uint32_t* flag = red + H::WARPS;                                  // [0..1] trigger flag of the groups in turn, [2] replay scratch
    static_assert(H::WARPS + 3 <= 64, "scratch words");

    const uint32_t t = threadIdx.x;
    const uint32_t rd_off = S::rd_off(t), wr_off = S::wr_off(t);      // exchange addresses: per-thread part (acs_cta.cuh)
    const size_t f = blockIdx.x;
    const HistConsts c = hist_consts<1>(p);

    uint32_t mpt[LB];                                                 // thread part of the table slot per phase
    hc_thread_slots<C, TIE_SIMD>(mpt, t, std::make_integer_sequence<int, LB>{});

    uint32_t x[NL];
    uint64_t acc = 0;
    {
        const uint32_t s0 = p.start_state & uint32_t(C::NS - 1);      // core.h:209-210; phase 0: position = state
#pragma unroll
        for (int q = 0; q < NL; q++)                                  // + the tag of step 0 where step 0 expects it
            x[q] = ((((uint32_t(q) << LOGT) | t) == s0) ? c.init_start : c.init_other) | (hc_tagged<TIE_SIMD>(q, LB - 1) ? 1u : 0u);
    }

    const uint32_t n_periods = (p.n_steps + HB - 1) / HB;
    uint32_t* rec = static_cast<uint32_t*>(p.dec) + f * size_t(n_periods) * (size_t(C::NS) / 2) + t * 4;
    const int16_t* row = reinterpret_cast<const int16_t*>(static_cast<const uint8_t*>(p.sym) + f * p.sym_row_bytes);

    // Branch metric tables of the n steps of the group that starts at first_step: thread t < LB * NP fills slot (t % NP) of step
    // (t / NP) - its own pattern is the one that sits at that slot index, the other three members of the slot are the values of the
    // lanes t ^ 1, t ^ 2, t ^ 3 (shuffles: the whole warp belongs to one step) - plain and with the next step's tag added.
    const uint32_t b_tph = t / NP, b_slot = t % NP;
    const uint32_t b_pat = pair_unapply_phase<C, LOGT, true, 2>(b_tph, b_slot, std::make_integer_sequence<int, LB>{});
    auto build_tables = [&](uint4* tbl, uint32_t first_step, uint32_t n) {
        if (t < uint32_t(LB * NP) && b_tph < n) {                   // uniform per warp: NP is a multiple of 32
            const uint32_t step = first_step + b_tph;
            const int16_t* sy = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(row) + step * uint32_t(R * sizeof(int16_t)));
            uint32_t tot = 0;
#pragma unroll
            for (int i = 0; i < R; i++) {
                const uint32_t sv = uint32_t(uint16_t(__ldg(sy + i))) << 16;
                const uint32_t lo = sv + c.c_low, hi = c.c_high - sv;       // s - low, high - s  (scalar.h:96-105 for s in [low, high])
                tot += ((b_pat >> i) & 1u) ? hi : lo;       // viterbi_branch_table.h:52 + scalar.h:66-73
            }
            const uint32_t t1 = __shfl_xor_sync(0xffffffffu, tot, 1), t2 = __shfl_xor_sync(0xffffffffu, tot, 2), t3 = __shfl_xor_sync(0xffffffffu, tot, 3);
            // the tag of the next step; none behind a record cut (the cut sets it)
            const uint32_t nxt = (2u << (step % uint32_t(HB))) & 0xffffu;
            uint4* dst = tbl + size_t(b_tph) * 2 * NP + b_slot;
            dst[0] = make_uint4(tot, t1, t2, t3);
            dst[NP] = make_uint4(tot + nxt, t1 + nxt, t2 + nxt, t3 + nxt);
        }
    };

    uint32_t pst = 0, r = 0;
    // history record of the period that just ended (position order), then clear the history fields and set the tag of the next
    // period's first step, whose path-1 operands are the registers with bit NEXTBIT set
    auto emit_record = [&](auto nextbit) {
        constexpr int NEXTBIT = decltype(nextbit)::value;         // -1: the last, partial record of the frame
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) {
            w[i] = __byte_perm(x[2 * i], x[2 * i + 1], 0x5410u);
            if constexpr (TIE_SIMD) w[i] = ~w[i];                                          // the tag marked path 0: decision = !tag
            if constexpr (TIE_SIMD || NEXTBIT < 0) w[i] &= 0x00010001u * ((1u << pst) - 1u);   // bits of steps not taken (and the last step's tag for a step that never comes)
        }
        uint32_t* dst = rec + size_t(r) * (size_t(C::NS) / 2);
#pragma unroll
        for (int v = 0; v < NW / 4; v++) *reinterpret_cast<uint4*>(dst + size_t(v) * (H::T * 4)) = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
#pragma unroll
        for (int q = 0; q < NL; q++) x[q] = (x[q] & 0xffff0000u) | ((NEXTBIT >= 0 && hc_tagged<TIE_SIMD>(q, NEXTBIT)) ? 1u : 0u);
        r++;
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

    // max_error << 16 = total + inverted error of any pattern (scalar.h:107); with twice the next step's tag for the tagged entries
    const uint32_t cplain = c.c_inv + uint32_t(R) * (c.c_low + c.c_high);
    const bool tag_behind_exchange = (((t >> (LOGT - 1)) & 1u) != 0u) != (TIE_SIMD != 0);
    auto run_bfly = [&](auto PHc, auto store_tag, const uint4* tbl, uint32_t* xout) {
        constexpr int PH = decltype(PHc)::value;
        constexpr bool STORE = decltype(store_tag)::value;
        const uint32_t cpre = cplain + ((4u << pst) & 0x1fffeu);
        if constexpr (PH == LB - 1) {
            const uint32_t cl = tag_behind_exchange ? cpre : cplain;
            hc_bfly_all<C, PH, TIE_SIMD, STORE>(x, tbl + PH * 2 * NP, mpt[PH], cl, cl, xout, std::make_integer_sequence<int, NL>{});
        } else {
            hc_bfly_all<C, PH, TIE_SIMD, false>(x, tbl + PH * 2 * NP, mpt[PH], cplain, cpre, xout, std::make_integer_sequence<int, NL>{});
        }
    };
    // the symbols of a group are pulled into L1 one group ahead of the table build that reads them
    auto prefetch_symbols = [&](uint32_t first_step) {
        if (t < uint32_t(LB)) {
            const uint32_t st = first_step + t;
            if (st < p.n_steps) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + size_t(st) * R));
        }
    };

    uint32_t done = 0, cur = 0;
    // prologue: buffer 0 holds the registers at the start of group 0, table set 0 its tables
#pragma unroll
    for (int q = 0; q < NL; q++) xch0[rd_off + uint32_t(q) * S::RD_STRIDE] = x[q];
    if (p.n_steps) build_tables(tbl0, 0u, p.n_steps < uint32_t(LB) ? p.n_steps : uint32_t(LB));
    prefetch_symbols(uint32_t(LB));
    __syncthreads();

    while (done < p.n_steps) {
        const uint32_t left = p.n_steps - done, span = left < uint32_t(LB) ? left : uint32_t(LB);
        const uint32_t pst0 = pst, r0 = r;
        const bool full = span == uint32_t(LB);
        uint32_t* xcur = xch0 + size_t(cur) * H::XCH_WORDS;
        uint32_t* xnew = xch0 + size_t(cur ^ 1u) * H::XCH_WORDS;
        const uint4* tcur = tbl0 + size_t(cur) * (H::TBL_WORDS / 4);
        uint4* tnew = tbl0 + size_t(cur ^ 1u) * (H::TBL_WORDS / 4);

        // ---- tables of the NEXT group into the other set (every thread left that set at the barrier of the previous group, or at
        //      the one behind its replay), symbols of the group after that on their way
        if (full && done + uint32_t(LB) < p.n_steps) {
            const uint32_t nleft = p.n_steps - done - uint32_t(LB);
            build_tables(tnew, done + uint32_t(LB), nleft < uint32_t(LB) ? nleft : uint32_t(LB));
            prefetch_symbols(done + 2u * uint32_t(LB));
        }

        // ---- speculative run of the group (no renormalisation); thread 0 keeps the running maximum of state 0's register.
        //      The last phase of a full group stores its results into the other exchange buffer as they are produced.
        uint32_t mx = 0u;
        auto spec_phase = [&](auto PHc, auto guard_tag) {
            constexpr int PH = decltype(PHc)::value;
            constexpr bool GUARD = decltype(guard_tag)::value;
            if constexpr (GUARD) { if (uint32_t(PH) >= span) return; }
            run_bfly(PHc, std::integral_constant<bool, !GUARD>{}, tcur, xnew + wr_off);
            mx = max(mx, x[0]);
            pst++;
            if constexpr (PH != LB - 1) { if (pst == uint32_t(HB)) emit_record(std::integral_constant<int, LB - 2 - PH>{}); }     // after the last phase: behind the exchange
        };
        if (full) {
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
        if (t == 0) flag[cur] = (mx >= c.thr) ? 1u : 0u;     // metric >= threshold  <=>  register >= threshold << 16
        __syncthreads();                                     // the ONE barrier of a group: steps done, results and next tables in place
        const uint32_t any_trig = flag[cur];                 // (the flag word of the next group is the other one)

        if (any_trig) {
            // ---- roll back (registers and record bookkeeping) and replay step by step with the reference's renormalisation
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xcur[rd_off + uint32_t(q) * S::RD_STRIDE];
            pst = pst0; r = r0;
            auto replay_phase = [&](auto PHc) {
                constexpr int PH = decltype(PHc)::value;
                if (uint32_t(PH) < span) {
                    run_bfly(PHc, std::false_type{}, tcur, xnew);
                    pst++;
                    if (t == 0) flag[2] = x[0];
                    __syncthreads();
                    const uint32_t x00 = flag[2];
                    if (x00 >= c.thr) {                                    // uniform across the CTA (scalar.h:48)
                        const uint32_t sub = cta_min() & 0xffff0000u;      // scalar.h:140-146
#pragma unroll
                        for (int q = 0; q < NL; q++) x[q] -= sub;          // scalar.h:148-150
                        acc += uint64_t(sub >> 16);                        // scalar.h:49, 152 (only thread 0's copy is stored)
                    } else {
                        __syncthreads();                                   // flag[2] may be rewritten by the next phase
                    }
                    if constexpr (PH != LB - 1) { if (pst == uint32_t(HB)) emit_record(std::integral_constant<int, LB - 2 - PH>{}); }
                }
            };
            replay_phase(std::integral_constant<int, 0>{});
            replay_phase(std::integral_constant<int, 1>{});
            replay_phase(std::integral_constant<int, 2>{});
            replay_phase(std::integral_constant<int, 3>{});
            replay_phase(std::integral_constant<int, 4>{});
            if (full) {
                // the exchange of the replayed group: value at (q, t) moves to PHI' = (t << LB) | q
#pragma unroll
                for (int q = 0; q < NL; q += 4) *reinterpret_cast<uint4*>(xnew + wr_off + q) = make_uint4(x[q], x[q + 1], x[q + 2], x[q + 3]);
            }
            __syncthreads();
        }

        done += span;
        if (full) {
            // ---- exchange, read side: positions back to PHI = s.  The buffer just read is the rollback copy of the next group; the
            //      other one is written again only in the last phase of the next group, behind every thread's loads.
#pragma unroll
            for (int q = 0; q < NL; q++) x[q] = xnew[rd_off + uint32_t(q) * S::RD_STRIDE];
            cur ^= 1u;
            if (pst == uint32_t(HB)) {
                emit_record(std::integral_constant<int, LB - 1>{});   // records are cut in the layout the traceback expects: PHI = rotr^(steps mod 5)(s)
                // the rollback copy must match the registers the next group starts from (history fields cleared); every thread
                // rewrites exactly the words it has just read and will read back on a rollback: no barrier needed
#pragma unroll
                for (int q = 0; q < NL; q++) xnew[rd_off + uint32_t(q) * S::RD_STRIDE] = x[q];
            }
        }
    }
    if (pst) emit_record(std::integral_constant<int, -1>{}); // last, partial record (its pst is still needed for the SIMD tie-break mask)

    // final metrics in logical order: state s sits at PHI = rotr^ph(s)   (core.h:195-199 reads old_metrics[end_state])
    const uint32_t ph = p.n_steps % uint32_t(LB);
    uint16_t* m = p.metrics + f * C::NS;
#pragma unroll
    for (int q = 0; q < NL; q++) m[rotl_rt((uint32_t(q) << LOGT) | t, ph, SB)] = uint16_t(x[q] >> 16);
    if (t == 0) p.acc[f] = acc;
}

}  // namespace vitb
