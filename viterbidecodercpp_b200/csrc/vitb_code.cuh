// vitb_code.cuh -- compile-time description of a convolutional code and the trellis index algebra the kernels share.
//
// The reference keeps K and R as template parameters and the generator polynomials G as a constructor argument used only to
// fill ViterbiBranchTable (include/viterbi/viterbi_branch_table.h:33-55).  On the GPU the table never exists as data: for
// catalogued codes G is a template parameter too, so "which branch metric does butterfly j use" folds into register names.
#pragma once
#include <cstdint>
#include <utility>

namespace vitb {

template <int K_, int R_, uint32_t... Gs>
struct Code {
    static_assert(sizeof...(Gs) == R_, "need R generator polynomials");
    static constexpr int K = K_;
    static constexpr int R = R_;
    static constexpr int SB = K_ - 1;                 // state bits        (core.h:162)
    static constexpr int NS = 1 << SB;                // states            (core.h:163)
    static constexpr int NP = 1 << R_;                // distinct R-bit branch patterns
    static constexpr uint32_t G[R_] = {Gs...};
};

__host__ __device__ constexpr uint32_t parity32(uint32_t x) {
    x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
    return x & 1u;
}

// Branch pattern of butterfly j (old state with leading bit 0): bit i = parity((j<<1) & G[i])   (viterbi_branch_table.h:45-54).
// bit i set  <=> the table entry BT[i][j] is soft_decision_high.
template <class C>
__host__ __device__ constexpr uint32_t bfly_pattern(uint32_t j) {
    uint32_t p = 0;
    for (int i = 0; i < C::R; i++) p |= parity32((j << 1) & C::G[i]) << i;
    return p;
}

// same, for a run-time index (lane part of the pattern in the group kernels); G stays a compile-time constant
template <class C, int... Is>
__device__ __forceinline__ uint32_t bfly_pattern_dyn_impl(uint32_t j, std::integer_sequence<int, Is...>) {
    uint32_t p = 0;
    ((p |= (uint32_t(__popc((j << 1) & std::integral_constant<uint32_t, C::G[Is]>::value)) & 1u) << Is), ...);
    return p;
}
template <class C>
__device__ __forceinline__ uint32_t bfly_pattern_dyn(uint32_t j) {
    return bfly_pattern_dyn_impl<C>(j, std::make_integer_sequence<int, C::R>{});
}

// runtime-G flavour of the same thing (generic kernels, host-side table checks)
__host__ __device__ inline uint32_t bfly_pattern_rt(const uint32_t* G, int R, uint32_t j) {
    uint32_t p = 0;
    for (int i = 0; i < R; i++) p |= parity32((j << 1) & G[i]) << i;
    return p;
}

// Direct symbol fetch: largest 32-bit word index of a row that still lies inside the caller's array (loads are clamped to it).
// Word indices inside a row fit 32 bits; the remaining size of a >= 16 GiB array does not, so it is saturated, not truncated.
__host__ __device__ inline uint32_t clamp_words_left(size_t total_bytes, size_t row_offset_bytes) {
    const size_t w = (total_bytes - row_offset_bytes - 4) >> 2;
    return w > size_t(0xffffffffu) ? 0xffffffffu : uint32_t(w);
}

// rotate an n-bit index left/right by r
__host__ __device__ constexpr uint32_t rotl_bits(uint32_t v, int r, int n) {
    r %= n;
    return r == 0 ? v : (((v << r) | (v >> (n - r))) & ((1u << n) - 1u));
}
__host__ __device__ constexpr uint32_t rotr_bits(uint32_t v, int r, int n) { return rotl_bits(v, (n - (r % n)) % n, n); }

}  // namespace vitb
