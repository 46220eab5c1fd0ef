// vitb_registry.h -- table of compiled ACS kernel variants; the host picks one at vitb_create time.
// (One GPU backend, many template instantiations: this is the CUDA analogue of the reference's <K,R,error_t,soft_t> template
// arguments, not a multi-backend dispatch.)
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "acs_pair.cuh"

namespace vitb {

enum KernelFamily { FAMILY_PAIR = 0 };

struct KernelEntry {
    int K, R;
    uint32_t G[16];
    int sh;            // 0: uint16_t metrics, 8: uint8_t metrics held as metric << 8
    int tie;           // VITB_TIE_*
    int consistent;    // max_error == R*(high-low): inverted error is the complementary table entry
    int family;
    const char* name;
    cudaError_t (*launch_pair)(const AcsPairParams&, unsigned n_blocks, cudaStream_t);
};

template <class C, int SH, bool TIE_SIMD, bool CONSISTENT>
cudaError_t launch_pair(const AcsPairParams& p, unsigned n_blocks, cudaStream_t s) {
    AcsPairParams q = p;
    q.n_blocks = n_blocks;
    acs_pair_kernel<C, SH, TIE_SIMD, CONSISTENT><<<(n_blocks + PAIR_WARPS - 1) / PAIR_WARPS, 32 * PAIR_WARPS, 0, s>>>(q);
    return cudaGetLastError();
}

template <class C, int SH, bool TIE_SIMD, bool CONSISTENT>
KernelEntry make_pair_entry(const char* name) {
    KernelEntry e{};
    e.K = C::K; e.R = C::R;
    for (int i = 0; i < C::R; i++) e.G[i] = C::G[i];
    e.sh = SH; e.tie = TIE_SIMD ? 1 : 0; e.consistent = CONSISTENT ? 1 : 0;
    e.family = FAMILY_PAIR; e.name = name;
    e.launch_pair = &launch_pair<C, SH, TIE_SIMD, CONSISTENT>;
    return e;
}

// all eight (metric type x tie-break x consistent-config) variants of one code
#define VITB_PAIR_VARIANTS(VEC, CODE, TAG)                                                         \
    VEC.push_back(make_pair_entry<CODE, 0, false, true>("acs_pair<" TAG ",u16,scalar-tie>"));       \
    VEC.push_back(make_pair_entry<CODE, 8, false, true>("acs_pair<" TAG ",u8,scalar-tie>"));        \
    VEC.push_back(make_pair_entry<CODE, 0, true, true>("acs_pair<" TAG ",u16,simd-tie>"));          \
    VEC.push_back(make_pair_entry<CODE, 8, true, true>("acs_pair<" TAG ",u8,simd-tie>"));           \
    VEC.push_back(make_pair_entry<CODE, 0, false, false>("acs_pair<" TAG ",u16,scalar-tie,cinv>")); \
    VEC.push_back(make_pair_entry<CODE, 8, false, false>("acs_pair<" TAG ",u8,scalar-tie,cinv>"));  \
    VEC.push_back(make_pair_entry<CODE, 0, true, false>("acs_pair<" TAG ",u16,simd-tie,cinv>"));    \
    VEC.push_back(make_pair_entry<CODE, 8, true, false>("acs_pair<" TAG ",u8,simd-tie,cinv>"));

// one translation unit per code family (parallel compilation)
void register_k7r2(std::vector<KernelEntry>& v);
void register_k7r3(std::vector<KernelEntry>& v);
void register_k7r4(std::vector<KernelEntry>& v);
void register_small(std::vector<KernelEntry>& v);

}  // namespace vitb
