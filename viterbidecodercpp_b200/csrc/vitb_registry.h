// vitb_registry.h -- table of compiled ACS kernel variants; the host picks one per call (by batch size) from the variants of the
// handle's code.  (One GPU backend, many template instantiations: this is the CUDA analogue of the reference's
// <K,R,error_t,soft_t> template arguments, not a multi-backend dispatch.)
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "acs_pair.cuh"
#include "acs_hist.cuh"
#include "acs_group.cuh"
#include "acs_cta.cuh"
#include "acs_generic.cuh"
#include "acs_hist_group.cuh"
#include "acs_hist_cta.cuh"

namespace vitb {

enum { LAYOUT_PAIR = 0, LAYOUT_GROUP = 1, LAYOUT_CTA = 2, LAYOUT_HISTGROUP = 3, LAYOUT_HISTCTA = 4 };

struct KernelEntry {
    int K, R;
    uint32_t G[16];
    int sh;            // 0: uint16_t metrics, 8: uint8_t metrics held as metric << 8
    int tie;           // VITB_TIE_*
    int consistent;    // max_error == R*(high-low): inverted error is the complementary table entry
    int logt;          // threads per frame pair = 1 << logt
    int layout;        // LAYOUT_PAIR (acs_pair.cuh), LAYOUT_GROUP (acs_group.cuh), LAYOUT_CTA (acs_cta.cuh): decision row format
    int ppw;           // frame pairs per warp block of the packed symbol stream (32 / lanes, 1 for CTA kernels)
    int dec_words;     // 32-bit decision words per thread per step (group/CTA kernels); pair kernels: one uint64 per frame
    const char* name;
    cudaError_t (*launch)(const AcsParams&, cudaStream_t);
    cudaError_t (*launch_direct)(const AcsParams&, cudaStream_t);   // pair kernels only: symbols read from the caller's rows (no ingest)
    // one-thread-per-pair entries: whole-frame batch decode with survivor-history records instead of decision rows (acs_hist.cuh).
    // uint8_t metrics: two frames per lane, 8-step records; uint16_t metrics: one frame per lane, 16-step records, packed stream in
    // the 16-pairs-per-warp-block layout
    cudaError_t (*launch_hist)(const AcsParams&, cudaStream_t);
    // generic entries (acs_generic.cuh): any generator polynomials and any rate R <= GENERIC_MAX_R for this K; R, G and consistent are
    // wildcards in the table and the branch patterns travel as a kernel argument
    int generic;
    cudaError_t (*launch_generic)(const AcsParams&, const GenericCode&, cudaStream_t);
    cudaError_t (*launch_hist_direct)(const AcsParams&, cudaStream_t);
    int variant_id;    // number vitb_set_variant / vitb_get_variants know this entry by; 0 = 1 << logt (lanes per frame pair)
    int tagged_rows;   // LAYOUT_PAIR: decision rows in the tagged-butterfly bit order (uint8_t metrics, not the saturating flavour)
};

inline int entry_variant(const KernelEntry* e) { return e->variant_id ? e->variant_id : (1 << e->logt); }

// Decision-row kernel of the one-thread-per-pair mapping (streaming calls; batch calls when the history kernel is switched off).
// A two-register-set (ping-pong) variant of it was measured and dropped: ptxas scheduled all compare-selects ahead of their
// predicated FADDs, ran out of predicate registers and spilled them (2.14 ms vs 1.42 ms on config 2, profiles/r01_summary.md).
template <class C, int SH, int TIE_SIMD, bool CONSISTENT>
cudaError_t launch_pair(const AcsParams& p, cudaStream_t s) {
    const unsigned grid = (p.n_blocks + PAIR_WARPS - 1) / PAIR_WARPS;
    acs_pair_kernel<C, SH, TIE_SIMD, CONSISTENT><<<grid, 32 * PAIR_WARPS, 0, s>>>(p);
    return cudaGetLastError();
}

template <class C, int SH, int TIE_SIMD, bool CONSISTENT>
cudaError_t launch_pair_direct(const AcsParams& p, cudaStream_t s) {
    const unsigned grid = (p.n_blocks + PAIR_WARPS - 1) / PAIR_WARPS;
    acs_pair_kernel<C, SH, TIE_SIMD, CONSISTENT, PairPeriod<C>::value, true><<<grid, 32 * PAIR_WARPS, 0, s>>>(p);
    return cudaGetLastError();
}

template <class C, int FMT, int TIE_SIMD, bool CONSISTENT, bool DIRECT>
cudaError_t launch_hist(const AcsParams& p, cudaStream_t s) {
    static const unsigned warps = getenv("VITB_HIST_WARPS") ? unsigned(atoi(getenv("VITB_HIST_WARPS"))) : unsigned(HIST_WARPS);
    const unsigned w = (warps >= 1 && warps <= unsigned(HIST_WARPS)) ? warps : unsigned(HIST_WARPS);
    const unsigned grid = (p.n_blocks + w - 1) / w;
    acs_hist_kernel<C, FMT, TIE_SIMD, CONSISTENT, DIRECT><<<grid, 32 * w, 0, s>>>(p);
    return cudaGetLastError();
}

template <class C, int LOGT, int SH, int TIE_SIMD, bool CONSISTENT>
cudaError_t launch_group(const AcsParams& p, cudaStream_t s) {
    constexpr int WARPS = GroupShape<C, LOGT>::WARPS;
    acs_group_kernel<C, LOGT, SH, TIE_SIMD, CONSISTENT><<<(p.n_blocks + WARPS - 1) / WARPS, 32 * WARPS, 0, s>>>(p);
    return cudaGetLastError();
}

// CONSISTENT is irrelevant for the CTA kernel (c_inv is folded into the shared-memory table)
template <class C, int LT, int SH, int TIE_SIMD>
cudaError_t launch_cta(const AcsParams& p, cudaStream_t s) {
    using S = CtaShape<C, LT>;
    // the attribute is per device: set it on every launch (cheap) so that handles on different GPUs all get it
    const cudaError_t e = cudaFuncSetAttribute(acs_cta_kernel<C, LT, SH, TIE_SIMD>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    acs_cta_kernel<C, LT, SH, TIE_SIMD><<<p.n_blocks, S::T, S::SMEM_BYTES, s>>>(p);
    return cudaGetLastError();
}

template <class C, int LT, int SH, int TIE_SIMD>
KernelEntry make_cta_entry(const char* name, int consistent) {
    KernelEntry e{};
    e.K = C::K; e.R = C::R;
    for (int i = 0; i < C::R; i++) e.G[i] = C::G[i];
    e.sh = SH; e.tie = TIE_SIMD; e.consistent = consistent; e.logt = LT; e.name = name;
    e.layout = LAYOUT_CTA; e.ppw = 1; e.dec_words = CtaKernel<C, LT, SH, TIE_SIMD>::W;
    e.launch = &launch_cta<C, LT, SH, TIE_SIMD>;
    return e;
}

#define VITB_CTA_VARIANTS(VEC, CODE, LT, TAG)                                                   \
    for (int cons = 0; cons < 2; cons++) {                                                      \
        VEC.push_back(make_cta_entry<CODE, LT, 0, false>("acs_cta<" TAG ",u16,scalar-tie>", cons)); \
        VEC.push_back(make_cta_entry<CODE, LT, 8, false>("acs_cta<" TAG ",u8,scalar-tie>", cons));  \
        VEC.push_back(make_cta_entry<CODE, LT, 0, true>("acs_cta<" TAG ",u16,simd-tie>", cons));    \
        VEC.push_back(make_cta_entry<CODE, LT, 8, true>("acs_cta<" TAG ",u8,simd-tie>", cons));     \
        VEC.push_back(make_cta_entry<CODE, LT, 0, 2>("acs_cta<" TAG ",u16,simd-sat>", cons));       \
        VEC.push_back(make_cta_entry<CODE, LT, 8, 2>("acs_cta<" TAG ",u8,simd-sat>", cons));        \
    }

template <class C, int LOGT, int SH, int TIE_SIMD, bool CONSISTENT>
KernelEntry make_entry(const char* name) {
    KernelEntry e{};
    e.K = C::K; e.R = C::R;
    for (int i = 0; i < C::R; i++) e.G[i] = C::G[i];
    e.sh = SH; e.tie = TIE_SIMD; e.consistent = CONSISTENT ? 1 : 0; e.logt = LOGT; e.name = name;
    if constexpr (LOGT == 0) {
        e.layout = LAYOUT_PAIR; e.ppw = 32; e.dec_words = 0;
        e.tagged_rows = (SH == 8 && TIE_SIMD != 2) ? 1 : 0;
        e.launch = &launch_pair<C, SH, TIE_SIMD, CONSISTENT>;
        if constexpr (DirectFetch<C, SH, PairPeriod<C>::value>::supported) e.launch_direct = &launch_pair_direct<C, SH, TIE_SIMD, CONSISTENT>;
        if constexpr (TIE_SIMD != 2) {      // the saturating flavour exists in the decision-row kernels only (acs_pair.cuh)
            e.launch_hist = &launch_hist<C, SH == 8 ? 0 : 1, TIE_SIMD, CONSISTENT, false>;
            e.launch_hist_direct = &launch_hist<C, SH == 8 ? 0 : 1, TIE_SIMD, CONSISTENT, true>;
        }
    } else {
        e.layout = LAYOUT_GROUP; e.ppw = GroupShape<C, LOGT>::PPW; e.dec_words = GroupShape<C, LOGT>::W;
        e.launch = &launch_group<C, LOGT, SH, TIE_SIMD, CONSISTENT>;
    }
    return e;
}

// all eight (metric type x tie-break x consistent-config) variants of one code at one lanes-per-pair setting
#define VITB_VARIANTS(VEC, CODE, LOGT, TAG)                                                        \
    VEC.push_back(make_entry<CODE, LOGT, 0, false, true>("acs<" TAG ",u16,scalar-tie>"));           \
    VEC.push_back(make_entry<CODE, LOGT, 8, false, true>("acs<" TAG ",u8,scalar-tie>"));            \
    VEC.push_back(make_entry<CODE, LOGT, 0, true, true>("acs<" TAG ",u16,simd-tie>"));              \
    VEC.push_back(make_entry<CODE, LOGT, 8, true, true>("acs<" TAG ",u8,simd-tie>"));               \
    VEC.push_back(make_entry<CODE, LOGT, 0, false, false>("acs<" TAG ",u16,scalar-tie,cinv>"));     \
    VEC.push_back(make_entry<CODE, LOGT, 8, false, false>("acs<" TAG ",u8,scalar-tie,cinv>"));      \
    VEC.push_back(make_entry<CODE, LOGT, 0, true, false>("acs<" TAG ",u16,simd-tie,cinv>"));        \
    VEC.push_back(make_entry<CODE, LOGT, 8, true, false>("acs<" TAG ",u8,simd-tie,cinv>"));         \
    VEC.push_back(make_entry<CODE, LOGT, 0, 2, true>("acs<" TAG ",u16,simd-sat>"));                 \
    VEC.push_back(make_entry<CODE, LOGT, 8, 2, true>("acs<" TAG ",u8,simd-sat>"));                  \
    VEC.push_back(make_entry<CODE, LOGT, 0, 2, false>("acs<" TAG ",u16,simd-sat,cinv>"));           \
    VEC.push_back(make_entry<CODE, LOGT, 8, 2, false>("acs<" TAG ",u8,simd-sat,cinv>"));

template <int K, int SH, int TIE_SIMD>
cudaError_t launch_generic(const AcsParams& p, const GenericCode& gc, cudaStream_t s) {
    constexpr unsigned W = GENERIC_THREADS / 32;
    acs_generic_kernel<K, SH, TIE_SIMD><<<(p.n_blocks + W - 1) / W, GENERIC_THREADS, 0, s>>>(p, gc);
    return cudaGetLastError();
}

template <int K, int SH, int TIE_SIMD>
KernelEntry make_generic_entry(const char* name) {
    KernelEntry e{};
    e.K = K; e.R = 0; e.sh = SH; e.tie = TIE_SIMD; e.consistent = -1; e.logt = 0; e.name = name;
    e.layout = LAYOUT_PAIR; e.ppw = 32; e.dec_words = 0; e.generic = 1;
    e.launch_generic = &launch_generic<K, SH, TIE_SIMD>;
    return e;
}

#define VITB_GENERIC_VARIANTS(VEC, KK, TAG)                                                              \
    VEC.push_back(make_generic_entry<KK, 0, false>("acs_generic<" TAG ",T1,u16,scalar-tie>"));             \
    VEC.push_back(make_generic_entry<KK, 8, false>("acs_generic<" TAG ",T1,u8,scalar-tie>"));              \
    VEC.push_back(make_generic_entry<KK, 0, true>("acs_generic<" TAG ",T1,u16,simd-tie>"));                \
    VEC.push_back(make_generic_entry<KK, 8, true>("acs_generic<" TAG ",T1,u8,simd-tie>"));

// survivor-history kernel with a frame over 4 lanes (acs_hist_group.cuh): K = 9, uint16_t metrics, whole-frame batch calls with
// directly fetchable symbols only - it has no decision-row form, so it is not one of the handle's streaming variants
template <class C, int TIE_SIMD, bool CONSISTENT>
cudaError_t launch_hist_group(const AcsParams& p, cudaStream_t s) {
    constexpr unsigned W = HistGroupShape<C>::WARPS;
    acs_hist_group_kernel<C, TIE_SIMD, CONSISTENT><<<(p.n_blocks + W - 1) / W, 32 * W, 0, s>>>(p);
    return cudaGetLastError();
}

template <class C, int TIE_SIMD, bool CONSISTENT>
KernelEntry make_hist_group_entry(const char* name) {
    KernelEntry e{};
    e.K = C::K; e.R = C::R;
    for (int i = 0; i < C::R; i++) e.G[i] = C::G[i];
    e.sh = 0; e.tie = TIE_SIMD; e.consistent = CONSISTENT ? 1 : 0; e.logt = HistGroupShape<C>::LOGT; e.name = name;
    e.layout = LAYOUT_HISTGROUP; e.ppw = 0; e.dec_words = 0;
    if (C::K == 7) e.variant_id = 104;       // 4 lanes per FRAME; 4 alone is the decision-row kernel with 4 lanes per frame pair
    e.launch_hist = &launch_hist_group<C, TIE_SIMD, CONSISTENT>;
    return e;
}

#define VITB_HIST_GROUP_VARIANTS(VEC, CODE, TAG)                                                            \
    VEC.push_back(make_hist_group_entry<CODE, false, true>("acs<" TAG ",T4,u16,scalar-tie>"));               \
    VEC.push_back(make_hist_group_entry<CODE, true, true>("acs<" TAG ",T4,u16,simd-tie>"));                  \
    VEC.push_back(make_hist_group_entry<CODE, false, false>("acs<" TAG ",T4,u16,scalar-tie,cinv>"));         \
    VEC.push_back(make_hist_group_entry<CODE, true, false>("acs<" TAG ",T4,u16,simd-tie,cinv>"));

// survivor-history kernel with one frame per 512-thread CTA (acs_hist_cta.cuh): K = 15, uint16_t metrics, whole-frame batch calls
// with unpunctured input only.  Variant number 256 (threads per frame PAIR would be 1024, which the decision-row kernel
// acs_cta<T1024> already answers to).
template <class C, int TIE_SIMD>
cudaError_t launch_hist_cta(const AcsParams& p, cudaStream_t s) {
    using H = HistCtaShape<C>;
    const cudaError_t e = cudaFuncSetAttribute(acs_hist_cta_kernel<C, TIE_SIMD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    acs_hist_cta_kernel<C, TIE_SIMD, 1><<<p.n_frames, H::T, H::SMEM_BYTES, s>>>(p);
    return cudaGetLastError();
}

template <class C, int TIE_SIMD>
KernelEntry make_hist_cta_entry(const char* name, int consistent) {
    KernelEntry e{};
    e.K = C::K; e.R = C::R;
    for (int i = 0; i < C::R; i++) e.G[i] = C::G[i];
    e.sh = 0; e.tie = TIE_SIMD; e.consistent = consistent; e.logt = HistCtaShape<C>::LOGT; e.name = name;
    e.layout = LAYOUT_HISTCTA; e.ppw = 0; e.dec_words = 0; e.variant_id = 256;
    e.launch_hist = &launch_hist_cta<C, TIE_SIMD>;
    return e;
}

#define VITB_HIST_CTA_VARIANTS(VEC, CODE, TAG)                                                          \
    for (int cons = 0; cons < 2; cons++) {                                                              \
        VEC.push_back(make_hist_cta_entry<CODE, false>("acs<" TAG ",T256,u16,scalar-tie>", cons));       \
        VEC.push_back(make_hist_cta_entry<CODE, true>("acs<" TAG ",T256,u16,simd-tie>", cons));          \
    }

// one translation unit per code family and lanes-per-pair setting (parallel compilation)
void register_small(std::vector<KernelEntry>& v);
void register_generic(std::vector<KernelEntry>& v);
void register_k9_hist_group(std::vector<KernelEntry>& v);
void register_k15_hist_cta(std::vector<KernelEntry>& v);
void register_k7_hist_group(std::vector<KernelEntry>& v);
void register_k7r2_t1(std::vector<KernelEntry>& v);
void register_k7r2_t2(std::vector<KernelEntry>& v);
void register_k7r2_t4(std::vector<KernelEntry>& v);
void register_k7r3_t1(std::vector<KernelEntry>& v);
void register_k7r3_t4(std::vector<KernelEntry>& v);
void register_k7r4_t1(std::vector<KernelEntry>& v);
void register_k7r4_t2(std::vector<KernelEntry>& v);
void register_k7r4_t4(std::vector<KernelEntry>& v);
void register_k9r2_t8(std::vector<KernelEntry>& v);
void register_k9r2_t16(std::vector<KernelEntry>& v);
void register_k9r4_t8(std::vector<KernelEntry>& v);
void register_k9r4_t16(std::vector<KernelEntry>& v);
void register_k15r6_cta512(std::vector<KernelEntry>& v);
void register_k15r6_cta1024(std::vector<KernelEntry>& v);

}  // namespace vitb
