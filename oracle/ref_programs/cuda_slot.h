// cuda_slot.h -- test-only shim: adds a SIMD_CUDA entry to the reference's decoder dispatch (examples/helpers/simd_type.h:21-112)
// BY INCLUSION.  The reference header is included first, unmodified, from where it lies; then the three names its programs use to
// enumerate decoders are re-pointed:
//     SIMD_Type_List        -> a list with one more entry, SIMD_CUDA
//     SELECT_FACTORY_ITEM   -> the reference's switch plus  case SIMD_CUDA: using it = viterbi_cuda::ViterbiDecoder_CUDA_Ref<K,R,...>
//     get_simd_type_string  -> prints "SIMD_CUDA" for it
// SIMD_CUDA takes the enumerator value the reference reserves for SIMD_NEON (3), which does not exist on an x86 host and lies
// inside the value range of the enum.  A translation unit includes this header and then the reference program's .cpp itself
// (oracle/ref_programs/*.cpp); nothing of the reference is copied or edited.  Built by oracle/Makefile into oracle/_ref/.
#pragma once
#include "helpers/simd_type.h"                                  // the reference's own header (-I /root/reference/examples)
#include "viterbi_cuda/viterbi_decoder_cuda_ref.h"

#if defined(__SIMD_NEON__)
#error "the CUDA slot borrows the NEON enumerator value; build on an x86 host"
#endif
constexpr SIMD_Type SIMD_CUDA = SIMD_Type(3);

inline const std::vector<SIMD_Type>& simd_type_list_with_cuda() {
    static const std::vector<SIMD_Type> list = [] {
        std::vector<SIMD_Type> l = SIMD_Type_List;
        l.push_back(SIMD_CUDA);
        return l;
    }();
    return list;
}

template <class factory_t> struct CudaDecoderOf;
template <> struct CudaDecoderOf<ViterbiDecoder_Factory_u16> {
    template <size_t K, size_t R> using type = viterbi_cuda::ViterbiDecoder_CUDA_Ref<K, R, uint16_t, int16_t>;
};
template <> struct CudaDecoderOf<ViterbiDecoder_Factory_u8> {
    template <size_t K, size_t R> using type = viterbi_cuda::ViterbiDecoder_CUDA_Ref<K, R, uint8_t, int8_t>;
};

constexpr const char* get_simd_type_string_with_cuda(const SIMD_Type simd_type) {
    return simd_type == SIMD_CUDA ? "SIMD_CUDA" : get_simd_type_string(simd_type);
}

#undef SELECT_FACTORY_ITEM
#define SELECT_FACTORY_ITEM(FACTORY, INDEX, K, R, BLOCK) do {\
    switch (INDEX) {\
    case SIMD_Type::SCALAR:   { using it = typename FACTORY::template SCALAR<K,R>; BLOCK }; break;\
    __SELECT_FACTORY_ITEM_SSE(FACTORY, INDEX, K, R, BLOCK)\
    __SELECT_FACTORY_ITEM_AVX(FACTORY, INDEX, K, R, BLOCK)\
    case SIMD_CUDA:           { using it = typename CudaDecoderOf<FACTORY>::template type<K,R>; BLOCK }; break;\
    default: break;\
    }\
} while(0)

#define SIMD_Type_List simd_type_list_with_cuda()
#define get_simd_type_string get_simd_type_string_with_cuda
