// The reference's examples/run_tests.cpp (8 codes x 3 decode types x every decoder of the host), unmodified, with the CUDA backend
// as one more decoder: every (code, type) is also decoded by SIMD_CUDA.  See cuda_slot.h.
#include "cuda_slot.h"
#include "run_tests.cpp"                                        // the reference program itself (-I /root/reference/examples)

// The CUDA backend reproduces the SCALAR decoder bit for bit, including the uint8_t metric overflow of Cassini with SOFT8 levels that
// makes the reference skip that case for SCALAR (run_tests.cpp:63-65): same skip for SIMD_CUDA.
static const bool cuda_skip_registered = [] {
    SKIP_TESTS.emplace(TestKey(SIMD_CUDA, DecodeType::SOFT8, 15, 6), "Same wrapping arithmetic as SCALAR (bit-exact with it): overflow in metrics");
    return true;
}();
