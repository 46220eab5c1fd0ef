// The reference's examples/run_benchmark.cpp (symbol rate of update / chainback per decoder: run_benchmark.cpp:188-245), unmodified,
// with SIMD_CUDA as one more decoder ("-s simd_cuda").  It times the single-frame streaming calls - the compatibility path; the
// batched entry points are what bench.py times.  See cuda_slot.h.
#include "cuda_slot.h"
#include "run_benchmark.cpp"
