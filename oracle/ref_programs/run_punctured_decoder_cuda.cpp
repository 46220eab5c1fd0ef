// The reference's examples/run_punctured_decoder.cpp (DAB fast information channel: 774 update calls per frame through
// decode_punctured_symbols, helpers/puncture_code_helpers.h:17-55), unmodified, with SIMD_CUDA as one more decoder.  See cuda_slot.h.
#include "cuda_slot.h"
#include "run_punctured_decoder.cpp"
