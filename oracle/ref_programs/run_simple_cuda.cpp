// The reference's examples/run_simple.cpp, unmodified, decoding on the GPU: the one line that names its decoder class
// (run_simple.cpp:74, "using Decoder = ViterbiDecoder_Scalar<K,R,uint16_t,int16_t>") is re-pointed at the CUDA decoder class.
#include <stddef.h>
#include <stdint.h>
#include "viterbi/viterbi_decoder_scalar.h"                     // the real one first (it is #pragma once), so the rename below hits only the program
#include "viterbi_cuda/viterbi_decoder_cuda_ref.h"
#define ViterbiDecoder_Scalar viterbi_cuda::ViterbiDecoder_CUDA_Ref
#include "run_simple.cpp"
