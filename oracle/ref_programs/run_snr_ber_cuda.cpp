// The reference's examples/run_snr_ber.cpp (BER sweep over Eb/N0, one decoder per pool thread: run_snr_ber.cpp:255-275), unmodified,
// with SIMD_CUDA as one more decoder (select it with "-s simd_cuda").  The CUDA decoder class keeps one GPU handle per host thread
// (viterbi_decoder_cuda_ref.h), so the reference's thread pool can drive it as it drives its own decoders.  See cuda_slot.h.
#include "cuda_slot.h"
#include "run_snr_ber.cpp"
