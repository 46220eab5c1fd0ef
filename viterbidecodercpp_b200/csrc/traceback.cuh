// traceback.cuh -- device replacement for ViterbiDecoder_Core::chainback (include/viterbi/viterbi_decoder_core.h:214-236).
//
// The chain itself is serial (state_{j-1} depends on the decision read at step j), but unlike the reference layout the
// WORD ADDRESS never depends on the state here: for K <= 7 the whole decision row of a frame is one 64-bit word, so every
// load of the walk can be issued ahead of time and the kernel runs at memory speed instead of at memory latency.
// One thread per frame; 32 consecutive frames read 256 contiguous bytes per step.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vitb {

struct TracebackParams {
    const uint64_t* dec;      // [n_blocks][dec_rows][64]
    uint32_t dec_rows;
    uint32_t n_frames;
    uint32_t total_bits;      // L
    uint32_t state_bits;      // K-1
    uint32_t end_state;
    uint8_t* out;             // [n_frames][out_stride]
    size_t out_stride;
};

// Decoded bit j is the decision read from row j + (K-1) at the current state; state <- (bit << (K-2)) | (state >> 1)
// (ViterbiTracebackBuffer::push_bit_in / get_state, core.h:96-113).  Bytes are MSB-first (core.h:234).  When
// total_bits % 8 != 0 the reference's last byte carries the leading bits of end_state below the decoded bits; same here.
template <int BATCH>
__global__ void __launch_bounds__(128) traceback_u64_kernel(const TracebackParams p) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const uint32_t SB = p.state_bits, L = p.total_bits;
    const uint64_t* d = p.dec + (size_t(f >> 6) * p.dec_rows + SB) * 64 + (f & 63);   // row of decoded bit 0
    uint8_t* out = p.out + size_t(f) * p.out_stride;
    uint32_t state = p.end_state;
    int64_t j = int64_t(L) - 1;
    uint32_t byte = 0;
    if (L & 7) {   // virtual bits L, L+1, ... = end_state from its top bit down, then zeros
        for (uint32_t jj = L; jj < ((L + 7) & ~7u); jj++) {
            const uint32_t k = jj - L;
            const uint32_t b = (k < SB) ? ((p.end_state >> (SB - 1 - k)) & 1u) : 0u;
            byte |= b << (7 - (jj & 7));
        }
    }
    // ragged head so the main loop works on whole bytes
    while (j >= 0 && ((j & 7) != 7)) {
        const uint32_t bit = uint32_t(d[size_t(j) * 64] >> state) & 1u;
        state = (bit << (SB - 1)) | (state >> 1);
        byte |= bit << (7 - (uint32_t(j) & 7));
        if ((j & 7) == 0) { out[j >> 3] = uint8_t(byte); byte = 0; }
        j--;
    }
    // main loop: BATCH rows in flight
    while (j >= BATCH - 1) {
        uint64_t w[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) w[k] = __ldcs(d + size_t(j - k) * 64);
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t bit = uint32_t(w[k] >> state) & 1u;
            state = (bit << (SB - 1)) | (state >> 1);
            byte |= bit << (k & 7);                         // j-k has (j-k)&7 == 7-(k&7) because j&7 == 7 and BATCH%8 == 0
            if ((k & 7) == 7) { out[(j - k) >> 3] = uint8_t(byte); byte = 0; }
        }
        j -= BATCH;
    }
    while (j >= 0) {
        const uint32_t bit = uint32_t(d[size_t(j) * 64] >> state) & 1u;
        state = (bit << (SB - 1)) | (state >> 1);
        byte |= bit << (7 - (uint32_t(j) & 7));
        if ((j & 7) == 0) { out[j >> 3] = uint8_t(byte); byte = 0; }
        j--;
    }
}

}  // namespace vitb
