// frontend.cuh -- device-side producers of decoder input: the test-data pipeline of the reference's
// BER sweep (examples/run_snr_ber.cpp:311-359) as CUDA kernels, so that sweeps run without crossing PCIe:
//     random data bytes -> rate 1/R shift-register convolutional encoder with K-1 zero tail bits
//     (include/viterbi/convolutional_encoder_shift_register.h:45-61, examples/helpers/test_helpers.h:17-64)
//     -> +-1.0 -> + N(0, sigma^2) -> x * (mag / sqrt(1 + sigma^2)) + mean -> round -> clamp to [low, high] -> soft_t
//     [-> puncture: only the symbols whose keep[] flag is set are stored, examples/helpers/puncture_code_helpers.h:57-98].
// The random streams are Philox4x32-10 (counter based: frame and position select the counter), not std::mt19937, so the
// frames are statistically - not bitwise - those of the reference program; the encoder and the quantiser are exact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <curand_kernel.h>

namespace vitb {

struct SynthParams {
    int K, R;
    uint32_t G[16];
    int high, low;
    uint32_t n_frames, total_bits;        // data bits per frame (multiple of 8)
    float sigma;                          // noise standard deviation; < 0: noise free
    float scale, mean;                    // quantiser: soft = round(x * scale + mean)
    unsigned long long seed;
    const int32_t* depuncture_map;        // nullable: [steps*R] -> index in the stored row or -1 (symbol not transmitted)
    uint8_t* tx_bytes;                    // [n_frames][total_bits/8]
    void* symbols;                        // [n_frames][row_stride] soft_t
    size_t row_stride;
};

// one thread per data byte
__global__ void synth_bytes_kernel(const SynthParams p) {
    const size_t nb = size_t(p.total_bits) / 8, i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nb * p.n_frames) return;
    curandStatePhilox4_32_10_t st;
    curand_init(p.seed, /*subsequence*/ i, /*offset*/ 0, &st);
    p.tx_bytes[i] = uint8_t(curand(&st) & 0xffu);
}

// reference quantiser (run_snr_ber.cpp:352-359); separate multiply and add so the result matches a two-rounding host evaluation
__device__ __forceinline__ int quantise_soft(float x, float scale, float mean, int low, int high) {
    const float y = roundf(__fadd_rn(__fmul_rn(x, scale), mean));       // std::round: halves away from zero
    int v = int(y);
    v = v > high ? high : v;
    v = v < low ? low : v;
    return v;
}

// one thread per (frame, trellis step): R symbols
template <typename soft_t>
__global__ void synth_symbols_kernel(const SynthParams p) {
    const uint32_t steps = p.total_bits + uint32_t(p.K) - 1;
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= size_t(steps) * p.n_frames) return;
    const uint32_t f = uint32_t(i / steps), t = uint32_t(i % steps);
    const uint8_t* tx = p.tx_bytes + size_t(f) * (p.total_bits / 8);
    // shift register after step t: bit k = data bit t-k (MSB-first within bytes, zero before the frame and in the tail)
    uint32_t reg = 0;
    for (int k = 0; k < p.K; k++) {
        const int64_t b = int64_t(t) - k;
        uint32_t bit = 0;
        if (b >= 0 && b < int64_t(p.total_bits)) bit = (tx[b >> 3] >> (7 - (b & 7))) & 1u;
        reg |= bit << k;
    }
    curandStatePhilox4_32_10_t st;
    if (p.sigma >= 0.f) curand_init(p.seed ^ 0x9e3779b97f4a7c15ull, /*subsequence*/ i, 0, &st);
    soft_t* row = static_cast<soft_t*>(p.symbols) + size_t(f) * p.row_stride;
    for (int r = 0; r < p.R; r++) {
        const uint32_t code_bit = uint32_t(__popc(reg & p.G[r])) & 1u;
        float x = code_bit ? 1.0f : -1.0f;
        if (p.sigma >= 0.f) x += p.sigma * curand_normal(&st);
        const uint32_t e = t * uint32_t(p.R) + uint32_t(r);
        int32_t dst = p.depuncture_map ? p.depuncture_map[e] : int32_t(e);
        if (dst >= 0) row[dst] = soft_t(quantise_soft(x, p.scale, p.mean, p.low, p.high));
    }
}

template <typename soft_t>
__global__ void quantise_kernel(const float* x, size_t n, float scale, float mean, int low, int high, soft_t* out) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = soft_t(quantise_soft(x[i], scale, mean, low, high));
}

// number of differing bits between two byte arrays (get_total_bit_errors, examples/helpers/test_helpers.h:94-103)
__global__ void bit_errors_kernel(const uint8_t* a, const uint8_t* b, size_t n, unsigned long long* count) {
    unsigned long long local = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        local += unsigned(__popc(uint32_t(a[i] ^ b[i])));
    for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

}  // namespace vitb
