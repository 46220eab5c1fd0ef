// run_simple_cuda.cpp -- the reference's usage demo (examples/run_simple.cpp: K=7 R=4 {109,79,83,109}, 1024 bytes, SOFT16,
// noise free) written against the CUDA facade, plus the same data decoded as a batch.  Own code: the encoder below is a plain
// shift register (reg = reg << 1 | bit, output i = parity(reg & G[i]), K-1 zero tail bits).
// Build: g++ -std=c++17 -Iinclude examples/run_simple_cuda.cpp -Lviterbidecodercpp_b200/csrc -lviterbi_b200 -o run_simple_cuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "viterbi_cuda/viterbi_decoder_cuda.h"

int main() {
    constexpr size_t K = 7, R = 4;
    const uint8_t G[R] = {109, 79, 83, 109};
    const int16_t high = +127, low = -127;
    const size_t total_input_bytes = 1024, total_input_bits = total_input_bytes * 8, steps = total_input_bits + K - 1;

    std::vector<uint8_t> tx(total_input_bytes), rx(total_input_bytes);
    for (auto& b : tx) b = uint8_t(std::rand() % 256);
    std::vector<int16_t> symbols(steps * R);
    uint32_t reg = 0;
    for (size_t t = 0; t < steps; t++) {
        const uint32_t bit = t < total_input_bits ? ((tx[t / 8] >> (7 - t % 8)) & 1u) : 0u;
        reg = (reg << 1) | bit;
        for (size_t i = 0; i < R; i++) symbols[t * R + i] = __builtin_parity(reg & G[i]) ? high : low;
    }

    // decode_type.h:21-34 (SOFT16 preset)
    const uint16_t max_error = uint16_t(high - low) * uint16_t(R), margin = uint16_t(max_error * 5u);
    viterbi_cuda::ViterbiDecoder_Config<uint16_t> config{max_error, 0, margin, uint16_t(0xFFFF - margin)};

    using Decoder = viterbi_cuda::ViterbiDecoder_CUDA<K, R, uint16_t, int16_t>;
    auto branch_table = viterbi_cuda::ViterbiBranchTable<K, R, int16_t>(G, high, low);
    auto vitdec = viterbi_cuda::ViterbiDecoder_Core<K, R, uint16_t, int16_t>(branch_table, config);

    vitdec.set_traceback_length(total_input_bits);                       // run_simple.cpp:76
    vitdec.reset();                                                      // :77
    const uint64_t accumulated_error = Decoder::update<uint64_t>(vitdec, symbols.data(), symbols.size());   // :78
    const uint64_t error = accumulated_error + uint64_t(vitdec.get_error());                              // :79
    vitdec.chainback(rx.data(), total_input_bits, 0u);                   // :80
    size_t bit_errors = 0;
    for (size_t i = 0; i < total_input_bytes; i++) bit_errors += size_t(__builtin_popcount(tx[i] ^ rx[i]));
    printf("error_metric=%llu\n%zu/%zu incorrect bits\n", (unsigned long long)error, bit_errors, total_input_bits);

    // the same frame 100 times as a batch
    const size_t F = 100;
    std::vector<int16_t> batch(F * symbols.size());
    for (size_t f = 0; f < F; f++) std::copy(symbols.begin(), symbols.end(), batch.begin() + f * symbols.size());
    std::vector<uint8_t> out(F * total_input_bytes);
    std::vector<uint64_t> acc(F);
    std::vector<uint32_t> fin(F);
    Decoder::decode_batch(vitdec, batch.data(), F, total_input_bits, out.data(), acc.data(), fin.data());
    size_t bad = 0;
    for (size_t f = 0; f < F; f++)
        for (size_t i = 0; i < total_input_bytes; i++) bad += (out[f * total_input_bytes + i] != tx[i]);
    printf("batch: %zu mismatching bytes, error_metric[0]=%llu\n", bad, (unsigned long long)(acc[0] + fin[0]));
    return (bit_errors == 0 && bad == 0 && error == 0) ? 0 : 1;
}
