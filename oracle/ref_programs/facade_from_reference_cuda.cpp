// The facade's converting constructors: a program that already holds the REFERENCE'S ViterbiBranchTable / ViterbiDecoder_Config
// objects hands them to the CUDA facade as they are, decodes a batch on the GPU, and gets the bytes and error the reference's scalar
// decoder gets.  Also checks operator[] / data() of the facade table and the m_metrics / m_decisions / m_current_decoded_bit accessors
// against the reference object's public members (core.h:238-242) after a streaming decode of the same frame.
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <random>
#include <vector>
#include "viterbi/convolutional_encoder_shift_register.h"
#include "viterbi/viterbi_decoder_core.h"
#include "viterbi/viterbi_decoder_scalar.h"
#include "helpers/test_helpers.h"
#include "viterbi_cuda/viterbi_decoder_cuda.h"

template <size_t K, size_t R, typename error_t, typename soft_t>
static int run(const char* name, const uint32_t (&G)[R], soft_t high, soft_t low, error_t threshold, unsigned seed) {
    const size_t n_bytes = 96, n_bits = n_bytes * 8, n_sym = (n_bits + K - 1) * R, F = 37;
    const error_t max_error = error_t((high - low) * int(R));
    ::ViterbiDecoder_Config<error_t> ref_config;
    ref_config.soft_decision_max_error = max_error;
    ref_config.initial_start_error = 0;
    ref_config.initial_non_start_error = error_t(max_error * 3);
    ref_config.renormalisation_threshold = threshold;
    ::ViterbiBranchTable<K, R, soft_t> ref_table(G, high, low);
    ::ViterbiDecoder_Core<K, R, error_t, soft_t> ref_core(ref_table, ref_config);

    // the reference's objects, handed over as they are
    viterbi_cuda::ViterbiBranchTable<K, R, soft_t> table(ref_table);
    viterbi_cuda::ViterbiDecoder_Config<error_t> config(ref_config);
    viterbi_cuda::ViterbiDecoder_Core<K, R, error_t, soft_t> core(table, config);
    using Cuda = viterbi_cuda::ViterbiDecoder_CUDA<K, R, error_t, soft_t>;
    using Scalar = ::ViterbiDecoder_Scalar<K, R, error_t, soft_t>;
    int bad = 0;
    for (size_t i = 0; i < R; i++) {
        if (memcmp(table[i], ref_table[i], sizeof(soft_t) * table.NUMSTATES) != 0) { printf("%s: table row %zu differs\n", name, i); bad++; }
        if ((table.polynomials()[i] ^ G[i]) & ((1u << (K - 1)) - 2u)) { printf("%s: polynomial %zu not recovered\n", name, i); bad++; }
    }
    if (table.data() != table[0]) bad++;

    std::mt19937 rng(seed);
    std::vector<soft_t> symbols(F * n_sym);
    std::vector<uint8_t> tx(F * n_bytes), ref_out(F * n_bytes), out(F * n_bytes);
    std::vector<uint64_t> ref_acc(F), acc(F);
    std::vector<uint32_t> ref_fin(F), fin(F);
    auto enc = ConvolutionalEncoder_ShiftRegister<uint32_t>(K, R, G);
    std::normal_distribution<float> noise(0.0f, 0.45f * float(high - low) / 2.0f);
    for (size_t f = 0; f < F; f++) {
        for (size_t i = 0; i < n_bytes; i++) tx[f * n_bytes + i] = uint8_t(rng());
        encode_data(&enc, &tx[f * n_bytes], n_bytes, &symbols[f * n_sym], n_sym, high, low);
        for (size_t i = 0; i < n_sym; i++) {
            float v = float(symbols[f * n_sym + i]) + noise(rng);
            v = v > float(high) ? float(high) : (v < float(low) ? float(low) : v);
            symbols[f * n_sym + i] = soft_t(v >= 0 ? v + 0.5f : v - 0.5f);
        }
        ref_core.set_traceback_length(n_bits);
        ref_core.reset();
        ref_acc[f] = Scalar::template update<uint64_t>(ref_core, &symbols[f * n_sym], n_sym);
        ref_fin[f] = uint32_t(ref_core.get_error());
        ref_core.chainback(&ref_out[f * n_bytes], n_bits, 0u);
    }
    Cuda::decode_batch(core, symbols.data(), F, n_bits, out.data(), acc.data(), fin.data());
    for (size_t f = 0; f < F; f++) {
        if (memcmp(&out[f * n_bytes], &ref_out[f * n_bytes], n_bytes) != 0 || acc[f] != ref_acc[f] || fin[f] != ref_fin[f]) {
            printf("%s: frame %zu differs from the reference scalar decoder\n", name, f);
            bad++;
        }
    }
    // streaming decode of the last frame: the public members of the reference object, read through the facade
    core.set_traceback_length(n_bits);
    core.reset();
    (void)Cuda::template update<uint64_t>(core, &symbols[(F - 1) * n_sym], n_sym);
    if (core.m_current_decoded_bit() != ref_core.m_current_decoded_bit) { printf("%s: m_current_decoded_bit differs\n", name); bad++; }
    const std::vector<error_t> m = core.m_metrics();
    for (size_t s = 0; s < core.NUMSTATES; s++) if (m[s] != ref_core.m_metrics.get_old()[s]) { printf("%s: m_metrics[%zu] differs\n", name, s); bad++; break; }
    const size_t rows = n_bits + K - 1;
    const std::vector<uint64_t> d = core.m_decisions(0, rows);
    for (size_t r = 0; r < rows && !bad; r++)
        for (size_t w = 0; w < ::ViterbiDecoder_Core<K, R, error_t, soft_t>::Decisions::TOTAL_BLOCKS; w++)
            if (d[r * core.DECISION_WORDS + w] != uint64_t(ref_core.m_decisions[r][w])) { printf("%s: m_decisions[%zu][%zu] differs\n", name, r, w); bad++; break; }
    printf("%s %s: %zu frames, table, polynomials, bytes, errors, metrics, decisions\n", bad ? "FAIL" : "PASS", name, F);
    return bad;
}

int main() {
    int bad = 0;
    const uint32_t voyager[2] = {109, 79}, dab[4] = {109, 79, 83, 109}, is95a[2] = {491, 369}, lte[3] = {91, 121, 117};
    bad += run<7, 2, uint16_t, int16_t>("Voyager soft16", voyager, int16_t(127), int16_t(-127), uint16_t(62995), 1);
    bad += run<7, 2, uint8_t, int8_t>("Voyager hard8", voyager, int8_t(1), int8_t(-1), uint8_t(243), 2);
    bad += run<7, 4, uint16_t, int16_t>("DAB soft16", dab, int16_t(127), int16_t(-127), uint16_t(60455), 3);
    bad += run<7, 3, uint8_t, int8_t>("LTE soft8", lte, int8_t(3), int8_t(-3), uint8_t(200), 4);
    bad += run<9, 2, uint16_t, int16_t>("IS-95A soft16", is95a, int16_t(127), int16_t(-127), uint16_t(62995), 5);
    return bad ? 1 : 0;
}
