// viterbi_decoder_cuda.h -- header-only C++17 facade over the C ABI (include/viterbi_b200.h) in the shapes of the reference:
//
//   reference                                                        here
//   ---------------------------------------------------------------  -----------------------------------------------------
//   ViterbiBranchTable<K,R,soft_t>(G, high, low)                     viterbi_cuda::ViterbiBranchTable<K,R,soft_t>      (same ctor)
//     include/viterbi/viterbi_branch_table.h:19-74
//   ViterbiDecoder_Config<error_t>                                   viterbi_cuda::ViterbiDecoder_Config<error_t>      (same fields)
//     include/viterbi/viterbi_decoder_config.h:11-18
//   ViterbiDecoder_Core<K,R,error_t,soft_t>(table, config)           viterbi_cuda::ViterbiDecoder_Core<K,R,error_t,soft_t>
//     .set_traceback_length/.get_traceback_length/.get_error/          same member names and default arguments
//     .reset/.chainback   include/viterbi/viterbi_decoder_core.h:157-243
//   ViterbiDecoder_Scalar<K,R,error_t,soft_t>::update<sum_t>(...)    viterbi_cuda::ViterbiDecoder_CUDA<K,R,error_t,soft_t>::update<sum_t>(...)
//     include/viterbi/viterbi_decoder_scalar.h:28-55                   + ::is_valid, + ::decode_batch (the batched entry point)
//
// A program written against the reference switches backend by replacing the three class names; see INTEGRATION.md.
// The reference reports misuse with assert(); here the C ABI returns an error code and the facade throws std::runtime_error
// (or aborts when exceptions are disabled), so a violated precondition is never silent.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include "../viterbi_b200.h"

namespace viterbi_cuda {

inline void check(int status, const char* what) {
    if (status != VITB_OK) throw std::runtime_error(std::string(what) + ": " + vitb_status_string(status));
}

template <typename error_t>
struct ViterbiDecoder_Config {            // viterbi_decoder_config.h:11-18
    error_t soft_decision_max_error;
    error_t initial_start_error;
    error_t initial_non_start_error;
    error_t renormalisation_threshold;
};

// On the GPU the table is folded into the kernels at compile time; this object only carries what the reference constructor takes.
template <size_t constraint_length, size_t code_rate, typename soft_t>
class ViterbiBranchTable {
public:
    static constexpr size_t K = constraint_length;
    static constexpr size_t R = code_rate;
    static constexpr size_t NUMSTATES = (size_t(1) << (K - 1)) / 2;      // viterbi_branch_table.h:26
    template <typename code_t>
    ViterbiBranchTable(const code_t* G, const soft_t _soft_decision_high, const soft_t _soft_decision_low)
        : soft_decision_high(_soft_decision_high), soft_decision_low(_soft_decision_low) {
        static_assert(K > 1u && R > 1u && R <= VITB_MAX_R, "unsupported code");
        for (size_t i = 0; i < R; i++) m_G[i] = uint32_t(G[i]);
    }
    // BT[i][j] = parity((j << 1) & G[i]) ? high : low   (viterbi_branch_table.h:45-54), computed on demand
    soft_t at(size_t index, size_t state) const {
        uint32_t v = (uint32_t(state) << 1) & m_G[index];
        v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
        return (v & 1u) ? soft_decision_high : soft_decision_low;
    }
    const uint32_t* polynomials() const { return m_G; }
    const soft_t soft_decision_high;
    const soft_t soft_decision_low;
private:
    uint32_t m_G[VITB_MAX_R] = {};
};

template <size_t constraint_length, size_t code_rate, typename error_t, typename soft_t>
class ViterbiDecoder_Core {
public:
    static constexpr size_t K = constraint_length;
    static constexpr size_t R = code_rate;
    static constexpr size_t TOTAL_STATE_BITS = K - 1;
    static constexpr size_t NUMSTATES = size_t(1) << TOTAL_STATE_BITS;
    using BranchTable = ViterbiBranchTable<K, R, soft_t>;
    using Config = ViterbiDecoder_Config<error_t>;
    static_assert((std::is_same<error_t, uint16_t>::value && std::is_same<soft_t, int16_t>::value) ||
                  (std::is_same<error_t, uint8_t>::value && std::is_same<soft_t, int8_t>::value),
                  "CUDA backend supports <uint16_t,int16_t> and <uint8_t,int8_t>");

    ViterbiDecoder_Core(const BranchTable& branch_table, const Config& config, int device = 0, int tie_break = VITB_TIE_SCALAR)
        : m_config(config) {
        const vitb_params p = make_params(branch_table, config, device, tie_break);
        check(vitb_create(&p, &m_handle), "vitb_create");       // leaves the decoder reset with traceback length 0 (core.h:175-176)
    }
    ~ViterbiDecoder_Core() { vitb_destroy(m_handle); }
    ViterbiDecoder_Core(const ViterbiDecoder_Core&) = delete;
    ViterbiDecoder_Core& operator=(const ViterbiDecoder_Core&) = delete;

    static vitb_params make_params(const BranchTable& bt, const Config& c, int device, int tie_break) {
        vitb_params p{};
        p.K = int32_t(K); p.R = int32_t(R);
        for (size_t i = 0; i < R; i++) p.G[i] = bt.polynomials()[i];
        p.soft_bytes = int32_t(sizeof(soft_t));
        p.soft_decision_high = bt.soft_decision_high; p.soft_decision_low = bt.soft_decision_low;
        p.soft_decision_max_error = c.soft_decision_max_error; p.initial_start_error = c.initial_start_error;
        p.initial_non_start_error = c.initial_non_start_error; p.renormalisation_threshold = c.renormalisation_threshold;
        p.tie_break = tie_break; p.device = device;
        return p;
    }

    void set_traceback_length(const size_t traceback_length) { check(vitb_set_traceback_length(m_handle, traceback_length), "set_traceback_length"); }
    size_t get_traceback_length() const { size_t n = 0; check(vitb_get_traceback_length(m_handle, &n), "get_traceback_length"); return n; }
    error_t get_error(const size_t end_state = 0u) { uint32_t e = 0; check(vitb_get_error(m_handle, end_state, &e), "get_error"); return error_t(e); }
    void reset(const size_t starting_state = 0u) { check(vitb_reset(m_handle, starting_state), "reset"); }
    void chainback(uint8_t* bytes_out, const size_t total_bits, const size_t end_state = 0u) {
        check(vitb_chainback(m_handle, bytes_out, total_bits, end_state), "chainback");
    }
    size_t current_decoded_bit() const { size_t n = 0; check(vitb_get_current_decoded_bit(m_handle, &n), "current_decoded_bit"); return n; }
    vitb_decoder* handle() { return m_handle; }
    const Config m_config;
private:
    vitb_decoder* m_handle = nullptr;
};

// The stateless decoder class of the reference (ViterbiDecoder_Scalar / _SSE_u16 / _AVX_u16 ...): same static interface.
template <size_t constraint_length, size_t code_rate, typename error_t, typename soft_t>
class ViterbiDecoder_CUDA {
public:
    using Base = ViterbiDecoder_Core<constraint_length, code_rate, error_t, soft_t>;
    // compile-time part of is_valid: the type pair; whether kernels exist for the polynomials is a run-time question (is_supported)
    static constexpr bool is_valid = Base::K >= 3;
    static bool is_supported(const typename Base::BranchTable& bt, const typename Base::Config& c) {
        const vitb_params p = Base::make_params(bt, c, 0, VITB_TIE_SCALAR);
        return vitb_is_supported(&p) != 0;
    }
    template <typename sum_error_t>
    static sum_error_t update(Base& base, const soft_t* symbols, const size_t N) {     // scalar.h:28-55
        uint64_t acc = 0;
        check(vitb_update(base.handle(), symbols, N, &acc), "update");
        return sum_error_t(acc);
    }
    // Batched entry point: n_frames frames of (total_bits + K-1)*R symbols each, decoded as reset() + update() + get_error() + chainback()
    static void decode_batch(Base& base, const soft_t* symbols, const size_t n_frames, const size_t total_bits,
                             uint8_t* bytes_out, uint64_t* accumulated_error, uint32_t* final_error) {
        check(vitb_decode_batch(base.handle(), symbols, n_frames, total_bits, nullptr, bytes_out, accumulated_error, final_error), "decode_batch");
    }
};

}  // namespace viterbi_cuda
