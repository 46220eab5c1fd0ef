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
#include <utility>
#include <vector>
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
    ViterbiDecoder_Config() = default;
    ViterbiDecoder_Config(error_t max_error, error_t start, error_t non_start, error_t threshold)
        : soft_decision_max_error(max_error), initial_start_error(start), initial_non_start_error(non_start), renormalisation_threshold(threshold) {}
    // from the reference's own ::ViterbiDecoder_Config<error_t> (any type with the same four fields)
    template <class RefConfig, typename = decltype(std::declval<const RefConfig&>().renormalisation_threshold)>
    ViterbiDecoder_Config(const RefConfig& c)
        : soft_decision_max_error(c.soft_decision_max_error), initial_start_error(c.initial_start_error),
          initial_non_start_error(c.initial_non_start_error), renormalisation_threshold(c.renormalisation_threshold) {}
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
        fill();
    }
    // From the reference's own ::ViterbiBranchTable<K,R,soft_t> (any type with its operator[]: table[i][j], j < 2^(K-2)).  The table
    // holds bits 1 .. K-2 of every polynomial; bit 0 (input tap) and bit K-1 (oldest bit) never enter it - the decoder assumes both
    // set.  Entry [i][0] is always the low soft level (parity of 0), any other value the high one (both are private in the reference).
    template <class RefTable, typename = decltype(std::declval<const RefTable&>()[0][0]), typename = typename std::enable_if<!std::is_pointer<RefTable>::value>::type>
    explicit ViterbiBranchTable(const RefTable& ref) : soft_decision_high(high_of(ref)), soft_decision_low(ref[0][0]) {
        for (size_t i = 0; i < R; i++) {
            uint32_t g = 1u | (1u << (K - 1));
            for (size_t b = 1; b + 1 < K; b++)
                if (ref[i][size_t(1) << (b - 1)] != soft_decision_low) g |= 1u << b;
            m_G[i] = g;
        }
        fill();
    }
    // BT[i][j] = parity((j << 1) & G[i]) ? high : low   (viterbi_branch_table.h:45-54)
    soft_t at(size_t index, size_t state) const { return m_table[index * NUMSTATES + state]; }
    const soft_t* operator[](size_t index) const { return &m_table[index * NUMSTATES]; }        // viterbi_branch_table.h:59-62
    const soft_t* data() const { return m_table.data(); }                                       // viterbi_branch_table.h:64-66 (rows are contiguous here)
    const uint32_t* polynomials() const { return m_G; }
    const soft_t soft_decision_high;
    const soft_t soft_decision_low;
private:
    template <class RefTable>
    static soft_t high_of(const RefTable& ref) {
        const soft_t lo = ref[0][0];
        soft_t hi = lo;
        for (size_t i = 0; i < R; i++)
            for (size_t j = 0; j < NUMSTATES; j++)
                if (ref[i][j] != lo) hi = ref[i][j];
        return hi;
    }
    void fill() {
        m_table.resize(R * NUMSTATES);
        for (size_t i = 0; i < R; i++) {
            for (size_t state = 0; state < NUMSTATES; state++) {
                uint32_t v = (uint32_t(state) << 1) & m_G[i];
                v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
                m_table[i * NUMSTATES + state] = (v & 1u) ? soft_decision_high : soft_decision_low;
            }
        }
    }
    uint32_t m_G[VITB_MAX_R] = {};
    std::vector<soft_t> m_table;       // host copy for callers that index the table; the kernels fold G into their code
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
    // The reference's public data members (core.h:238-242) live on the device here; these read them out / assign them.
    size_t m_current_decoded_bit() const { return current_decoded_bit(); }
    void set_current_decoded_bit(size_t bit) { check(vitb_set_current_decoded_bit(m_handle, bit), "set_current_decoded_bit"); }
    std::vector<error_t> m_metrics() {                                          // m_metrics.get_old()[0 .. NUMSTATES)
        std::vector<uint32_t> raw(NUMSTATES);
        check(vitb_get_metrics(m_handle, raw.data()), "get_metrics");
        return std::vector<error_t>(raw.begin(), raw.end());
    }
    void set_metrics(const error_t* metrics) {
        std::vector<uint32_t> raw(metrics, metrics + NUMSTATES);
        check(vitb_set_metrics(m_handle, raw.data()), "set_metrics");
    }
    // m_decisions[first_row .. first_row + n_rows) in the reference layout: max(NUMSTATES/64, 1) uint64 per row (core.h:49-83)
    static constexpr size_t DECISION_WORDS = (NUMSTATES + 63) / 64;
    std::vector<uint64_t> m_decisions(size_t first_row, size_t n_rows) {
        std::vector<uint64_t> rows(n_rows * DECISION_WORDS);
        if (n_rows) check(vitb_get_decisions(m_handle, first_row, n_rows, rows.data()), "get_decisions");
        return rows;
    }
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
