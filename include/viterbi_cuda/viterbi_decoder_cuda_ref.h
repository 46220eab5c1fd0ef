// viterbi_decoder_cuda_ref.h -- the CUDA backend as ONE MORE DECODER CLASS OF THE REFERENCE: same static interface as
// ViterbiDecoder_Scalar / _SSE_u16 / _AVX_u16 (include/viterbi/viterbi_decoder_scalar.h:25-55), operating on the REFERENCE'S OWN
// ViterbiDecoder_Core object.  This is the "SIMD_CUDA" slot of examples/helpers/simd_type.h:21-112: with it the reference's
// programs (run_tests.cpp, run_punctured_decoder.cpp, run_simple.cpp) drive the GPU unmodified - see oracle/ref_programs/.
//
//     using Decoder = viterbi_cuda::ViterbiDecoder_CUDA_Ref<K, R, uint16_t, int16_t>;
//     ViterbiDecoder_Core<K, R, uint16_t, int16_t> vitdec(branch_table, config);          // the reference's class
//     vitdec.set_traceback_length(L);  vitdec.reset();
//     acc = Decoder::update<uint64_t>(vitdec, symbols, n);                                // add-compare-select on the GPU
//     err = vitdec.get_error();  vitdec.chainback(bytes, L, 0);                           // the reference's own code, on GPU-made state
//
// The decoder state stays where the reference keeps it - m_metrics, m_decisions, m_current_decoded_bit are public members
// (core.h:238-242).  update() hands the current metrics and row counter to a GPU handle (vitb_set_metrics,
// vitb_set_current_decoded_bit), runs vitb_update, and copies the new metrics and the new decision rows (reference bit order,
// vitb_get_decisions) back into the object; get_error and chainback then run as the reference wrote them.  One GPU handle per
// (code, soft levels, config) AND PER HOST THREAD is created on first use and shared by every Core of that type the thread decodes
// (the handle holds no state between calls; per thread because the reference's programs decode from a thread pool,
// examples/run_snr_ber.cpp:260-264, and a handle is single-threaded).  It is a compatibility path (one host round trip per update
// call); batches go through viterbi_decoder_cuda.h.
//
// Needs the reference headers on the include path (<viterbi/viterbi_decoder_core.h>); nothing of them is copied here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>
#include "viterbi/viterbi_decoder_core.h"
#include "../viterbi_b200.h"

namespace viterbi_cuda {

// The branch table holds parity((j << 1) & G[i]) for j < 2^(K-2) (viterbi_branch_table.h:45-54): bits 1 .. K-2 of every polynomial.
// Bit 0 (the input tap) and bit K-1 (the tap on the oldest bit, the one the two butterfly predecessors differ in) never enter the
// table - the decoder assumes both are set, as they are in every code of constraint length K.
// The two soft levels are private members of the reference table; entry [i][0] is always the low one (parity of 0) and any other
// value in the table is the high one.
template <size_t K, size_t R, typename soft_t>
inline void polynomials_of(const ::ViterbiBranchTable<K, R, soft_t>& bt, uint32_t (&G)[VITB_MAX_R], int32_t& high, int32_t& low) {
    static_assert(R <= VITB_MAX_R, "code rate not supported by the CUDA backend");
    using Table = ::ViterbiBranchTable<K, R, soft_t>;
    const soft_t lo = bt[0][0];
    soft_t hi = lo;
    for (size_t i = 0; i < R; i++)
        for (size_t j = 0; j < Table::NUMSTATES; j++)
            if (bt[i][j] != lo) hi = bt[i][j];
    for (size_t i = 0; i < R; i++) {
        uint32_t g = 1u | (1u << (K - 1));
        for (size_t b = 1; b + 1 < K; b++)
            if (bt[i][size_t(1) << (b - 1)] != lo) g |= 1u << b;
        G[i] = g;
    }
    high = int32_t(hi);
    low = int32_t(lo);
}

template <size_t constraint_length, size_t code_rate, typename error_t, typename soft_t>
class ViterbiDecoder_CUDA_Ref {
public:
    using Base = ::ViterbiDecoder_Core<constraint_length, code_rate, error_t, soft_t>;
    static constexpr size_t K = constraint_length, R = code_rate;
    static constexpr bool is_valid =
        ((std::is_same<error_t, uint16_t>::value && std::is_same<soft_t, int16_t>::value) ||
         (std::is_same<error_t, uint8_t>::value && std::is_same<soft_t, int8_t>::value)) && K >= 3 && R <= VITB_MAX_R;

    template <typename sum_error_t>
    static sum_error_t update(Base& base, const soft_t* symbols, const size_t N) {              // scalar.h:28-55
        vitb_decoder* h = handle_for(base);
        const size_t steps = N / R, first = base.m_current_decoded_bit;
        size_t tl = 0;
        ok(vitb_get_traceback_length(h, &tl), "get_traceback_length");
        if (tl != base.get_traceback_length()) ok(vitb_set_traceback_length(h, base.get_traceback_length()), "set_traceback_length");
        // state in: the reference object's metrics (whatever reset() or an earlier update left there) and its row counter
        uint32_t m[Base::NUMSTATES];
        error_t* old_metrics = base.m_metrics.get_old();
        for (size_t s = 0; s < Base::NUMSTATES; s++) m[s] = uint32_t(old_metrics[s]);
        ok(vitb_set_metrics(h, m), "set_metrics");
        ok(vitb_set_current_decoded_bit(h, first), "set_current_decoded_bit");
        uint64_t acc = 0;
        ok(vitb_update(h, symbols, N, &acc), "update");
        // state out: new metrics, and the decision rows of the steps just taken in the reference's layout (core.h:49-83)
        ok(vitb_get_metrics(h, m), "get_metrics");
        for (size_t s = 0; s < Base::NUMSTATES; s++) old_metrics[s] = error_t(m[s]);
        constexpr size_t W64 = (Base::NUMSTATES + 63) / 64;                                      // uint64 words per row of the C ABI
        static_assert(sizeof(typename Base::Decisions::format_t) == 8, "decision blocks are 64-bit on the hosts this adapter is built for");
        std::vector<uint64_t> rows(steps * W64);
        if (steps) ok(vitb_get_decisions(h, first, steps, rows.data()), "get_decisions");
        for (size_t r = 0; r < steps; r++) {
            auto* dst = base.m_decisions[first + r];
            for (size_t w = 0; w < Base::Decisions::TOTAL_BLOCKS; w++) dst[w] = rows[r * W64 + w];
        }
        base.m_current_decoded_bit = first + steps;                                              // scalar.h:52
        return sum_error_t(acc);
    }

private:
    static void ok(int status, const char* what) {
        if (status != VITB_OK) throw std::runtime_error(std::string("ViterbiDecoder_CUDA_Ref: ") + what + ": " + vitb_status_string(status));
    }
    // one handle per (polynomials, soft levels, config) and host thread; created on first use, destroyed when the thread ends
    struct Cache {
        std::map<std::vector<uint32_t>, vitb_decoder*> handles;
        ~Cache() { for (auto& kv : handles) vitb_destroy(kv.second); }
    };
    static vitb_decoder* handle_for(const Base& base) {
        static thread_local Cache cache;
        vitb_params p{};
        p.K = int32_t(K); p.R = int32_t(R);
        polynomials_of<K, R, soft_t>(base.m_branch_table, p.G, p.soft_decision_high, p.soft_decision_low);
        p.soft_bytes = int32_t(sizeof(soft_t));
        p.soft_decision_max_error = base.m_config.soft_decision_max_error;
        p.initial_start_error = base.m_config.initial_start_error;
        p.initial_non_start_error = base.m_config.initial_non_start_error;
        p.renormalisation_threshold = base.m_config.renormalisation_threshold;
        p.tie_break = VITB_TIE_SCALAR; p.device = 0;
        std::vector<uint32_t> key(p.G, p.G + R);
        for (uint32_t v : {uint32_t(p.soft_decision_high), uint32_t(p.soft_decision_low), p.soft_decision_max_error, p.initial_start_error,
                           p.initial_non_start_error, p.renormalisation_threshold}) key.push_back(v);
        auto it = cache.handles.find(key);
        if (it != cache.handles.end()) return it->second;
        vitb_decoder* h = nullptr;
        ok(vitb_create(&p, &h), "vitb_create");
        cache.handles.emplace(key, h);
        return h;
    }
};

}  // namespace viterbi_cuda
