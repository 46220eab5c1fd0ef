// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" shim around the UNMODIFIED reference headers, which are included where they lie
// (-I/root/reference/include -I/root/reference/examples); no reference source is copied into this repo.
// Built by oracle/Makefile into oracle/_ref/libvitref.so (git-ignored, shipped to the GPU box by gpurun).
//
// Used (a) to pin oracle/viterbi_oracle.c bit-for-bit, (b) to generate tests/golden/*.npz, and (c) as the
// CPU baseline (reference AVX2 / SSE / scalar decoders on pinned host threads) that bench.py reports.
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <vector>
#include <thread>
#include <chrono>
#include <atomic>
#include <pthread.h>
#include <sched.h>

#include "viterbi/viterbi_branch_table.h"
#include "viterbi/viterbi_decoder_config.h"
#include "viterbi/viterbi_decoder_core.h"
#include "viterbi/viterbi_decoder_scalar.h"
#include "viterbi/x86/viterbi_decoder_sse_u16.h"
#include "viterbi/x86/viterbi_decoder_sse_u8.h"
#include "viterbi/x86/viterbi_decoder_avx_u16.h"
#include "viterbi/x86/viterbi_decoder_avx_u8.h"
#include "viterbi/convolutional_encoder_shift_register.h"
#include "helpers/puncture_code_helpers.h"
#include "helpers/test_helpers.h"

namespace {

enum { IMPL_SCALAR = 0, IMPL_SSE = 1, IMPL_AVX = 2 };

template <size_t K, size_t R, int SOFT_BYTES> struct Types;
template <size_t K, size_t R> struct Types<K, R, 2> {
    using soft_t = int16_t; using error_t = uint16_t;
    using Scalar = ViterbiDecoder_Scalar<K, R, uint16_t, int16_t>;
    using SSE = ViterbiDecoder_SSE_u16<K, R>;
    using AVX = ViterbiDecoder_AVX_u16<K, R>;
};
template <size_t K, size_t R> struct Types<K, R, 1> {
    using soft_t = int8_t; using error_t = uint8_t;
    using Scalar = ViterbiDecoder_Scalar<K, R, uint8_t, int8_t>;
    using SSE = ViterbiDecoder_SSE_u8<K, R>;
    using AVX = ViterbiDecoder_AVX_u8<K, R>;
};

struct Args {
    const uint32_t* G; int high, low; const uint64_t* cfg;
    const void* symbols; size_t n_frames, L;
    uint8_t* out_bytes; uint64_t* acc; uint32_t* final_error;
    uint64_t* decisions;   // optional: [F][L+K-1][words] reference layout
    uint32_t* metrics;     // optional: [F][2^(K-1)] final metrics
    int n_threads; double* seconds;
    int passes = 1;        // timing runs: every thread decodes its range `passes` times (threads persist; seconds = mean per pass)
};

template <class Decoder, size_t K, size_t R, typename error_t, typename soft_t>
int decode_range(const ViterbiBranchTable<K, R, soft_t>& table, const ViterbiDecoder_Config<error_t>& config,
                 const Args& a, size_t f0, size_t f1) {
    using Core = ViterbiDecoder_Core<K, R, error_t, soft_t>;
    if constexpr (!Decoder::is_valid) {
        return -3;
    } else {
        Core core(table, config);
        const size_t steps = a.L + (K - 1), n_sym = steps * R, out_stride = (a.L + 7) / 8;
        const size_t words = Core::Decisions::TOTAL_BLOCKS;
        core.set_traceback_length(a.L);
        for (size_t f = f0; f < f1; f++) {
            const soft_t* sym = static_cast<const soft_t*>(a.symbols) + f * n_sym;
            core.reset();
            const uint64_t acc = Decoder::template update<uint64_t>(core, sym, n_sym);
            if (a.acc) a.acc[f] = acc;
            if (a.final_error) a.final_error[f] = uint32_t(core.get_error());
            if (a.out_bytes) core.chainback(a.out_bytes + f * out_stride, a.L, 0u);
            if (a.decisions) {
                for (size_t t = 0; t < steps; t++)
                    memcpy(a.decisions + (f * steps + t) * words, core.m_decisions[t], words * sizeof(uint64_t));
            }
            if (a.metrics) {
                for (size_t s = 0; s < Core::NUMSTATES; s++) a.metrics[f * Core::NUMSTATES + s] = uint32_t(core.get_error(s));
            }
        }
        return 0;
    }
}

void pin_to_cpu(int cpu) {
    cpu_set_t set; CPU_ZERO(&set); CPU_SET(cpu, &set);
    pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

template <class Decoder, size_t K, size_t R, typename error_t, typename soft_t>
int decode_all(const Args& a) {
    using Table = ViterbiBranchTable<K, R, soft_t>;
    std::vector<uint32_t> g(a.G, a.G + R);
    Table table(g.data(), soft_t(a.high), soft_t(a.low));
    ViterbiDecoder_Config<error_t> config;
    config.soft_decision_max_error = error_t(a.cfg[0]);
    config.initial_start_error = error_t(a.cfg[1]);
    config.initial_non_start_error = error_t(a.cfg[2]);
    config.renormalisation_threshold = error_t(a.cfg[3]);
    const int nt = a.n_threads > 0 ? a.n_threads : 1;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = 0;
    if (nt == 1) {
        for (int pass = 0; pass < a.passes && !rc; pass++) rc = decode_range<Decoder, K, R, error_t, soft_t>(table, config, a, 0, a.n_frames);
    } else {
        // one ViterbiDecoder_Core per thread, contiguous frame ranges, each thread pinned to one allowed logical CPU
        cpu_set_t allowed; CPU_ZERO(&allowed);
        sched_getaffinity(0, sizeof(allowed), &allowed);
        std::vector<int> cpus;
        for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
        std::vector<std::thread> threads;
        std::atomic<int> worst{0};
        for (int t = 0; t < nt; t++) {
            const size_t f0 = a.n_frames * size_t(t) / size_t(nt), f1 = a.n_frames * size_t(t + 1) / size_t(nt);
            threads.emplace_back([&, t, f0, f1]() {
                if (!cpus.empty()) pin_to_cpu(cpus[size_t(t) % cpus.size()]);
                for (int pass = 0; pass < a.passes; pass++) {
                    const int r = decode_range<Decoder, K, R, error_t, soft_t>(table, config, a, f0, f1);
                    if (r) { worst = r; break; }
                }
            });
        }
        for (auto& th : threads) th.join();
        rc = worst;
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (a.seconds) *a.seconds = std::chrono::duration<double>(t1 - t0).count() / double(a.passes > 0 ? a.passes : 1);
    return rc;
}

template <size_t K, size_t R, int SOFT_BYTES>
int dispatch_impl(int impl, const Args& a) {
    using T = Types<K, R, SOFT_BYTES>;
    using soft_t = typename T::soft_t; using error_t = typename T::error_t;
    switch (impl) {
    case IMPL_SCALAR: return decode_all<typename T::Scalar, K, R, error_t, soft_t>(a);
    case IMPL_SSE:    return decode_all<typename T::SSE, K, R, error_t, soft_t>(a);
    case IMPL_AVX:
        if (!__builtin_cpu_supports("avx2")) return -4;
        return decode_all<typename T::AVX, K, R, error_t, soft_t>(a);
    default: return -2;
    }
}

#define FOR_KR(K_, R_) if (K == K_ && R == R_) return soft_bytes == 2 ? dispatch_impl<K_, R_, 2>(impl, a) : dispatch_impl<K_, R_, 1>(impl, a);

int dispatch(int K, int R, int soft_bytes, int impl, const Args& a) {
    if (soft_bytes != 1 && soft_bytes != 2) return -2;
    FOR_KR(3, 2) FOR_KR(5, 2) FOR_KR(7, 2) FOR_KR(7, 3) FOR_KR(7, 4) FOR_KR(9, 2) FOR_KR(9, 4) FOR_KR(15, 6)
    return -1;   // (K,R) pair not instantiated
}

}  // namespace

extern "C" {

// Decode n_frames frames of (L+K-1)*R soft symbols each with the reference decoder `impl` (0 scalar, 1 SSE, 2 AVX2).
// soft_bytes 2 -> <uint16_t,int16_t>, 1 -> <uint8_t,int8_t>. Returns 0, or <0 (unsupported combination).
int vitref_decode(int K, int R, const uint32_t* G, int soft_bytes, int high, int low, const uint64_t cfg[4], int impl,
                  const void* symbols, size_t n_frames, size_t L,
                  uint8_t* out_bytes, uint64_t* acc, uint32_t* final_error, uint64_t* decisions, uint32_t* metrics,
                  int n_threads, double* seconds) {
    Args a{G, high, low, cfg, symbols, n_frames, L, out_bytes, acc, final_error, decisions, metrics, n_threads, seconds};
    return dispatch(K, R, soft_bytes, impl, a);
}

// same, `passes` passes over the batch with the worker threads (one pinned Core each) kept alive in between; *seconds = mean per pass
int vitref_decode_passes(int K, int R, const uint32_t* G, int soft_bytes, int high, int low, const uint64_t cfg[4], int impl,
                         const void* symbols, size_t n_frames, size_t L, uint8_t* out_bytes, uint64_t* acc, uint32_t* final_error,
                         int n_threads, int passes, double* seconds) {
    Args a{G, high, low, cfg, symbols, n_frames, L, out_bytes, acc, final_error, nullptr, nullptr, n_threads, seconds};
    a.passes = passes > 0 ? passes : 1;
    return dispatch(K, R, soft_bytes, impl, a);
}

// ConvolutionalEncoder_ShiftRegister + encode_data (examples/helpers/test_helpers.h:17-64): bits -> {high, low} int16 symbols
size_t vitref_encode(int K, int R, const uint32_t* G, const uint8_t* bytes, size_t nbytes, int16_t* out, int16_t high, int16_t low) {
    ConvolutionalEncoder_ShiftRegister<uint32_t> enc(size_t(K), size_t(R), G);
    enc.reset();
    const size_t max_symbols = (nbytes * 8 + size_t(K - 1)) * size_t(R);
    return encode_data<int16_t>(&enc, bytes, nbytes, out, max_symbols, high, low);
}

// DAB fast-information-channel punctured round trip, following examples/run_punctured_decoder.cpp:196-286 with the
// puncture tables passed in by the caller (pi16/pi15: 32 entries, pix: 24 entries).
// Encodes 96 bytes -> 2304 punctured symbols.
size_t vitref_fic_encode(const uint32_t* G, const uint8_t* bytes, int16_t* out, size_t max_out, int16_t high, int16_t low,
                         const bool* pi16, const bool* pi15, const bool* pix) {
    constexpr size_t K = 7, R = 4;
    ConvolutionalEncoder_ShiftRegister<uint32_t> enc(K, R, G);
    enc.reset();
    size_t n = 0, ib = 0;
    for (int i = 0; i < 21; i++, ib += 4) n += encode_punctured_data<int16_t>(&enc, bytes + ib, 4, out + n, max_out - n, pi16, 32, high, low);
    for (int i = 0; i < 3; i++, ib += 4)  n += encode_punctured_data<int16_t>(&enc, bytes + ib, 4, out + n, max_out - n, pi15, 32, high, low);
    n += encode_punctured_tail<int16_t>(&enc, out + n, max_out - n, pix, 24, high, low);
    return n;
}

// decode_punctured_symbols x3 + chainback for one FIC frame with the scalar (impl 0) / SSE / AVX reference decoder.
int vitref_fic_decode(const uint32_t* G, int soft_bytes, int high, int low, const uint64_t cfg[4], int impl,
                      const void* symbols, size_t total_symbols, const bool* pi16, const bool* pi15, const bool* pix,
                      uint8_t* out_bytes, uint64_t* acc_out, uint32_t* final_error) {
    constexpr size_t K = 7, R = 4, L = 768;
    auto run = [&](auto tag_soft, auto tag_err, auto tag_dec) -> int {
        using soft_t = decltype(tag_soft); using error_t = decltype(tag_err); using Dec = decltype(tag_dec);
        std::vector<uint32_t> g(G, G + R);
        ViterbiBranchTable<K, R, soft_t> table(g.data(), soft_t(high), soft_t(low));
        ViterbiDecoder_Config<error_t> config{error_t(cfg[0]), error_t(cfg[1]), error_t(cfg[2]), error_t(cfg[3])};
        ViterbiDecoder_Core<K, R, error_t, soft_t> core(table, config);
        core.set_traceback_length(L);
        core.reset();
        const soft_t* p = static_cast<const soft_t*>(symbols);
        size_t remain = total_symbols;
        uint64_t acc = 0;
        auto r1 = decode_punctured_symbols<Dec>(core, soft_t(0), p, remain, pi16, 32, 32 * R * 21);
        acc += r1.accumulated_error; p += r1.index_punctured_symbol; remain -= r1.index_punctured_symbol;
        auto r2 = decode_punctured_symbols<Dec>(core, soft_t(0), p, remain, pi15, 32, 32 * R * 3);
        acc += r2.accumulated_error; p += r2.index_punctured_symbol; remain -= r2.index_punctured_symbol;
        auto r3 = decode_punctured_symbols<Dec>(core, soft_t(0), p, remain, pix, 24, 24);
        acc += r3.accumulated_error; remain -= r3.index_punctured_symbol;
        core.chainback(out_bytes, L, 0u);
        *acc_out = acc;
        *final_error = uint32_t(core.get_error());
        return remain == 0 ? 0 : -5;
    };
    if (soft_bytes == 2) {
        if (impl == IMPL_SCALAR) return run(int16_t(0), uint16_t(0), ViterbiDecoder_Scalar<K, R, uint16_t, int16_t>());
        if (impl == IMPL_SSE) return run(int16_t(0), uint16_t(0), ViterbiDecoder_SSE_u16<K, R>());
        if (impl == IMPL_AVX) return run(int16_t(0), uint16_t(0), ViterbiDecoder_AVX_u16<K, R>());
    } else if (soft_bytes == 1) {
        if (impl == IMPL_SCALAR) return run(int8_t(0), uint8_t(0), ViterbiDecoder_Scalar<K, R, uint8_t, int8_t>());
        if (impl == IMPL_SSE) return run(int8_t(0), uint8_t(0), ViterbiDecoder_SSE_u8<K, R>());
        if (impl == IMPL_AVX) return run(int8_t(0), uint8_t(0), ViterbiDecoder_AVX_u8<K, R>());
    }
    return -2;
}

int vitref_host_threads() {
    cpu_set_t allowed; CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof(allowed), &allowed);
    return CPU_COUNT(&allowed);
}

}  // extern "C"
