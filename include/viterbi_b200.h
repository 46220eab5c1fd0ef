/*
 * viterbi_b200.h -- C ABI of the B200 (sm_100a) batched Viterbi decoder.
 *
 * Drop-in CUDA backend for the hot path of williamyang98/ViterbiDecoderCpp:
 *     reset -> update (add-compare-select) -> get_error -> chainback
 * The reference is a header-only C++ library with no FFI of its own; each entry point below names the reference interface
 * it replaces (paths relative to the reference tree).  The header-only C++ facade in include/viterbi_cuda/ wraps this ABI
 * in the reference's template shapes; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 (VITB_OK) or a negative vitb_status;
 * host pointers are never retained after return; a handle is single-threaded mutable state (like ViterbiDecoder_Core), make
 * one per host thread / per GPU.  There is no CPU fallback: without a CUDA device every call fails with VITB_ERR_CUDA.
 */
#ifndef VITERBI_B200_H
#define VITERBI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITB_MAX_R 16
#define VITB_END_STATE_BEST ((size_t)-1)

typedef enum vitb_status {
    VITB_OK = 0,
    VITB_ERR_ARG = -1,          /* violates a reference assert (symbol count not multiple of R, too many bits, bad state ...) */
    VITB_ERR_UNSUPPORTED = -2,  /* (K, R, types) combination has no kernel (the analogue of Decoder::is_valid == false)      */
    VITB_ERR_CUDA = -3,         /* CUDA runtime error; vitb_last_cuda_error() has the cudaError_t                             */
    VITB_ERR_STATE = -4,        /* call protocol violated (e.g. chainback before enough update steps)                        */
    VITB_ERR_NOMEM = -5
} vitb_status;

/* tie-break of the compare-select */
#define VITB_TIE_SCALAR 0   /* decision = path0 >  path1  (viterbi_decoder_scalar.h:123-124) -- the parity oracle            */
#define VITB_TIE_SIMD   1   /* decision = path0 >= path1  (x86/viterbi_decoder_avx_u16.h:112-115, SSE/NEON alike)            */
/* VITB_TIE_SIMD keeps the scalar decoder's wrapping arithmetic and only breaks ties the SIMD way: equal to the SIMD decoders as
 * long as no metric saturates there (true for the stock presets except uint8_t metrics of the long codes).  VITB_TIE_SIMD_SAT
 * reproduces them exactly: saturating metric adds (avx_u16.h:107-110 adds_epu16 / avx_u8.h adds_epu8), inverted error formed as
 * max_error -sat total (:106), saturating sum of the branch errors (:95-97), ties - saturated candidates included - to path 1.
 * It runs on the decision-row kernels (the survivor-history kernels keep their decisions where a saturated metric would clobber
 * them), so batch calls are about half as fast as with the other two flavours.  Catalogue codes only. */
#define VITB_TIE_SIMD_SAT 2

/* Everything the reference passes to ViterbiBranchTable<K,R,soft_t>(G, high, low) (viterbi_branch_table.h:33-35) and
 * ViterbiDecoder_Core<K,R,error_t,soft_t>(branch_table, config) (viterbi_decoder_core.h:170), flattened. */
typedef struct vitb_params {
    int32_t K;                          /* constraint length                                                     */
    int32_t R;                          /* code rate 1/R                                                         */
    uint32_t G[VITB_MAX_R];             /* generator polynomials, LSB = newest input bit (viterbi_branch_table.h:31) */
    int32_t soft_bytes;                 /* 2: soft_t=int16_t, error_t=uint16_t;  1: soft_t=int8_t, error_t=uint8_t */
    int32_t soft_decision_high;         /* viterbi_branch_table.h:34                                             */
    int32_t soft_decision_low;          /* viterbi_branch_table.h:35                                             */
    uint32_t soft_decision_max_error;   /* viterbi_decoder_config.h:14                                           */
    uint32_t initial_start_error;       /* viterbi_decoder_config.h:15                                           */
    uint32_t initial_non_start_error;   /* viterbi_decoder_config.h:16                                           */
    uint32_t renormalisation_threshold; /* viterbi_decoder_config.h:17                                           */
    int32_t tie_break;                  /* VITB_TIE_SCALAR (default), VITB_TIE_SIMD or VITB_TIE_SIMD_SAT         */
    int32_t device;                     /* CUDA device ordinal                                                   */
} vitb_params;

typedef struct vitb_decoder vitb_decoder;

/* ---- lifetime: replaces constructing ViterbiBranchTable + ViterbiDecoder_Core (core.h:170-177) ---- */
int vitb_create(const vitb_params* params, vitb_decoder** out);
int vitb_destroy(vitb_decoder* h);
/* 1 if a compiled kernel exists for (K, R, G, soft_bytes): the analogue of `static constexpr bool Decoder::is_valid` (scalar.h:25) */
int vitb_is_supported(const vitb_params* params);

/* ---- single-frame streaming API: same call protocol as the reference (examples/run_simple.cpp:76-80) ---- */
int vitb_set_traceback_length(vitb_decoder* h, size_t traceback_length);        /* core.h:180-186 */
int vitb_get_traceback_length(const vitb_decoder* h, size_t* traceback_length); /* core.h:189-192 */
int vitb_reset(vitb_decoder* h, size_t starting_state);                         /* core.h:202-211 */
/* Decoder::update<uint64_t>(base, symbols, N) (scalar.h:28-55): `symbols` is a HOST array of N soft_t, N % R == 0, may be called
 * repeatedly; *accumulated_error receives the sum of renormalisation minima of THIS call.  One synchronous round trip per call
 * (calls of up to 64 KB of symbols are staged in mapped pinned memory: two kernel launches and one synchronize, ~30 us on a B200;
 * the batch calls below are the fast path). */
int vitb_update(vitb_decoder* h, const void* symbols, size_t n_symbols, uint64_t* accumulated_error);
int vitb_get_error(vitb_decoder* h, size_t end_state, uint32_t* error);         /* core.h:195-199; VITB_END_STATE_BEST: the minimum */
int vitb_chainback(vitb_decoder* h, uint8_t* bytes_out, size_t total_bits, size_t end_state); /* core.h:214-236; or VITB_END_STATE_BEST */
/* public fields of the reference Core, copied out on request */
int vitb_get_current_decoded_bit(const vitb_decoder* h, size_t* bit);           /* m_current_decoded_bit  core.h:242 */
int vitb_get_metrics(vitb_decoder* h, uint32_t* metrics_out /* [2^(K-1)] */);   /* m_metrics.get_old()    core.h:240 */
/* m_decisions[first_row .. first_row+n_rows) in the reference layout: max(2^(K-1)/64,1) uint64 per row, bit s%64 of word s/64 (core.h:49-83) */
int vitb_get_decisions(vitb_decoder* h, size_t first_row, size_t n_rows, uint64_t* rows_out);
/* m_metrics and m_current_decoded_bit are public members of the reference's core (core.h:238-242), so a caller may also ASSIGN
 * them; the adapter that lets the reference's own programs drive this backend (include/viterbi_cuda/viterbi_decoder_cuda_ref.h)
 * keeps the decoder state in the reference object and hands it to the GPU around every update.  metrics_in: [2^(K-1)] values. */
int vitb_set_metrics(vitb_decoder* h, const uint32_t* metrics_in);
int vitb_set_current_decoded_bit(vitb_decoder* h, size_t bit);

/* ---- batched API: n_frames independent frames per call, each decoded as reset(start) + update(all) + get_error(end) +
 *      chainback(L, end).  symbols: [n_frames][row_stride] soft_t with (L+K-1)*R symbols used per frame (or the punctured
 *      count, see vitb_set_puncture_schedule); row_stride = 0 means densely packed.
 *      out_bytes [n_frames][ceil(L/8)], acc_error [n_frames] (sum of renormalisation minima), final_error [n_frames]
 *      (get_error(end_state)); any output pointer may be NULL.  Total path error of a frame = acc_error + final_error
 *      (examples/run_simple.cpp:78-79). ---- */
typedef struct vitb_batch_opts {
    size_t row_stride;       /* elements between frames in `symbols`; 0 = dense */
    size_t starting_state;   /* reset(starting_state)          core.h:202 */
    size_t end_state;        /* get_error / chainback end_state core.h:195,214; VITB_END_STATE_BEST = per frame, the state with
                              * the smallest final path metric (lowest index on a tie) */
} vitb_batch_opts;

/* host pointers: H2D copy, kernels, D2H copy, synchronous */
int vitb_decode_batch(vitb_decoder* h, const void* symbols, size_t n_frames, size_t total_bits, const vitb_batch_opts* opts,
                      uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error);
/* device pointers on the handle's device, enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream), asynchronous.
 * All batch calls of ONE handle share its device workspace, so the library orders them on the device: a call enqueued on another
 * stream first waits (cudaStreamWaitEvent) for the previous batch call of the same handle.  Calls on one handle therefore never
 * overlap each other, whatever streams they use; to overlap two batches (double buffering) use two handles.  The host side of a
 * handle stays single-threaded. */
int vitb_decode_batch_dev(vitb_decoder* h, const void* d_symbols, size_t n_frames, size_t total_bits, const vitb_batch_opts* opts,
                          uint8_t* d_out_bytes, uint64_t* d_acc_error, uint32_t* d_final_error, void* stream);
/* Pipelined batch calls.  Inside the library the add-compare-select kernel of a chunk runs on the caller's stream and its traceback /
 * result gather on a second, handle-owned stream (different bottlenecks: instruction issue vs DRAM round trips), with two workspace
 * slots; by default the caller's stream waits for the traceback before vitb_decode_batch_dev returns control of it, so results are
 * in stream order as documented above.  With pipelining enabled that wait is DEFERRED by one call: when call i+1 returns, the
 * results of call i are complete in the order of call i+1's stream, while the traceback of call i+1 is still free to run next to the
 * ACS of call i+2 (config 2: 0.83 -> ~0.66 ms per batch of 65 536 frames).  vitb_batch_flush(h, stream) makes `stream` wait for
 * everything outstanding.  Output buffers of consecutive calls may be the same (tracebacks run in call order).  Twice the workspace.
 * Pipelined calls also run their ACS kernel on a handle-owned stream (so that the caller's stream can already run the next call's
 * ingest pass: a three-stage pipeline ingest | ACS | traceback), which means the INPUT buffer of call i is still being read after the
 * call returned: like its outputs, it must be left alone until call i+1 has returned (stream order) or the flush. */
int vitb_set_pipelining(vitb_decoder* h, int enabled);
int vitb_batch_flush(vitb_decoder* h, void* stream);

/* pinned host pointers, enqueued on `stream`: H2D copy, kernels and D2H copies are all asynchronous; the caller synchronises the
 * stream before reading the outputs (pageable memory works too but then the copies block). */
int vitb_decode_batch_async(vitb_decoder* h, const void* symbols, size_t n_frames, size_t total_bits, const vitb_batch_opts* opts,
                            uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error, void* stream);

/* Punctured input for the batched API (examples/helpers/puncture_code_helpers.h:17-55, schedule as in
 * examples/run_punctured_decoder.cpp:248-286): keep[i] != 0 means depunctured symbol i was transmitted and is taken from the
 * frame's row in order; keep[i] == 0 inserts `unpunctured_value`.  n_depunctured must equal (L+K-1)*R of later batch calls.
 * n_depunctured = 0 clears the schedule. */
int vitb_set_puncture_schedule(vitb_decoder* h, const uint8_t* keep, size_t n_depunctured, int32_t unpunctured_value);

/* ---- device-side front end (the "next" row after the hot path: the producers of decoder input).  The test-data pipeline of the
 *      reference's BER sweep (examples/run_snr_ber.cpp:311-359) on the GPU: random bytes -> convolutional encoder with K-1 zero tail
 *      bits (include/viterbi/convolutional_encoder_shift_register.h:45-61) -> BPSK + AWGN at EbNo_dB (NaN = noise free) -> the
 *      reference quantiser (round, clamp to [low, high]) -> soft_t rows, punctured if a schedule is set.  Encoder and quantiser are
 *      exact; the random streams are Philox4x32-10, not std::mt19937. ---- */
int vitb_synth_frames_dev(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed,
                          uint8_t* d_tx_bytes /* [F][L/8] out */, void* d_symbols /* [F][row_stride] out */, size_t row_stride, void* stream);
/* same, results copied to host arrays (tests, small runs) */
int vitb_synth_frames(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed,
                      uint8_t* tx_bytes, void* symbols, size_t row_stride);
/* the quantiser alone on host arrays: soft = clamp(round(x * scale + mean), low, high)   (run_snr_ber.cpp:352-359) */
int vitb_quantise(vitb_decoder* h, const float* x, size_t n, float scale, float mean, void* soft_out);
/* generate n_frames frames at EbNo_dB, decode them and count bit errors, all on the device (run_snr_ber.cpp:336-372 for one batch) */
int vitb_ber_trial(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed, uint64_t* bit_errors);

/* ---- multi-GPU: frames are independent, so the batch is split into contiguous ranges, one per handle (each created on its own
 *      device), decoded concurrently from host memory; no collective.  With pinned host buffers (cudaHostAlloc / cudaHostRegister)
 *      one host thread enqueues every device's copies and kernels on that device's streams and then waits for all of them; with
 *      pageable buffers (whose asynchronous copies block) one host thread per device is used.  Synchronous.  ---- */
int vitb_decode_batch_multi(vitb_decoder* const* handles, int n_handles, const void* symbols, size_t n_frames, size_t total_bits,
                            const vitb_batch_opts* opts, uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error);

/* ---- introspection ---- */
/* bytes of device workspace a batch call of this shape needs (decision rows dominate: n_frames * (L+K-1) * 2^(K-1)/8) */
int vitb_workspace_bytes(const vitb_decoder* h, size_t n_frames, size_t total_bits, size_t* bytes);
/* cap on the workspace; larger batches are processed in chunks of frames.  0 = default (VITB_WORKSPACE_MB, else 60 % of device memory) */
int vitb_set_workspace_limit(vitb_decoder* h, size_t bytes);
/* number of CUDA kernels this handle has launched so far (bench.py reports it as gpu_launches) */
int vitb_kernel_launch_count(const vitb_decoder* h, uint64_t* count);
/* stage profiling: when enabled every batch chunk records CUDA events around its stages on the launching stream.
 * vitb_get_stage_ms (call after synchronising the stream) returns the summed device time of the LAST batch call per stage:
 * ms[0] ingest, ms[1] add-compare-select, ms[2] traceback, ms[3] result gather. */
int vitb_set_profiling(vitb_decoder* h, int enabled);
int vitb_get_stage_ms(vitb_decoder* h, float ms[4]);
/* Kernel variants: a frame pair can be spread over 1, 2, 4 ... lanes (more lanes = more warps for small batches).  By default the
 * variant is chosen per batch call from the batch size; vitb_set_variant pins it (0 = automatic).  vitb_get_variants lists the
 * compiled lanes-per-pair settings of this handle's code and returns their number.  Numbers above 100 name batch-only
 * survivor-history kernels that share a lane count with a decision-row kernel: 104 = K = 7, uint16_t metrics, one FRAME over 4
 * lanes (the small-batch kernel); 256 = K = 15, uint16_t metrics, one frame per 512-thread CTA.  A pinned batch-only variant
 * applies when its input requirements hold (unpunctured rows, 4-byte aligned for the lane-group kernels); otherwise the call falls
 * back to the automatic choice. */
int vitb_set_variant(vitb_decoder* h, int lanes_per_pair);
int vitb_get_variants(const vitb_decoder* h, int* lanes_per_pair, int capacity);
/* One-lane-per-pair variants (K <= 7) decode whole-frame batches with the survivor-history kernel (csrc/acs_hist.cuh: decisions ride
 * below the path metrics, traceback moves 8 / 16 steps per lookup) - same results, about twice the speed.  enabled = 0 keeps the
 * decision-row kernels for batch calls too (the streaming calls always use them); default 1, or 0 when VITB_NO_HIST is set. */
int vitb_set_history_kernel(vitb_decoder* h, int enabled);
/* Traceback of the survivor-history records walks every frame in concurrent segments: each segment warms up over overlap_records
 * records from a guessed state (survivor paths merge) and is verified - and re-walked if it did not merge - against the segment
 * above, so the result is exact.  seg_records = 0 / overlap_records = -1 choose automatically (by batch size / 96 steps); tests use
 * overlap 0 to force the repair path.  Units: history records (8 or 16 steps) for K <= 7, 8 decoded bits / 8 decision rows for the
 * K = 15 row walk; the lane-group kernels (K = 9) stream whole rows and are not segmented. */
int vitb_set_traceback_segments(vitb_decoder* h, int seg_records, int overlap_records);
/* Sliding-window traceback (K = 15, whose decision rows are 2 KB per frame and step: 34.4 GB for 1024 frames of 16384 bits).  The
 * reference keeps every row of a frame until chainback (core.h:180-186); with a window the library keeps a ring of 2 x window_bits
 * rows per frame: the ACS kernel runs one window per launch and the decoded bits of a window are walked out as soon as the next
 * window is in, starting window_bits - (K-1) rows above them from state 0 (survivor paths merge within a few constraint lengths;
 * the last window starts from the true end state).  Path metrics and errors are not affected.  The decoded bytes equal the exact
 * result whenever every warm-up merged, and the library COUNTS the window boundaries where it did not:
 * vitb_get_window_mismatches (of the last batch call; synchronises the device) == 0 certifies the bytes.  window_bits is rounded
 * up to a multiple of 40; 0 = keep every row (exact, verified and repaired: the default).  vitb_workspace_bytes reflects the
 * setting (config 5: 34.4 GB -> 1.7 GB at window_bits = 400).  VITB_ERR_UNSUPPORTED for K <= 9 (rows are small there). */
int vitb_set_traceback_window(vitb_decoder* h, size_t window_bits);
int vitb_get_window_mismatches(vitb_decoder* h, uint64_t* count);
/* name of the ACS kernel variant selected for this handle, e.g. "acs_pair<K7,R2,u8,scalar-tie>" */
const char* vitb_kernel_name(const vitb_decoder* h);
int vitb_last_cuda_error(const vitb_decoder* h);
const char* vitb_status_string(int status);
/* library-level: number of CUDA devices visible (0 when no driver/GPU: every create then fails with VITB_ERR_CUDA) */
int vitb_device_count(void);
const char* vitb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* VITERBI_B200_H */
