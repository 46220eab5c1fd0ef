/*
 * viterbi_oracle.h -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Plain-C CPU restatement of the reference hot path
 *     reset -> update (add-compare-select) -> get_error -> chainback
 * of williamyang98/ViterbiDecoderCpp, plus the test-data producers the parity
 * tests need (shift-register convolutional encoder, puncturing).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this port
 * bit-for-bit (decoded bytes, every decision row, sum of renormalisation
 * minima, final metric) against the reference's own headers compiled in place
 * (oracle/ref_harness.cpp -> oracle/_ref/libvitref.so) and against the golden
 * vectors under tests/golden/ that were generated from that build, and
 * reproduces the reference programs' known answers (run_simple: error 0;
 * run_punctured_decoder: 100584 / 2376 / 792).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef VITERBI_ORACLE_H
#define VITERBI_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_MAX_R 16

/* arithmetic flavour of the add-compare-select */
enum { VO_MODE_SCALAR = 0,   /* viterbi_decoder_scalar.h: wrapping error_t, strict '>' select      */
       VO_MODE_SIMD   = 1 }; /* x86/viterbi_decoder_avx_u16.h: saturating, decision = (min==path1) */

typedef struct vo_decoder vo_decoder;

/* err_bits: 8 | 16 | 32 (error_t = uint8_t/uint16_t/uint32_t); soft_bits: 8 | 16 (soft_t = int8_t/int16_t).
 * cfg = { soft_decision_max_error, initial_start_error, initial_non_start_error, renormalisation_threshold } */
vo_decoder* vo_create(int K, int R, const uint32_t* G, int err_bits, int soft_bits, int high, int low, const uint64_t cfg[4], int mode);
void vo_destroy(vo_decoder* d);

void vo_branch_table(const vo_decoder* d, int32_t* out /* [R][2^(K-2)] */);
void vo_set_traceback_length(vo_decoder* d, size_t traceback_length);
size_t vo_get_traceback_length(const vo_decoder* d);
void vo_reset(vo_decoder* d, size_t starting_state);
/* symbols: soft_t array (int8_t or int16_t per soft_bits); n multiple of R. Returns the sum of renormalisation minima. -1 on misuse. */
int64_t vo_update(vo_decoder* d, const void* symbols, size_t n);
uint32_t vo_get_error(const vo_decoder* d, size_t end_state);
int vo_chainback(const vo_decoder* d, uint8_t* bytes_out, size_t total_bits, size_t end_state);
size_t vo_current_decoded_bit(const vo_decoder* d);
/* decision rows in the reference layout: row t = max(2^(K-1)/64,1) uint64 words, bit s%64 of word s/64 */
const uint64_t* vo_decision_row(const vo_decoder* d, size_t t);
size_t vo_decision_words_per_row(const vo_decoder* d);
const uint32_t* vo_metrics(const vo_decoder* d);
uint32_t vo_max_metric_seen(const vo_decoder* d);
uint32_t vo_max_metric_seen_all(vo_decoder* d, int clear);
uint64_t vo_clipped(vo_decoder* d, int clear);

/* whole-frame convenience: reset(0) + update(all) + get_error(0) + chainback(L, 0) over n_frames frames laid out [frame][(L+K-1)*R] */
int vo_decode_frames(vo_decoder* d, const void* symbols, size_t n_frames, size_t L,
                     uint8_t* out_bytes /* [F][ceil(L/8)] */, uint64_t* acc_error /* [F] */, uint32_t* final_error /* [F] */);

/* depuncture + update one decoded bit at a time (examples/helpers/puncture_code_helpers.h:17-55).
 * Returns number of punctured symbols consumed; *acc += sum of renormalisation minima. */
size_t vo_update_punctured(vo_decoder* d, int unpunctured_value, const void* punctured_symbols, size_t total_symbols,
                           const uint8_t* puncture_code, size_t puncture_code_length, size_t requested_output_symbols, uint64_t* acc);

/* test-data producers */
/* shift-register encoder + K-1 zero tail bits; out_bits[(nbytes*8+K-1)*R] in {0,1}
 * (include/viterbi/convolutional_encoder_shift_register.h:45-61, examples/helpers/test_helpers.h:17-64) */
size_t vo_encode(int K, int R, const uint32_t* G, const uint8_t* bytes, size_t nbytes, uint8_t* out_bits);
int vo_parity(uint64_t x);

#ifdef __cplusplus
}
#endif
#endif
