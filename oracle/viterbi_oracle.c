/*
 * viterbi_oracle.c -- TEST INFRASTRUCTURE ONLY (see viterbi_oracle.h).
 *
 * CPU restatement, in plain C, of the reference decoder's arithmetic.  Every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference).  Nothing here is used by the CUDA product path.
 */
#include "viterbi_oracle.h"
#include <stdlib.h>
#include <string.h>

struct vo_decoder {
    int K, R, err_bits, soft_bits, high, low, mode;
    uint32_t G[VO_MAX_R];
    uint32_t max_error, start_error, non_start_error, renorm_threshold; /* viterbi_decoder_config.h:14-17 */
    size_t n_states, n_half, words_per_row;
    int32_t* branch;            /* [R][n_half], values high/low      viterbi_branch_table.h:45-54 */
    uint32_t* metric[2];        /* double buffer                      viterbi_decoder_core.h:22-45 */
    int cur;                    /* index of the "old" buffer                                        */
    uint64_t* decisions;        /* [rows][words_per_row]              viterbi_decoder_core.h:49-83 */
    size_t rows;                /* traceback_length + K-1             viterbi_decoder_core.h:180-186 */
    size_t decoded_bit;         /* m_current_decoded_bit                                            */
    uint32_t max_seen;
    uint64_t n_clipped;      /* see addu / subu */
    uint32_t max_seen_all;   /* like max_seen, but kept across reset(): the largest metric of a whole batch */
};

/* include/viterbi/parity_table.h:46-54 : parity of all bits of x */
int vo_parity(uint64_t x) {
    x ^= x >> 32; x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
    return (int)(x & 1u);
}

static uint32_t err_mask(const vo_decoder* d) { return d->err_bits >= 32 ? 0xFFFFFFFFu : ((1u << d->err_bits) - 1u); }

/* value -> soft_t (two's complement truncation of an int expression assigned to int8_t/int16_t) */
static int32_t to_soft(const vo_decoder* d, int32_t v) {
    if (d->soft_bits == 8) return (int32_t)(int8_t)(uint8_t)(uint32_t)v;
    return (int32_t)(int16_t)(uint16_t)(uint32_t)v;
}

/* signed saturation to soft_t (_mm256_subs_epi16 / _mm256_subs_epi8) */
static int32_t sat_soft(const vo_decoder* d, int32_t v) {
    const int32_t hi = (d->soft_bits == 8) ? 127 : 32767, lo = -hi - 1;
    return v > hi ? hi : (v < lo ? lo : v);
}

/* n_clipped counts the additions / subtractions whose exact result left the range of error_t: where the SIMD mode saturated, or the
 * scalar mode wrapped (test instrumentation: 0 certifies that wrapping and saturating arithmetic cannot have differed) */
static uint32_t addu(const vo_decoder* d, uint32_t a, uint32_t b) {
    const uint32_t m = err_mask(d);
    const uint64_t s = (uint64_t)a + b;
    if (s > m) ((vo_decoder*)d)->n_clipped++;
    if (d->mode == VO_MODE_SIMD) return s > m ? m : (uint32_t)s;   /* _mm256_adds_epu16: x86/viterbi_decoder_avx_u16.h:107-110 */
    return (a + b) & m;                      /* plain error_t '+': viterbi_decoder_scalar.h:113-116 */
}

static uint32_t subu(const vo_decoder* d, uint32_t a, uint32_t b) {
    if (b > a) ((vo_decoder*)d)->n_clipped++;
    if (d->mode == VO_MODE_SIMD) return a > b ? a - b : 0u;   /* _mm256_subs_epu16: avx_u16.h:106,165 */
    return (a - b) & err_mask(d);                             /* scalar.h:107,149 */
}

vo_decoder* vo_create(int K, int R, const uint32_t* G, int err_bits, int soft_bits, int high, int low, const uint64_t cfg[4], int mode) {
    if (K < 2 || K > 24 || R < 1 || R > VO_MAX_R) return NULL;
    if (!(err_bits == 8 || err_bits == 16 || err_bits == 32) || !(soft_bits == 8 || soft_bits == 16)) return NULL;
    vo_decoder* d = (vo_decoder*)calloc(1, sizeof(*d));
    if (!d) return NULL;
    d->K = K; d->R = R; d->err_bits = err_bits; d->soft_bits = soft_bits; d->high = high; d->low = low; d->mode = mode;
    for (int i = 0; i < R; i++) d->G[i] = G[i];
    const uint32_t m = err_mask(d);
    d->max_error = (uint32_t)cfg[0] & m; d->start_error = (uint32_t)cfg[1] & m;
    d->non_start_error = (uint32_t)cfg[2] & m; d->renorm_threshold = (uint32_t)cfg[3] & m;
    d->n_states = (size_t)1 << (K - 1);
    d->n_half = d->n_states / 2;
    d->words_per_row = d->n_states / 64 ? d->n_states / 64 : 1;   /* core.h:64 */
    d->branch = (int32_t*)malloc(sizeof(int32_t) * (size_t)R * (d->n_half ? d->n_half : 1));
    d->metric[0] = (uint32_t*)malloc(sizeof(uint32_t) * d->n_states);
    d->metric[1] = (uint32_t*)malloc(sizeof(uint32_t) * d->n_states);
    /* viterbi_branch_table.h:45-54 : y = P{(0|X|0) & G[i]} -> high or low */
    for (size_t st = 0; st < d->n_half; st++)
        for (int i = 0; i < R; i++)
            d->branch[(size_t)i * d->n_half + st] = vo_parity((uint64_t)(st << 1) & G[i]) ? high : low;
    vo_reset(d, 0);                       /* core.h:175-176 */
    vo_set_traceback_length(d, 0);
    return d;
}

void vo_destroy(vo_decoder* d) {
    if (!d) return;
    free(d->branch); free(d->metric[0]); free(d->metric[1]); free(d->decisions); free(d);
}

void vo_branch_table(const vo_decoder* d, int32_t* out) {
    memcpy(out, d->branch, sizeof(int32_t) * (size_t)d->R * d->n_half);
}

/* core.h:180-186 */
void vo_set_traceback_length(vo_decoder* d, size_t traceback_length) {
    const size_t rows = traceback_length + (size_t)(d->K - 1);
    d->decisions = (uint64_t*)realloc(d->decisions, sizeof(uint64_t) * rows * d->words_per_row);
    d->rows = rows;
    if (d->decoded_bit > rows) d->decoded_bit = rows;
}

/* core.h:189-192 */
size_t vo_get_traceback_length(const vo_decoder* d) { return d->rows - (size_t)(d->K - 1); }

/* core.h:202-211 */
void vo_reset(vo_decoder* d, size_t starting_state) {
    d->decoded_bit = 0;
    d->max_seen = 0;
    uint32_t* old = d->metric[d->cur];
    for (size_t s = 0; s < d->n_states; s++) old[s] = d->non_start_error;
    old[starting_state & (d->n_states - 1)] = d->start_error;
}

/* One trellis step: viterbi_decoder_scalar.h:58-136 (mode SCALAR) or x86/viterbi_decoder_avx_u16.h:73-136 (mode SIMD) */
static void step(vo_decoder* d, const int32_t* sym, uint64_t* row, const uint32_t* old, uint32_t* nw) {
    const uint32_t m = err_mask(d);
    const size_t H = d->n_half;
    for (size_t w = 0; w < d->words_per_row; w++) row[w] = 0;             /* scalar.h:60-62 */
    for (size_t j = 0; j < H; j++) {
        uint32_t total = 0;
        for (int i = 0; i < d->R; i++) {
            const int32_t expected = d->branch[(size_t)i * H + j];
            uint32_t abs_err;
            if (d->mode == VO_MODE_SIMD) {                                /* avx_u16.h:95-97 */
                int32_t e = sat_soft(d, expected - sym[i]);
                e = to_soft(d, e < 0 ? -e : e);                           /* abs_epi16(INT16_MIN) stays INT16_MIN */
                abs_err = (uint32_t)e & ((d->soft_bits == 8) ? 0xFFu : 0xFFFFu);
                total = addu(d, total, abs_err & m);
            } else {                                                      /* scalar.h:66-73,155-158 */
                const int32_t e = to_soft(d, expected - sym[i]);          /* const soft_t error = expected_sym - sym */
                const int32_t a = to_soft(d, (e > 0) ? e : -e);           /* get_abs<soft_t> returns soft_t           */
                abs_err = (uint32_t)a & m;                                /* error_t(...)                            */
                total = (total + abs_err) & m;
            }
        }
        const uint32_t inv = subu(d, d->max_error, total);                /* scalar.h:107 / avx_u16.h:106 */
        const uint32_t e00 = addu(d, old[j], total);                      /* (0|X) -> (X|0) */
        const uint32_t e10 = addu(d, old[j + H], inv);                    /* (1|X) -> (X|0) */
        const uint32_t e01 = addu(d, old[j], inv);                        /* (0|X) -> (X|1) */
        const uint32_t e11 = addu(d, old[j + H], total);                  /* (1|X) -> (X|1) */
        uint32_t d0, d1;
        if (d->mode == VO_MODE_SIMD) {                                    /* avx_u16.h:112-115: min, then cmpeq(min, path1) */
            d0 = ((e00 < e10 ? e00 : e10) == e10);
            d1 = ((e01 < e11 ? e01 : e11) == e11);
        } else {                                                          /* scalar.h:123-124: strict '>' */
            d0 = e00 > e10;
            d1 = e01 > e11;
        }
        nw[2 * j]     = d0 ? e10 : e00;                                   /* scalar.h:127-128 */
        nw[2 * j + 1] = d1 ? e11 : e01;
        if (nw[2 * j] > d->max_seen) d->max_seen = nw[2 * j];
        if (nw[2 * j + 1] > d->max_seen) d->max_seen = nw[2 * j + 1];
        if (d->max_seen > d->max_seen_all) d->max_seen_all = d->max_seen;
        const size_t s0 = 2 * j;                                          /* scalar.h:131-134 */
        row[s0 / 64] |= ((uint64_t)(d0 | (d1 << 1))) << (s0 % 64);
    }
}

/* scalar.h:139-153 / avx_u16.h:138-170 */
static uint32_t renormalise(vo_decoder* d, uint32_t* metric) {
    uint32_t mn = metric[0];
    for (size_t s = 1; s < d->n_states; s++) if (metric[s] < mn) mn = metric[s];
    for (size_t s = 0; s < d->n_states; s++) metric[s] = subu(d, metric[s], mn);
    return mn;
}

static int32_t load_soft(const vo_decoder* d, const void* symbols, size_t i) {
    return d->soft_bits == 8 ? (int32_t)((const int8_t*)symbols)[i] : (int32_t)((const int16_t*)symbols)[i];
}

/* scalar.h:28-55 */
int64_t vo_update(vo_decoder* d, const void* symbols, size_t n) {
    if (n % (size_t)d->R) return -1;                                       /* scalar.h:37 */
    const size_t steps = n / (size_t)d->R;
    if (steps + d->decoded_bit > d->rows) return -1;                       /* scalar.h:38-40 */
    uint64_t acc = 0;
    int32_t sym[VO_MAX_R];
    for (size_t t = 0; t < steps; t++) {
        for (int i = 0; i < d->R; i++) sym[i] = load_soft(d, symbols, t * (size_t)d->R + (size_t)i);
        uint32_t* old = d->metric[d->cur];
        uint32_t* nw = d->metric[1 - d->cur];
        step(d, sym, d->decisions + d->decoded_bit * d->words_per_row, old, nw);
        if (nw[0] >= d->renorm_threshold) acc += renormalise(d, nw);      /* scalar.h:48-50 : state 0 only */
        d->cur = 1 - d->cur;                                               /* scalar.h:51 */
        d->decoded_bit++;                                                  /* scalar.h:52 */
    }
    return (int64_t)acc;
}

/* core.h:195-199 */
uint32_t vo_get_error(const vo_decoder* d, size_t end_state) { return d->metric[d->cur][end_state & (d->n_states - 1)]; }

/* core.h:214-236 with ViterbiTracebackBuffer core.h:87-153.
 * The buffer is (K-1 state bits | shift_state padding); a decoded bit enters at the top, the state is the top K-1 bits,
 * the emitted byte is the top 8 bits.  bytes_out[j/8] is rewritten on every bit, so byte b ends up holding decoded bits
 * 8b..8b+7 MSB-first; if total_bits % 8 != 0 the last byte's low bits are the leading bits of end_state (then zeros). */
int vo_chainback(const vo_decoder* d, uint8_t* bytes_out, size_t total_bits, size_t end_state) {
    const size_t sb = (size_t)(d->K - 1);
    if (vo_get_traceback_length(d) < total_bits) return -1;               /* core.h:216 */
    if (d->decoded_bit < sb + total_bits) return -1;                      /* core.h:217 */
    if (end_state >= d->n_states) return -1;                              /* core.h:218 */
    const size_t ignore = sb < 8 ? sb : 8;                                /* core.h:136-140 */
    const size_t shift_state = 8 - ignore, shift_tail = sb - ignore, width = sb + shift_state;
    uint64_t buf = (uint64_t)end_state << shift_state;                    /* set_state core.h:101-103 */
    for (size_t i = 0; i < total_bits; i++) {
        const size_t j = (total_bits - 1) - i;
        const uint64_t* row = d->decisions + (j + sb) * d->words_per_row; /* core.h:226-227 */
        const size_t state = (size_t)(buf >> shift_state);                /* get_state core.h:96-98 */
        const uint64_t bit = (row[state / 64] >> (state % 64)) & 1u;      /* core.h:229-232 */
        buf = (buf >> 1) | (bit << (width - 1));                          /* push_bit_in core.h:109-113 */
        bytes_out[j / 8] = (uint8_t)((buf >> shift_tail) & 0xFFu);        /* get_data core.h:105-107, core.h:234 */
    }
    return 0;
}

size_t vo_current_decoded_bit(const vo_decoder* d) { return d->decoded_bit; }
const uint64_t* vo_decision_row(const vo_decoder* d, size_t t) { return d->decisions + t * d->words_per_row; }
size_t vo_decision_words_per_row(const vo_decoder* d) { return d->words_per_row; }
const uint32_t* vo_metrics(const vo_decoder* d) { return d->metric[d->cur]; }
uint32_t vo_max_metric_seen(const vo_decoder* d) { return d->max_seen; }
/* number of additions / subtractions that saturated (SIMD mode) or wrapped (scalar mode) since creation / the last clearing call */
uint64_t vo_clipped(vo_decoder* d, int clear) { const uint64_t v = d->n_clipped; if (clear) d->n_clipped = 0; return v; }
/* largest metric since the decoder was created or the value was last cleared (clear != 0) */
uint32_t vo_max_metric_seen_all(vo_decoder* d, int clear) { const uint32_t v = d->max_seen_all; if (clear) d->max_seen_all = 0; return v; }

/* The call protocol of examples/run_simple.cpp:76-80 applied to each frame of a batch */
int vo_decode_frames(vo_decoder* d, const void* symbols, size_t n_frames, size_t L,
                     uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error) {
    const size_t steps = L + (size_t)(d->K - 1), per_frame = steps * (size_t)d->R, out_stride = (L + 7) / 8;
    const size_t soft_bytes = (size_t)d->soft_bits / 8;
    vo_set_traceback_length(d, L);
    for (size_t f = 0; f < n_frames; f++) {
        vo_reset(d, 0);
        const int64_t acc = vo_update(d, (const uint8_t*)symbols + f * per_frame * soft_bytes, per_frame);
        if (acc < 0) return -1;
        if (acc_error) acc_error[f] = (uint64_t)acc;
        if (final_error) final_error[f] = vo_get_error(d, 0);
        if (out_bytes && vo_chainback(d, out_bytes + f * out_stride, L, 0)) return -1;
    }
    return 0;
}

/* examples/helpers/puncture_code_helpers.h:17-55 */
size_t vo_update_punctured(vo_decoder* d, int unpunctured_value, const void* punctured_symbols, size_t total_symbols,
                           const uint8_t* puncture_code, size_t puncture_code_length, size_t requested_output_symbols, uint64_t* acc) {
    size_t i_code = 0, i_out = 0, i_in = 0;
    int8_t s8[VO_MAX_R]; int16_t s16[VO_MAX_R];
    while (i_out < requested_output_symbols) {
        for (int i = 0; i < d->R; i++) {
            int32_t v;
            if (puncture_code[i_code]) {                 /* 1 = symbol was transmitted */
                if (i_in >= total_symbols) return i_in;  /* helpers:36-38 */
                v = load_soft(d, punctured_symbols, i_in++);
            } else {
                v = unpunctured_value;
            }
            s8[i] = (int8_t)v; s16[i] = (int16_t)v;
            i_code = (i_code + 1) % puncture_code_length;
            i_out++;
        }
        const int64_t a = vo_update(d, d->soft_bits == 8 ? (const void*)s8 : (const void*)s16, (size_t)d->R);
        if (a < 0) return i_in;
        if (acc) *acc += (uint64_t)a;
    }
    return i_in;
}

/* convolutional_encoder_shift_register.h:45-61 (input bits MSB-first into reg = reg<<1 | bit; output bit j = parity(G[j] & reg))
 * + test_helpers.h:17-64 (K-1 zero tail bits).  Output one byte per code bit in {0,1}, order [step][R]. */
size_t vo_encode(int K, int R, const uint32_t* G, const uint8_t* bytes, size_t nbytes, uint8_t* out_bits) {
    const uint64_t kmask = (K >= 64) ? ~0ull : (((uint64_t)1 << K) - 1u);
    uint64_t reg = 0;
    size_t n = 0;
    const size_t total = nbytes * 8 + (size_t)(K - 1);
    for (size_t t = 0; t < total; t++) {
        const unsigned bit = (t < nbytes * 8) ? ((bytes[t / 8] >> (7 - (t % 8))) & 1u) : 0u;
        reg = (reg << 1) | bit;
        for (int j = 0; j < R; j++) out_bits[n++] = (uint8_t)vo_parity(((uint64_t)G[j] & kmask) & reg);
    }
    return n;
}
