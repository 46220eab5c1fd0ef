// traceback.cuh -- device replacement for ViterbiDecoder_Core::chainback (include/viterbi/viterbi_decoder_core.h:214-236).
//
// The chain itself is serial (state_{j-1} depends on the decision read at step j), but unlike the reference layout the
// WORD ADDRESS never depends on the state here: for K <= 7 the whole decision row of a frame is one 64-bit word, so every
// load of the walk can be issued ahead of time and the kernel runs at memory speed instead of at memory latency.
// One thread per frame; 32 consecutive frames read 256 contiguous bytes per step.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_pipeline.h>

namespace vitb {

struct TracebackParams {
    const uint64_t* dec;      // [n_blocks][dec_rows][64]
    uint32_t dec_rows;
    uint32_t n_frames;
    uint32_t total_bits;      // L
    uint32_t state_bits;      // K-1
    uint32_t end_state;
    const uint32_t* end_states;   // nullable: per-frame end states (VITB_END_STATE_BEST), overrides end_state
    uint8_t* out;             // [n_frames][out_stride]
    size_t out_stride;
    uint32_t tag_layout;      // 0: bit s of the word = state s;  1: rows written by the tagged butterfly (acs_pair.cuh):
                              //    byte 2k+h holds states 2J+h for J = 8k..8k+7, first butterfly in the top bit
};

// end state of frame f: the call's end_state, or the frame's own (argmin of its final metrics) when the caller asked for the best one
template <class P>
__device__ __forceinline__ uint32_t frame_end_state(const P& p, size_t f) {
    return p.end_states ? p.end_states[f < p.n_frames ? f : p.n_frames - 1] : p.end_state;
}

// bit position of state s inside a frame's 64-bit decision word
__device__ __forceinline__ uint32_t dec_bit_index(uint32_t s, uint32_t tag_layout) {
    if (!tag_layout) return s;
    const uint32_t J = s >> 1;
    return (((J >> 3) * 2 + (s & 1u)) << 3) + 7u - (J & 7u);
}

// Decoded bit j is the decision read from row j + (K-1) at the current state; state <- (bit << (K-2)) | (state >> 1)
// (ViterbiTracebackBuffer::push_bit_in / get_state, core.h:96-113).  Bytes are MSB-first (core.h:234).  When
// total_bits % 8 != 0 the reference's last byte carries the leading bits of end_state below the decoded bits; same here.
template <int BATCH>
__global__ void __launch_bounds__(128) traceback_u64_kernel(const TracebackParams p) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const uint32_t SB = p.state_bits, L = p.total_bits;
    const uint64_t* d = p.dec + (size_t(f >> 6) * p.dec_rows + SB) * 64 + (f & 63);   // row of decoded bit 0
    uint8_t* out = p.out + size_t(f) * p.out_stride;
    const uint32_t es = frame_end_state(p, f);
    uint32_t state = es;
    int64_t j = int64_t(L) - 1;
    uint32_t byte = 0;
    if (L & 7) {   // virtual bits L, L+1, ... = end_state from its top bit down, then zeros
        for (uint32_t jj = L; jj < ((L + 7) & ~7u); jj++) {
            const uint32_t k = jj - L;
            const uint32_t b = (k < SB) ? ((es >> (SB - 1 - k)) & 1u) : 0u;
            byte |= b << (7 - (jj & 7));
        }
    }
    // ragged head so the main loop works on whole bytes
    const uint32_t tagl = p.tag_layout;
    while (j >= 0 && ((j & 7) != 7)) {
        const uint32_t bit = uint32_t(d[size_t(j) * 64] >> dec_bit_index(state, tagl)) & 1u;
        state = (bit << (SB - 1)) | (state >> 1);
        byte |= bit << (7 - (uint32_t(j) & 7));
        if ((j & 7) == 0) { out[j >> 3] = uint8_t(byte); byte = 0; }
        j--;
    }
    // main loop: BATCH rows in flight
    while (j >= BATCH - 1) {
        uint64_t w[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) w[k] = __ldcs(d + size_t(j - k) * 64);
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t bit = uint32_t(w[k] >> dec_bit_index(state, tagl)) & 1u;
            state = (bit << (SB - 1)) | (state >> 1);
            byte |= bit << (k & 7);                         // j-k has (j-k)&7 == 7-(k&7) because j&7 == 7 and BATCH%8 == 0
            if ((k & 7) == 7) { out[(j - k) >> 3] = uint8_t(byte); byte = 0; }
        }
        j -= BATCH;
    }
    while (j >= 0) {
        const uint32_t bit = uint32_t(d[size_t(j) * 64] >> dec_bit_index(state, tagl)) & 1u;
        state = (bit << (SB - 1)) | (state >> 1);
        byte |= bit << (7 - (uint32_t(j) & 7));
        if ((j & 7) == 0) { out[j >> 3] = uint8_t(byte); byte = 0; }
        j--;
    }
}

// ---- decision rows written by acs_group_kernel (acs_group.cuh) -------------------------------------------------------------
// T = 2^logt lanes serve one frame pair, with the same lane <-> group mapping as the ACS kernel: lane t loads ITS OWN decision
// words (coalesced, addresses independent of the state, so whole batches of rows are in flight), and the word that holds the
// decision of the current state is fetched from lane t(state) with one __shfl per frame and row.  All lanes of the group track
// both states redundantly; lane 0 writes frame A's bytes, lane 1 frame B's.
struct TracebackGroupParams {
    const uint32_t* dec;      // [n_wblocks][dec_rows][32][W]
    uint32_t dec_rows;
    uint32_t n_frames;
    uint32_t total_bits;
    uint32_t state_bits;      // SB = K-1
    uint32_t logt;            // lanes per pair = 1 << logt (>= 1)
    uint32_t end_state;
    const uint32_t* end_states;   // nullable: per-frame end states (VITB_END_STATE_BEST), overrides end_state
    uint8_t* out;
    size_t out_stride;
};

__device__ __forceinline__ uint32_t rotr_rt(uint32_t v, uint32_t r, uint32_t n) {
    r %= n;
    return r == 0 ? v : (((v >> r) | (v << (n - r))) & ((1u << n) - 1u));
}

// grid = ceil(n_wblocks / 4), block = 128 (4 warps, one warp block each)
template <int W, int BATCH>
__global__ void __launch_bounds__(128) traceback_group_kernel(const TracebackGroupParams p) {
    const uint32_t lane = threadIdx.x & 31, wblk = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t SB = p.state_bits, g = p.logt, T = 1u << g, LB = SB - g, PPW = 32u >> g, L = p.total_bits;
    const uint32_t n_pairs = (p.n_frames + 1) / 2;
    if (size_t(wblk) * PPW >= n_pairs) return;
    const uint32_t pw = lane >> g, t = lane & (T - 1), grp0 = lane & ~(T - 1);
    const size_t fA = (size_t(wblk) * PPW + pw) * 2;
    const uint32_t* d = p.dec + (size_t(wblk) * p.dec_rows * 32 + lane) * W;      // this lane's word(s) of row 0
    const uint32_t esA = frame_end_state(p, fA), esB = frame_end_state(p, fA + 1);
    uint32_t stA = esA, stB = esB;
    uint32_t byteA = 0, byteB = 0;
    const bool writerA = (t == 0) && (fA < p.n_frames), writerB = (t == 1) && (fA + 1 < p.n_frames);
    uint8_t* outA = p.out + fA * p.out_stride;
    uint8_t* outB = outA + p.out_stride;
    if (L & 7) {
        for (uint32_t jj = L; jj < ((L + 7) & ~7u); jj++) {
            const uint32_t k = jj - L;
            byteA |= ((k < SB) ? ((esA >> (SB - 1 - k)) & 1u) : 0u) << (7 - (jj & 7));
            byteB |= ((k < SB) ? ((esB >> (SB - 1 - k)) & 1u) : 0u) << (7 - (jj & 7));
        }
    }
    const uint32_t smask = (1u << SB) - 1u;
    // n = row % LB is carried along (rows are visited in decreasing order); rot = n + 1 is in [1, LB] so never 0 or SB
    auto one_row = [&](int64_t j, uint32_t n, uint32_t w0, uint32_t w1) {
        const uint32_t rot = n + 1;
        const uint32_t phiA = ((stA >> rot) | (stA << (SB - rot))) & smask, phiB = ((stB >> rot) | (stB << (SB - rot))) & smask;
        const uint32_t qA = phiA >> g, qB = phiB >> g;
        const uint32_t wA = __shfl_sync(0xffffffffu, w0, int(grp0 + (phiA & (T - 1))));
        const uint32_t wB = __shfl_sync(0xffffffffu, W == 1 ? w0 : w1, int(grp0 + (phiB & (T - 1))));
        const uint32_t bitA = (wA >> qA) & 1u;
        const uint32_t bitB = (wB >> (W == 1 ? 16 + qB : qB)) & 1u;
        stA = (bitA << (SB - 1)) | (stA >> 1);
        stB = (bitB << (SB - 1)) | (stB >> 1);
        byteA |= bitA << (7 - (uint32_t(j) & 7));
        byteB |= bitB << (7 - (uint32_t(j) & 7));
        if ((j & 7) == 0) {
            if (writerA) outA[j >> 3] = uint8_t(byteA);
            if (writerB) outB[j >> 3] = uint8_t(byteB);
            byteA = 0; byteB = 0;
        }
    };
    int64_t j = int64_t(L) - 1;
    uint32_t n = uint32_t((uint64_t(j) + SB) % LB);     // phase of the top row
    while (j >= 0 && ((j + 1) % BATCH) != 0) {      // ragged head
        const uint32_t* r = d + (size_t(j) + SB) * 32 * W;
        one_row(j, n, r[0], W == 2 ? r[W - 1] : 0u);
        n = (n == 0) ? LB - 1 : n - 1;
        j--;
    }
    // main loop, software pipelined: the next batch of rows is already in flight while this one is walked
    uint32_t nw0[BATCH], nw1[BATCH];
    if (j >= BATCH - 1) {
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            const uint32_t* r = d + (size_t(j - k) + SB) * 32 * W;
            if (W == 2) { const uint2 v = __ldcs(reinterpret_cast<const uint2*>(r)); nw0[k] = v.x; nw1[k] = v.y; }
            else { nw0[k] = __ldcs(r); nw1[k] = 0u; }
        }
    }
    while (j >= BATCH - 1) {
        uint32_t w0[BATCH], w1[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++) { w0[k] = nw0[k]; w1[k] = nw1[k]; }
        if (j - BATCH >= BATCH - 1) {
#pragma unroll
            for (int k = 0; k < BATCH; k++) {
                const uint32_t* r = d + (size_t(j - BATCH - k) + SB) * 32 * W;
                if (W == 2) { const uint2 v = __ldcs(reinterpret_cast<const uint2*>(r)); nw0[k] = v.x; nw1[k] = v.y; }
                else { nw0[k] = __ldcs(r); nw1[k] = 0u; }
            }
        }
#pragma unroll
        for (int k = 0; k < BATCH; k++) {
            one_row(j - k, n, w0[k], w1[k]);
            n = (n == 0) ? LB - 1 : n - 1;
        }
        j -= BATCH;
    }
}

// Staged variant of the above: a warp walks 32 FRAMES (16 pairs) at once instead of 32/T.  The decision rows of those pairs
// (32 bytes per frame and row, state independent) are streamed into shared memory with cp.async, two batches of ROWS rows in
// flight, and every lane picks the one word its state needs with a single LDS.  8-16x fewer instructions per frame than the
// shuffle version, which replicates the walk in all T lanes of a pair.
// grid = ceil(n_frames / 128), block = 128 (4 warps), dynamic smem = 4 * 2 * ROWS * row_words * 4 bytes
template <int W, int ROWS>
__global__ void __launch_bounds__(128) traceback_group_staged_kernel(const TracebackGroupParams p) {
    extern __shared__ uint32_t tb_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t SB = p.state_bits, g = p.logt, T = 1u << g, LB = SB - g, PPW = 32u >> g, L = p.total_bits;
    const uint32_t fbase = (blockIdx.x * 4 + warp) * 32;
    if (fbase >= p.n_frames) return;
    const uint32_t n_wb = 16u / PPW;                       // warp blocks covered by this warp's 16 pairs (PPW <= 16)
    const uint32_t wb_words = 32u * W;                     // words of one warp block per row
    const uint32_t row_words = n_wb * wb_words;            // = 16 * T * W
    uint32_t* stage = tb_smem + size_t(warp) * 2 * ROWS * row_words;
    const size_t wblk0 = size_t(fbase / 2) / PPW;
    const uint32_t* dbase = p.dec + wblk0 * p.dec_rows * wb_words;     // row r of warp block w: dbase + (w * dec_rows + r) * wb_words

    const uint32_t f = fbase + lane, pr = lane >> 1, half = lane & 1u;
    const uint32_t my_off = ((pr / PPW) * 32u + (pr % PPW) * T) * W + (W == 2 ? half : 0u);   // + t * W selects the lane's word
    const uint32_t bit_base = (W == 2) ? 0u : half * 16u;
    const bool writer = f < p.n_frames;
    uint8_t* out = p.out + size_t(f) * p.out_stride;
    const uint32_t smask = (1u << SB) - 1u;

    const uint32_t es = frame_end_state(p, f);
    uint32_t byte = 0;
    if (L & 7) {
        for (uint32_t jj = L; jj < ((L + 7) & ~7u); jj++) {
            const uint32_t k = jj - L;
            const uint32_t b = (k < SB) ? ((es >> (SB - 1 - k)) & 1u) : 0u;
            byte |= b << (7 - (jj & 7));
        }
    }
    // batch b covers rows top - b*ROWS - (ROWS-1) .. top - b*ROWS (clamped at 0; rows below SB are loaded but not walked)
    const int top = int(L + SB) - 1;
    const int n_batches = int((L + ROWS - 1) / ROWS);
    // all sizes are powers of two: 16-byte pieces per row, per warp block
    const uint32_t lg_row_pieces = 31u - uint32_t(__clz(row_words / 4)), lg_wb_pieces = 31u - uint32_t(__clz(wb_words / 4));
    const uint32_t pieces = uint32_t(ROWS) << lg_row_pieces;
    auto issue = [&](int b) {
        uint32_t* dst = stage + size_t(b & 1) * ROWS * row_words;
        for (uint32_t pc = lane; pc < pieces; pc += 32) {
            const uint32_t k = pc >> lg_row_pieces, in_row = pc & ((1u << lg_row_pieces) - 1u);   // k-th row from the top of the batch
            int r = top - b * ROWS - int(k);
            if (r < 0) r = 0;
            const uint32_t w = in_row >> lg_wb_pieces, o = in_row & ((1u << lg_wb_pieces) - 1u);
            __pipeline_memcpy_async(dst + pc * 4, dbase + (size_t(w) * p.dec_rows + size_t(r)) * wb_words + o * 4, 16);
        }
        __pipeline_commit();
    };
    issue(0);
    // The walk tracks PHI = rotr^(n+1)(state) instead of the state: inside an exchange period going one row down only replaces
    // the bit at position SB-1-n by the decision just read; at a period boundary PHI is re-rotated once.
    uint32_t n = uint32_t(top) % LB;
    uint32_t phi;
    {
        const uint32_t rot = n + 1, st = es;
        phi = ((st >> rot) | (st << (SB - rot))) & smask;
    }
    const uint32_t lane_mask = T - 1;
    for (int b = 0; b < n_batches; b++) {
        if (b + 1 < n_batches) { issue(b + 1); __pipeline_wait_prior(1); } else { __pipeline_wait_prior(0); }
        __syncwarp();
        const uint32_t* src = stage + size_t(b & 1) * ROWS * row_words + my_off;
        const int r_top = top - b * ROWS;
#pragma unroll
        for (int k = 0; k < ROWS; k++) {
            const int r = r_top - k;
            if (r < int(SB)) break;
            const uint32_t wv = src[uint32_t(k) * row_words + (phi & lane_mask) * W];
            const uint32_t bit = (wv >> (bit_base + (phi >> g))) & 1u;
            if (n != 0) {                                     // uniform: n depends on the row only
                const uint32_t pos = SB - 1 - n;
                phi = (phi & ~(1u << pos)) | (bit << pos);
                n--;
            } else {                                          // PHI = rotr^1(state_r): replacing its top bit gives state_(r-1);
                const uint32_t st = (phi & (smask >> 1)) | (bit << (SB - 1));   // the row below has phase LB-1: PHI = rotr^LB
                phi = ((st >> LB) | (st << g)) & smask;       // SB - LB = g
                n = LB - 1;
            }
            const uint32_t j = uint32_t(r) - SB;
            byte |= bit << (7 - (j & 7));
            if ((j & 7) == 0) { if (writer) out[j >> 3] = uint8_t(byte); byte = 0; }
        }
        __syncwarp();       // everyone done with this buffer before it is refilled two batches later
    }
}

// ---- decision rows written by acs_cta_kernel (acs_cta.cuh): [pair][row][T threads] x {A word, B word} -----------------------
// A row is 2 KB per frame, so unlike the short codes the walk must read only the word it needs: address = f(state).  Within one
// exchange period (LB rows) the owning thread t(state) does not change, so the LB loads of a period are issued together; the
// next period's address needs this period's decisions (a dependent chain of L/LB memory round trips per frame, all frames in
// parallel).  One thread per frame.
struct TracebackCtaParams {
    uint32_t words;           // 2: {A word, B word} per thread and row (32 registers per thread); 1: A bits | B bits << 16 (16 registers)
    const uint32_t* dec;
    uint32_t dec_rows;
    uint32_t n_frames;
    uint32_t total_bits;
    uint32_t state_bits;
    uint32_t logt;
    uint32_t end_state;
    const uint32_t* end_states;   // nullable: per-frame end states (VITB_END_STATE_BEST), overrides end_state
    uint8_t* out;
    size_t out_stride;
    uint32_t ring_rows;       // sliding-window calls: row r lives at r % ring_rows (a multiple of the exchange period); 0 = every row is kept
};

// Walk decision rows r_hi .. r_lo (downwards) of one frame from `state` (the state after row r_hi); returns the state after row
// r_lo - 1 ... i.e. the state the walk arrives with below r_lo.  WRITE: decoded bit j = r - SB goes to the output (MSB-first bytes,
// flushed when j is a multiple of 8; `byte` carries the bits above r_hi that share its byte - only the ragged top of a frame has any).
template <int MAXLB, bool WRITE>
__device__ __forceinline__ uint32_t cta_walk(const TracebackCtaParams& p, const uint32_t* d, uint8_t* out, uint32_t bit_base,
                                             int64_t r_hi, int64_t r_lo, uint32_t state, uint32_t byte) {
    const uint32_t SB = p.state_bits, g = p.logt, T = 1u << g, LB = SB - g, W = p.words;
    int64_t r = r_hi;
    while (r >= r_lo) {
        const uint32_t n = uint32_t(r) % LB;               // rows r, r-1, .., r-n belong to the same exchange period
        int64_t r_lo_p = r - n;
        if (r_lo_p < r_lo) r_lo_p = r_lo;
        const uint32_t cnt = uint32_t(r - r_lo_p) + 1;
        const uint32_t t = rotr_rt(state, n + 1, SB) & (T - 1);
        uint32_t w[MAXLB];
#pragma unroll
        for (int k = 0; k < MAXLB; k++) {
            const size_t row = p.ring_rows ? size_t(uint64_t(r - k) % p.ring_rows) : size_t(r - k);
            w[k] = (uint32_t(k) < cnt) ? __ldcs(d + (row * T + t) * W) : 0u;
        }
#pragma unroll
        for (int k = 0; k < MAXLB; k++) {
            if (uint32_t(k) < cnt) {
                const uint32_t q = rotr_rt(state, n - uint32_t(k) + 1, SB) >> g;
                const uint32_t bit = (w[k] >> (bit_base + q)) & 1u;
                state = (bit << (SB - 1)) | (state >> 1);
                if (WRITE) {
                    const int64_t j = r - k - int64_t(SB);
                    byte |= bit << (7 - (uint32_t(j) & 7));
                    if ((j & 7) == 0) { out[j >> 3] = uint8_t(byte); byte = 0; }
                }
            }
        }
        r = r_lo_p - 1;
    }
    return state;
}

// virtual bits L, L+1, ... of a ragged last byte = end_state from its top bit down, then zeros (the reference's traceback buffer
// still holds them, core.h:96-113)
__device__ __forceinline__ uint32_t ragged_top_byte(uint32_t L, uint32_t SB, uint32_t end_state) {
    uint32_t byte = 0;
    if (L & 7) {
        for (uint32_t jj = L; jj < ((L + 7) & ~7u); jj++) {
            const uint32_t k = jj - L;
            const uint32_t b = (k < SB) ? ((end_state >> (SB - 1 - k)) & 1u) : 0u;
            byte |= b << (7 - (jj & 7));
        }
    }
    return byte;
}

// one thread per frame, the whole chain
template <int MAXLB>
__global__ void __launch_bounds__(64) traceback_cta_kernel(const TracebackCtaParams p) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const uint32_t SB = p.state_bits, T = 1u << p.logt, L = p.total_bits, half = f & 1u, W = p.words;
    const uint32_t* d = p.dec + size_t(f >> 1) * p.dec_rows * T * W + (W == 2 ? half : 0u);
    const uint32_t bit_base = (W == 2) ? 0u : half * 16u;
    uint8_t* out = p.out + size_t(f) * p.out_stride;
    const uint32_t es = frame_end_state(p, f);
    (void)cta_walk<MAXLB, true>(p, d, out, bit_base, int64_t(L) + SB - 1, SB, es, ragged_top_byte(L, SB, es));
}

// Segmented walk (same scheme as traceback_hist_seg_kernel below): segments of seg_bits decoded bits (a multiple of 8), warm-up over
// `overlap` rows from state 0, the arriving state recorded for the check against the segment above.
struct TracebackCtaSegParams {
    uint32_t n_seg, seg_bits, overlap;
    uint32_t* spec;           // [n_seg][n_frames]
    uint32_t* fin;            // [n_seg][n_frames]
    uint32_t seg0;            // first segment of this launch (grid.y counts from it): sliding-window calls walk one segment per launch
};

// grid = (ceil(n_frames / 64), n_seg)
template <int MAXLB>
__global__ void __launch_bounds__(64) traceback_cta_seg_kernel(const TracebackCtaParams p, const TracebackCtaSegParams sp) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y + sp.seg0;
    if (f >= p.n_frames) return;
    const uint32_t SB = p.state_bits, T = 1u << p.logt, L = p.total_bits, half = f & 1u, W = p.words;
    const uint32_t* d = p.dec + size_t(f >> 1) * p.dec_rows * T * W + (W == 2 ? half : 0u);
    const uint32_t bit_base = (W == 2) ? 0u : half * 16u;
    uint8_t* out = p.out + size_t(f) * p.out_stride;
    const int64_t r_top = int64_t(L) + SB - 1;
    const int64_t r_lo = int64_t(g) * sp.seg_bits + SB;
    int64_t r_hi = r_lo + sp.seg_bits - 1;
    const uint32_t es = frame_end_state(p, f);
    uint32_t state, byte = 0;
    if (g + 1 == sp.n_seg || r_hi >= r_top) {
        r_hi = r_top;
        state = es;                                           // the top segment starts from the truth
        byte = ragged_top_byte(L, SB, es);
    } else {
        int64_t r_warm = r_hi + sp.overlap;
        if (r_warm >= r_top) { r_warm = r_top; state = es; }
        else state = 0u;
        state = cta_walk<MAXLB, false>(p, d, out, bit_base, r_warm, r_hi + 1, state, 0u);
        sp.spec[size_t(g) * p.n_frames + f] = state;
    }
    sp.fin[size_t(g) * p.n_frames + f] = cta_walk<MAXLB, true>(p, d, out, bit_base, r_hi, r_lo, state, byte);
}

// Sliding-window calls cannot repair (the rows of a segment are overwritten before the segment above has been walked): they count,
// per call, the segment boundaries at which the warm-up from state 0 did NOT arrive in the state the segment above ended in - 0
// means the windowed result is the exact one.
__global__ void __launch_bounds__(128) traceback_seg_count_mismatch_kernel(const uint32_t* spec, const uint32_t* fin, uint32_t n_seg,
                                                                             uint32_t n_frames, unsigned long long* count) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    uint32_t bad = 0;
    for (uint32_t g = 0; g + 1 < n_seg; g++) bad += spec[size_t(g) * n_frames + f] != fin[size_t(g + 1) * n_frames + f];
    if (bad) atomicAdd(count, (unsigned long long)bad);
}

// one thread per frame: top-down check of the segment boundaries, re-walk on mismatch
template <int MAXLB>
__global__ void __launch_bounds__(64) traceback_cta_fix_kernel(const TracebackCtaParams p, const TracebackCtaSegParams sp) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    uint32_t truth = sp.fin[size_t(sp.n_seg - 1) * p.n_frames + f];
    for (int g = int(sp.n_seg) - 2; g >= 0; g--) {
        const size_t i = size_t(g) * p.n_frames + f;
        if (sp.spec[i] == truth) { truth = sp.fin[i]; continue; }
        const uint32_t SB = p.state_bits, T = 1u << p.logt, half = f & 1u, W = p.words;
        const uint32_t* d = p.dec + size_t(f >> 1) * p.dec_rows * T * W + (W == 2 ? half : 0u);
        const int64_t r_lo = int64_t(g) * sp.seg_bits + SB;
        truth = cta_walk<MAXLB, true>(p, d, p.out + size_t(f) * p.out_stride, (W == 2) ? 0u : half * 16u, r_lo + sp.seg_bits - 1, r_lo, truth, 0u);
    }
}

}  // namespace vitb

namespace vitb {

// ---- history records written by acs_hist_kernel (acs_hist.cuh) ---------------------------------------------------------------
// Record r of a frame holds, for every state s at time min(HB(r+1), S), the decisions D[HB r .. HB r + HB-1] along the survivor path
// that ends in s (bit k = step HB r + k; HB = 8 for uint8_t metrics, 16 for uint16_t metrics).  chainback (core.h:214-236) reads, for
// decoded bit j, the decision of row j + (K-1) at the current state and then moves to state (bit << (K-2)) | (state >> 1): along
// the survivor path those are exactly the bits of the history field, so one lookup replaces HB dependent row lookups:
//     decoded[j] = D[j + SB];   state one period earlier = bit-reversal of the low SB bits of the field (SB <= 8 <= HB).
// Bytes are MSB-first (core.h:234); when total_bits % 8 != 0 the last byte carries the leading bits of end_state below the decoded
// bits (the reference's traceback buffer still holds them), modelled here as virtual decisions D[S + k] = bit (SB-1-k) of end_state.
struct TracebackHistParams {
    const uint8_t* dec;       // records, layout of acs_hist.cuh
    uint32_t n_periods;
    uint32_t n_frames;
    uint32_t total_bits;      // L
    uint32_t state_bits;      // SB = K-1 (<= 8)
    uint32_t end_state;
    const uint32_t* end_states;   // nullable: per-frame end states (VITB_END_STATE_BEST), overrides end_state
    uint32_t n_steps;         // S = L + SB
    uint32_t hist_bits;       // HB: 8 (two frames per lane, 64 per block) or 16 (one frame per lane, 32 per block)
    uint32_t logt;            // acs_hist_group.cuh: a frame spans 2^logt lanes (HB = 16), records in position order; else 0
    uint8_t* out;
    size_t out_stride;
};

// Walk records r_hi .. r_lo (downwards) of one frame.  `state` is the state at the boundary above r_hi (time min(HB (r_hi + 1), S)).
// The history above a record only enters its output through its low SB bits, and those are the bit-reversed state at the
// boundary between the two records, so a walk can start at any record boundary from the state alone.  Returns the state at the
// boundary below r_lo.
struct HistFrame {
    const uint8_t* base;      // first record of the frame's warp block
    uint8_t* out;
    size_t rec_bytes;
    uint32_t lane, half, vw_log, row_log, SB, HB, r_last, nv, VS, n_out, end_state, logt, n_steps;
    bool wide;
};

__device__ __forceinline__ HistFrame hist_frame(const TracebackHistParams& p, uint32_t f) {
    HistFrame c;
    c.SB = p.state_bits; c.HB = p.hist_bits; c.wide = c.HB == 16;
    c.logt = p.logt; c.n_steps = p.n_steps;
    const uint32_t NS = 1u << c.SB, nw = (NS >> c.logt) / 2;     // record words per lane
    c.vw_log = nw >= 4 ? 2u : (nw == 2 ? 1u : 0u);
    // logt >= 5 (acs_hist_cta.cuh): a frame IS a block of 2^logt threads; else a block is one warp
    const bool cta = c.logt >= 5u;
    const uint32_t fpb_log = cta ? 0u : (c.wide ? (5u - c.logt) : 6u);   // frames per block: 1, or 64, 32, 32 >> logt per warp
    c.row_log = cta ? c.logt : 5u;                               // threads per row of 16-byte record chunks
    c.lane = cta ? 0u : (c.wide ? ((f & ((1u << fpb_log) - 1u)) << c.logt) : ((f & 63u) >> 1));      // first lane of the frame
    c.half = c.wide ? 0u : (f & 1u);
    c.rec_bytes = cta ? size_t(2) * NS : ((size_t(64) * NS) >> c.logt);   // one block, one period
    c.base = p.dec + size_t(f >> fpb_log) * p.n_periods * c.rec_bytes;
    c.out = p.out + size_t(f) * p.out_stride;
    c.n_out = (p.total_bits + 7) / 8;
    c.r_last = (p.n_steps - 1) / c.HB;
    c.nv = p.n_steps - c.HB * c.r_last;
    c.end_state = frame_end_state(p, f);
    const uint32_t E = c.SB ? (__brev(c.end_state) >> (32 - c.SB)) : 0u;     // virtual decisions behind the last step
    c.VS = E << c.nv;                                            // relative to the first step of record r_last (at most 16 + 8 bits)
    return c;
}

template <bool WRITE>
__device__ __forceinline__ uint32_t hist_walk(const HistFrame& c, int64_t r_hi, int64_t r_lo, uint32_t state) {
    const uint32_t SB = c.SB, HB = c.HB, hmask = (1u << HB) - 1u;
    for (int64_t r = r_hi; r >= r_lo; r--) {
        const bool last = uint32_t(r) == c.r_last;
        // history of the record above, as far as this record's output needs it
        const uint32_t hnext = last ? ((c.VS >> HB) & hmask) : (SB ? (__brev(state) >> (32 - SB)) : 0u);
        // where the history of `state` sits in the record: register `reg` of lane `ln`.  One lane per frame (pair): register = state.
        // Frame over 2^logt lanes (acs_hist_group.cuh): position PHI = rotr^m(state), m = steps done at the end of the record mod LB.
        uint32_t reg = state, ln = c.lane;
        if (c.logt) {
            const uint32_t done = last ? c.n_steps : HB * (uint32_t(r) + 1u);
            const uint32_t phi = rotr_rt(state, done % (SB - c.logt), SB);
            reg = phi >> c.logt;
            ln = c.lane | (phi & ((1u << c.logt) - 1u));
        }
        // word w = reg >> 1 of the lane: FMT 0 bytes [A:2w, B:2w, A:2w+1, B:2w+1], FMT 1 halfwords [2w, 2w+1]
        const uint32_t w = reg >> 1;
        const uint32_t off = ((((w >> c.vw_log) << c.row_log) + ln) << (c.vw_log + 2)) + ((w & ((1u << c.vw_log) - 1u)) << 2) + ((reg & 1u) << 1) + c.half;
        const uint8_t* ptr = c.base + size_t(r) * c.rec_bytes + off;
        const uint32_t h = c.wide ? uint32_t(*reinterpret_cast<const uint16_t*>(ptr)) : uint32_t(*ptr);
        uint32_t hext = h;
        if (last) {
            hext = (h & ((1u << c.nv) - 1u)) | (c.VS & hmask);
            for (int k = int(c.nv) - 1; k >= 0; k--) state = (((h >> k) & 1u) << (SB - 1)) | (state >> 1);
        } else {
            state = __brev(h) >> (32 - SB);            // one period back: the last SB pushed bits, newest on top
        }
        if (WRITE) {
            // decoded bits of this record's steps: window of 2 HB decisions shifted down by SB, MSB-first bytes
            const uint64_t win = (uint64_t(hext) | (uint64_t(hnext) << HB)) >> SB;
            if (!c.wide) {
                if (uint32_t(r) < c.n_out) c.out[r] = uint8_t(__brev(uint32_t(win) & 0xffu) >> 24);
            } else {
                const uint32_t b0 = 2u * uint32_t(r);
                if (b0 < c.n_out) c.out[b0] = uint8_t(__brev(uint32_t(win) & 0xffu) >> 24);
                if (b0 + 1 < c.n_out) c.out[b0 + 1] = uint8_t(__brev(uint32_t(win >> 8) & 0xffu) >> 24);
            }
        }
    }
    return state;
}

// one thread per frame, the whole chain
__global__ void __launch_bounds__(128) traceback_hist_kernel(const TracebackHistParams p) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const HistFrame c = hist_frame(p, f);
    (void)hist_walk<true>(c, c.r_last, 0, c.end_state);
}

// ---- segmented walk -----------------------------------------------------------------------------------------------------------
// The chain of a frame is S / HB dependent memory round trips (about 630 ns each on B200), which is what a traceback launch costs
// when the batch has too few frames to saturate DRAM.  Survivor paths merge: a walk started `overlap` records above a segment from
// ANY state has, with overwhelming probability, joined the true path by the time it reaches the segment.  So every frame is cut
// into n_seg segments of seg_records records that are walked concurrently, each (except the top one, which starts from the true end
// state) after a warm-up over `overlap` records from state 0; the state it arrives with is recorded next to the state the segment
// above ended in, and traceback_hist_fix_kernel re-walks, serially and from the true state, any segment whose two states differ.
// The result is therefore exact, not probabilistic.
struct TracebackSegParams {
    uint32_t n_seg, seg_records, overlap;
    uint32_t* spec;           // [n_seg][n_frames] state a segment's warm-up arrived with (top segment: unused)
    uint32_t* fin;            // [n_seg][n_frames] state at the segment's lower boundary
};

// grid = (ceil(n_frames / 128), n_seg)
__global__ void __launch_bounds__(128) traceback_hist_seg_kernel(const TracebackHistParams p, const TracebackSegParams sp) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
    if (f >= p.n_frames) return;
    const HistFrame c = hist_frame(p, f);
    const int64_t r_lo = int64_t(g) * sp.seg_records;
    int64_t r_hi = r_lo + sp.seg_records - 1;
    uint32_t state;
    if (g + 1 == sp.n_seg || r_hi >= int64_t(c.r_last)) {
        r_hi = c.r_last;
        state = c.end_state;                                  // the top segment starts from the truth
    } else {
        int64_t r_warm = r_hi + sp.overlap;
        if (r_warm >= int64_t(c.r_last)) { r_warm = c.r_last; state = c.end_state; }
        else state = 0u;
        state = hist_walk<false>(c, r_warm, r_hi + 1, state);
        sp.spec[size_t(g) * p.n_frames + f] = state;
    }
    sp.fin[size_t(g) * p.n_frames + f] = hist_walk<true>(c, r_hi, r_lo, state);
}

// one thread per frame: top-down check of the segment boundaries, re-walk on mismatch
__global__ void __launch_bounds__(128) traceback_hist_fix_kernel(const TracebackHistParams p, const TracebackSegParams sp) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    uint32_t truth = sp.fin[size_t(sp.n_seg - 1) * p.n_frames + f];
    for (int g = int(sp.n_seg) - 2; g >= 0; g--) {
        const size_t i = size_t(g) * p.n_frames + f;
        if (sp.spec[i] == truth) { truth = sp.fin[i]; continue; }
        const HistFrame c = hist_frame(p, f);
        const int64_t r_lo = int64_t(g) * sp.seg_records;
        truth = hist_walk<true>(c, r_lo + sp.seg_records - 1, r_lo, truth);
    }
}

}  // namespace vitb
