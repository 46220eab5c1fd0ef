// ingest.cuh -- front end of the batch decoder: soft symbols as the caller lays them out -> the packed stream the ACS
// kernels read with fully coalesced 128-byte warp loads.
//
// Caller layout (the reference's, one frame after another): symbols[frame][step][R] of soft_t (int8_t / int16_t), what
// Decoder::update(base, symbols, N) walks (viterbi_decoder_scalar.h:43-46).  Optionally punctured: a depuncture map gives, for
// every depunctured symbol index, the index of the received symbol in the frame's row or -1 for "insert the unpunctured
// value", which is decode_punctured_symbols' behaviour (examples/helpers/puncture_code_helpers.h:17-55) done as a gather.
//
// Pair stream: frame pair gp = (frames 2gp, 2gp+1) belongs to warp block gp / PPW at slot gp % PPW (PPW = pairs per warp of the
// ACS kernel: 32 for one-thread-per-pair, 32/T when a pair spans T lanes);
//     word[(wblk * n_sym + e) * PPW + slot] = (sA << SH) & 0xffff | (sB << SH) << 16,   e = step*R + i.
// SH = 8 for uint8_t error metrics (held as metric << 8), else 0.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vitb {

struct IngestParams {
    const void* symbols;      // device pointer, soft_t
    size_t row_stride;        // elements between consecutive frames
    uint32_t n_frames;
    uint32_t n_sym;           // depunctured symbols per frame in this call (= steps * R)
    const int32_t* depuncture_map;  // nullable; [n_sym] -> received index or -1
    int32_t fill_value;       // unpunctured symbol value
    uint32_t* pk;             // output
    uint32_t ppw;             // pairs per warp block (power of two, 1..32)
};

constexpr int INGEST_TILE = 256;

// grid = (ceil(n_sym / 256), ceil(n_frames / 64)), block = 256 threads.  64 frames x 256 symbols transposed through shared memory.
template <typename soft_t, int SH>
__global__ void __launch_bounds__(INGEST_TILE) ingest_pairs_kernel(const IngestParams p) {
    constexpr int RS = INGEST_TILE + 1;        // odd row stride (in int16) -> conflict-free column reads
    __shared__ uint16_t tile[64 * RS];
    const uint32_t blk = blockIdx.y, e0 = blockIdx.x * INGEST_TILE, tid = threadIdx.x;
    const soft_t* sym = static_cast<const soft_t*>(p.symbols);

    const uint32_t e = e0 + tid;
    int32_t src = -2;                          // -2: beyond the row, -1: punctured
    if (e < p.n_sym) src = p.depuncture_map ? p.depuncture_map[e] : int32_t(e);
    // Column `tid` of the tile: one symbol position of 64 consecutive frames.  Everything about the position (received index,
    // punctured or not) is loop invariant, so the loop over the frames is a pointer walk (this kernel was issue bound at 25
    // instructions per element when the address was recomputed per frame: profiles/r01_summary.md).
    const uint32_t rows = (p.n_frames - blk * 64u < 64u) ? (p.n_frames - blk * 64u) : 64u;
    uint16_t* tcol = tile + tid;
    if (src >= 0) {
        const soft_t* q = sym + size_t(blk) * 64 * p.row_stride + size_t(src);
        if (rows == 64u) {
#pragma unroll 16
            for (int r = 0; r < 64; r++) { tcol[r * RS] = uint16_t(uint32_t(int32_t(*q)) << SH); q += p.row_stride; }
        } else {
            for (uint32_t r = 0; r < 64u; r++) {
                tcol[r * RS] = (r < rows) ? uint16_t(uint32_t(int32_t(*q)) << SH) : uint16_t(0);
                q += p.row_stride;
            }
        }
    } else {
        const uint16_t v = (src == -1) ? uint16_t(uint32_t(p.fill_value) << SH) : uint16_t(0);
#pragma unroll 16
        for (uint32_t r = 0; r < 64u; r++) tcol[r * RS] = (r < rows) ? v : uint16_t(0);
    }
    __syncthreads();
    const uint32_t ppw = p.ppw;
    const uint32_t n_valid = (p.n_sym - e0 < uint32_t(INGEST_TILE)) ? (p.n_sym - e0) : uint32_t(INGEST_TILE);
    // 32 pairs x 256 symbols leave as runs of (256 * ppw) words, one run per warp block touched by this tile (ppw is a power of two):
    // word ((wblk * n_sym + e0 + k) << lp) + slot.  A thread keeps its slot and walks k; consecutive threads write consecutive words.
    const uint32_t lp = 31u - uint32_t(__clz(ppw)), slot = tid & (ppw - 1u), kk = tid >> lp, kstep = uint32_t(INGEST_TILE) >> lp;
    for (uint32_t wb = 0; wb < (32u >> lp); wb++) {
        const uint32_t pr = (wb << lp) + slot;               // pair inside the tile
        const uint16_t* ta = tile + (2u * pr) * RS;
        uint32_t* outw = p.pk + (((size_t(blk) * (32u >> lp) + wb) * p.n_sym + e0) << lp) + slot;
        if (n_valid == uint32_t(INGEST_TILE)) {
#pragma unroll 4
            for (uint32_t k = kk; k < uint32_t(INGEST_TILE); k += kstep) outw[size_t(k) << lp] = uint32_t(ta[k]) | (uint32_t(ta[RS + k]) << 16);
        } else {
            for (uint32_t k = kk; k < n_valid; k += kstep) outw[size_t(k) << lp] = uint32_t(ta[k]) | (uint32_t(ta[RS + k]) << 16);
        }
    }
}

}  // namespace vitb
