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
#pragma unroll 8
    for (int r = 0; r < 64; r++) {
        const size_t f = size_t(blk) * 64 + r;
        int32_t v = 0;
        if (f < p.n_frames) {
            if (src >= 0) v = int32_t(sym[f * p.row_stride + size_t(src)]);
            else if (src == -1) v = p.fill_value;
        }
        tile[r * RS + tid] = uint16_t(uint32_t(v) << SH);
    }
    __syncthreads();
    const uint32_t ppw = p.ppw;
    const uint32_t n_valid = (p.n_sym - e0 < uint32_t(INGEST_TILE)) ? (p.n_sym - e0) : uint32_t(INGEST_TILE);
    if (ppw == 32) {
        // one warp block per tile: lane = pair, 8 symbols in flight per pass
        const uint32_t lane = tid & 31, k0 = tid >> 5;
        uint32_t* out = p.pk + (size_t(blk) * p.n_sym + e0) * 32 + lane;
#pragma unroll 4
        for (uint32_t k = k0; k < uint32_t(INGEST_TILE); k += INGEST_TILE / 32) {
            if (k < n_valid) {
                const uint32_t a = tile[(2 * lane) * RS + k], b = tile[(2 * lane + 1) * RS + k];
                out[size_t(k) * 32] = a | (b << 16);
            }
        }
    } else {
        // 32 pairs x 256 symbols leave as (256 * ppw)-word runs, one run per warp block touched by this tile (ppw is a power of two)
        const uint32_t lp = 31u - uint32_t(__clz(ppw)), run_shift = 8u + lp;
#pragma unroll 4
        for (uint32_t l = tid; l < 32u * INGEST_TILE; l += INGEST_TILE) {
            const uint32_t wb = l >> run_shift, rem = l & ((1u << run_shift) - 1u), k = rem >> lp, slot = rem & (ppw - 1u);
            if (k < n_valid) {
                const uint32_t pr = (wb << lp) + slot;               // pair inside the tile
                const uint32_t a = tile[(2 * pr) * RS + k], b = tile[(2 * pr + 1) * RS + k];
                const size_t wblk = size_t(blk) * (32u >> lp) + wb;
                p.pk[((wblk * p.n_sym + e0 + k) << lp) + slot] = a | (b << 16);
            }
        }
    }
}

}  // namespace vitb
