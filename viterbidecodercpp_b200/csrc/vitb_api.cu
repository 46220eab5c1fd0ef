// vitb_api.cu -- extern "C" boundary (include/viterbi_b200.h) and host-side orchestration of the decode pipeline
//     ingest (layout + depuncture) -> ACS (add-compare-select, decision rows to HBM) -> traceback -> result gather
// No CPU fallback: every entry point needs a CUDA device.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/viterbi_b200.h"
#include "vitb_registry.h"
#include "ingest.cuh"
#include "traceback.cuh"
#include "frontend.cuh"
#include <cmath>

namespace vitb {

static const std::vector<KernelEntry>& registry() {
    static std::vector<KernelEntry> entries;
    static std::once_flag once;
    std::call_once(once, [] {
        register_small(entries);
        register_k7r2_t1(entries); register_k7r2_t2(entries); register_k7r2_t4(entries);
        register_k7r3_t1(entries); register_k7r3_t4(entries);
        register_k7r4_t1(entries); register_k7r4_t2(entries); register_k7r4_t4(entries);
        register_k9r2_t8(entries); register_k9r2_t16(entries);
        register_k9r4_t8(entries); register_k9r4_t16(entries);
        register_k15r6_cta512(entries); register_k15r6_cta1024(entries);
        register_generic(entries);
        register_k9_hist_group(entries);
        register_k15_hist_cta(entries);
        register_k7_hist_group(entries);
    });
    return entries;
}

// per-frame results picked out of the workspace: acc_error[f], final_error[f] = metrics[f][end_state of the frame]
__global__ void gather_results_kernel(const uint64_t* acc, const uint16_t* metrics, uint32_t n_states, uint32_t end_state,
                                      const uint32_t* end_states, uint32_t n_frames, uint64_t* acc_out, uint32_t* final_out) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    if (acc_out) acc_out[f] = acc[f];
    if (final_out) final_out[f] = metrics[size_t(f) * n_states + (end_states ? end_states[f] : end_state)];
}

// VITB_END_STATE_BEST: the end state of a frame is the state with the smallest final path metric (lowest index on a tie): what a
// caller of the reference does with get_error(s) over all s before chainback(..., s) when the encoder was not flushed to state 0.
// One warp per frame.
__global__ void __launch_bounds__(128) best_state_kernel(const uint16_t* metrics, uint32_t n_states, uint32_t n_frames, uint32_t* best) {
    const uint32_t f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    const uint16_t* m = metrics + size_t(f) * n_states;
    uint32_t key = 0xffffffffu;                         // metric << 16 | state for n_states <= 65536: min(key) = lowest metric, lowest state
    for (uint32_t s = lane; s < n_states; s += 32) {
        const uint32_t k = (uint32_t(m[s]) << 16) | (s & 0xffffu);
        key = k < key ? k : key;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, key, d);
        key = o < key ? o : key;
    }
    if (lane == 0) best[f] = key & 0xffffu;
}

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; bytes = 0; }
        const cudaError_t e = cudaMalloc(&ptr, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
};

// Pinned host memory mapped into the device's address space: the streaming calls stage their (small) input there and take their
// result from there, so that a call is two kernel launches and one synchronize - no copy calls.
struct MappedBuffer {
    void* host = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        release();
        cudaError_t e = cudaHostAlloc(&host, need, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) { host = nullptr; return e; }
        e = cudaHostGetDevicePointer(&dev, host, 0);
        if (e != cudaSuccess) { cudaFreeHost(host); host = nullptr; return e; }
        bytes = need;
        return cudaSuccess;
    }
    void release() { if (host) cudaFreeHost(host); host = nullptr; dev = nullptr; bytes = 0; }
};

// Workspace of one chunk of a batch call.  acs_done: recorded on the caller's stream behind the ACS kernel (the traceback on
// tb_stream waits for it); tb_done: recorded on tb_stream behind the chunk's last kernel / copy (the next ACS that reuses the slot,
// and whoever wants the results, wait for it).
struct BatchSlot {
    DeviceBuffer dec, metrics, acc, tb_spec, tb_fin, end_states, pk;
    cudaEvent_t acs_done = nullptr, tb_done = nullptr, in_ready = nullptr;
    bool tb_pending = false;
};

}  // namespace vitb

using namespace vitb;

struct vitb_decoder {
    vitb_params prm{};
    std::vector<const KernelEntry*> variants;   // every compiled lanes-per-pair variant of this code/config, ascending logt
    const KernelEntry* entry = nullptr;         // variant used by the single-frame streaming API (fixes its decision layout)
    const KernelEntry* hg_entry = nullptr;      // frame-over-4-lanes (K = 9, K = 7) / frame-per-CTA (K = 15) survivor-history kernel, uint16_t metrics, batch calls only
    const KernelEntry* last_batch = nullptr;    // variant the last batch call selected
    bool last_batch_hist = false;               // ... and whether it ran as the survivor-history kernel (acs_hist.cuh)
    std::string name_buf, name_override;        // name_override: a batch call that used two kernels (cta_wave_split)
    int forced_variant = 0;                     // vitb_set_variant (0 = automatic)
    bool use_hist = getenv("VITB_NO_HIST") == nullptr;   // vitb_set_history_kernel
    int seg_records_forced = 0, seg_overlap_forced = -1;   // vitb_set_traceback_segments (0 / -1 = automatic)
    GenericCode gcode{};                        // branch patterns for the generic kernels (codes outside the compiled catalogue)
    int n_sm = 148;
    int n_states = 0;
    int sh = 0;
    int last_cuda = 0;
    uint64_t launches = 0;
    size_t ws_limit = 0;                 // vitb_set_workspace_limit; 0 = ws_default
    size_t ws_default = size_t(24) << 30;   // queried once at create time
    cudaStream_t stream = nullptr;    // owned; used by the host-pointer entry points
    cudaStream_t copy_stream = nullptr;   // owned; host->device copies of the pipelined host-pointer path
    std::vector<cudaEvent_t> copy_ev;     // one per pipeline chunk + fork/join
    // Every batch call reuses the handle's workspace (pk, dec, metrics, ...), so calls on one handle are ordered on the device even
    // when they are enqueued on different streams: each call's stream first waits for `batch_done` of the previous call.
    cudaEvent_t batch_done = nullptr;
    bool batch_pending = false;
    // batch workspace.  Everything the traceback / gather of a chunk reads lives in one of two SLOTS, so that the traceback of chunk
    // c (tb_stream; memory bound) runs next to the add-compare-select of chunk c+1 (caller's stream; issue bound): see BatchSlot
    DeviceBuffer pk, d_in, d_out, d_accout, d_finout, map, end_states;
    BatchSlot slots[2];
    int slot_next = 0;
    cudaStream_t tb_stream = nullptr;     // owned; traceback, best-state and gather kernels of the batch calls
    cudaStream_t acs_stream = nullptr;    // owned; ACS kernels of pipelined calls (the caller's stream keeps only the ingest pass)
    bool pipelined = false;               // vitb_set_pipelining: the join with the last chunk's traceback is deferred to the next call / flush
    size_t window_bits = 0;               // vitb_set_traceback_window (K = 15): decision rows kept per frame = 2 windows; 0 = all
    DeviceBuffer win_count;               // mismatch counter of the sliding-window calls
    bool overlap_hint = false;            // this chunk's traceback will run next to another chunk's ACS (pipelined call, multi-chunk call)
    size_t n_depunctured = 0, n_received = 0;
    int32_t unpunctured_value = 0;
    // single-frame streaming state (one 64-frame block, frame 0 is the user's)
    DeviceBuffer s_pk, s_dec, s_metrics, s_acc, s_in, s_out;
    MappedBuffer s_map;                 // streaming calls: [0, 512) accumulated minima of the call, [512, ...) the call's symbols
    DeviceBuffer g_tx, g_sym, g_cnt;     // host-pointer conveniences of the device-side front end
    size_t traceback_length = 0;
    size_t current_decoded_bit = 0;
    // stage profiling
    bool profiling = false;
    std::vector<cudaEvent_t> ev;      // 5 events per chunk of the last batch call
    size_t ev_used = 0;
};

namespace {

inline uint32_t pack2(uint32_t v) { return (v & 0xffffu) | (v << 16); }

int cuda_fail(vitb_decoder* h, cudaError_t e) {
    if (h) h->last_cuda = int(e);
    return VITB_ERR_CUDA;
}
#define VITB_CUDA(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail((h), e__); } while (0)

std::vector<const KernelEntry*> find_entries(const vitb_params& p) {
    std::vector<const KernelEntry*> out;
    if (p.K < 2 || p.R < 1 || p.R > VITB_MAX_R) return out;
    const int sh = (p.soft_bytes == 1) ? 8 : 0;
    const uint32_t span = uint32_t(p.soft_decision_high - p.soft_decision_low);
    const uint32_t mask = (p.soft_bytes == 1) ? 0xffu : 0xffffu;
    const int consistent = ((p.soft_decision_max_error & mask) == ((span * uint32_t(p.R)) & mask)) ? 1 : 0;
    const uint32_t kmask = (p.K >= 32) ? 0xffffffffu : ((1u << p.K) - 1u);
    for (const KernelEntry& e : registry()) {
        if (e.generic || e.layout == LAYOUT_HISTGROUP || e.layout == LAYOUT_HISTCTA || e.K != p.K || e.R != p.R || e.sh != sh || e.tie != p.tie_break || e.consistent != consistent) continue;
        bool same = true;
        for (int i = 0; i < p.R; i++) same = same && ((e.G[i] & kmask) == (p.G[i] & kmask));
        if (same) out.push_back(&e);
    }
    if (out.empty() && p.R <= GENERIC_MAX_R) {      // not in the compiled catalogue: the generic kernel of this K, if there is one
        for (const KernelEntry& e : registry())
            if (e.generic && e.K == p.K && e.sh == sh && e.tie == p.tie_break) out.push_back(&e);
    }
    std::sort(out.begin(), out.end(), [](const KernelEntry* a, const KernelEntry* b) { return a->logt < b->logt; });
    return out;
}

// the batch-only survivor-history kernel of this code, if it has one: frame over 4 lanes (K = 9) or frame per CTA (K = 15), uint16_t
// metrics
const KernelEntry* find_hist_group(const vitb_params& p) {
    if (p.soft_bytes != 2) return nullptr;
    const uint32_t span = uint32_t(p.soft_decision_high - p.soft_decision_low);
    const int consistent = ((p.soft_decision_max_error & 0xffffu) == ((span * uint32_t(p.R)) & 0xffffu)) ? 1 : 0;
    const uint32_t kmask = (p.K >= 32) ? 0xffffffffu : ((1u << p.K) - 1u);
    for (const KernelEntry& e : registry()) {
        if ((e.layout != LAYOUT_HISTGROUP && e.layout != LAYOUT_HISTCTA) || e.K != p.K || e.R != p.R || e.tie != p.tie_break || e.consistent != consistent) continue;
        bool same = true;
        for (int i = 0; i < p.R; i++) same = same && ((e.G[i] & kmask) == (p.G[i] & kmask));
        if (same) return &e;
    }
    return nullptr;
}

// Variant for a batch of n_frames.  Measured on B200 (profiles/r01_summary.md): fewer lanes per pair = fewer instructions per ACS,
// but every SM sub-partition needs a few warps to overlap the 2-cycle dispatch of the packed-integer instructions.  Pick the
// fewest lanes that still give `want` warps per sub-partition (1.5 for the one-thread-per-pair kernel, whose warps carry 32
// independent butterflies each; 4 for the lane-group kernels), else the most parallel variant.
const KernelEntry* choose_variant(const vitb_decoder* h, size_t n_frames) {
    if (h->forced_variant > 0) {
        for (const KernelEntry* e : h->variants) if (entry_variant(e) == h->forced_variant) return e;
    }
    const size_t pairs = (n_frames + 1) / 2, smsp = size_t(h->n_sm) * 4;
    if (h->variants.front()->layout == LAYOUT_CTA) return h->variants.front();  // K = 15: 512 threads x 32 registers measured fastest (56.0 vs 61.1 ms)
    // survivor-history kernel (acs_hist.cuh, 4 instructions per butterfly): about twice as fast per add-compare-select as the
    // predicate kernels and no ingest pass, so it wins unless the batch is so small that only the lane-group variants can spread
    // it over the GPU (a warp of the history kernel runs alone at full speed: 16 384 config-2 frames = 256 warps take 0.34 ms)
    const KernelEntry* e0 = h->variants.front();
    if (h->use_hist && e0->logt == 0 && e0->launch_hist) {
        const size_t fpw = (e0->sh == 8) ? 64 : 32;
        if ((n_frames + fpw - 1) / fpw >= smsp / 8) return e0;
    }
    for (const KernelEntry* e : h->variants) {
        const size_t warps = ((pairs << e->logt) + 31) / 32;
        const size_t want = (e->logt == 0) ? smsp * 3 / 2 : smsp * 4;
        if (e->layout == LAYOUT_CTA || warps >= want) return e;
    }
    return h->variants.back();
}

bool params_valid(const vitb_params& p) {
    if (p.K < 2 || p.K > 24 || p.R < 1 || p.R > VITB_MAX_R) return false;
    if (p.soft_bytes != 1 && p.soft_bytes != 2) return false;
    if (p.soft_decision_high <= p.soft_decision_low) return false;        // viterbi_branch_table.h:43
    if (p.tie_break < VITB_TIE_SCALAR || p.tie_break > VITB_TIE_SIMD_SAT) return false;
    const int lim = (p.soft_bytes == 1) ? 127 : 32767;
    if (p.soft_decision_high > lim || p.soft_decision_low < -lim - 1) return false;
    return true;
}

void fill_acs_params(const vitb_decoder* h, AcsParams& a) {
    const vitb_params& p = h->prm;
    const int sh = h->sh;
    const uint32_t emask = (p.soft_bytes == 1) ? 0xffu : 0xffffu;
    a.c_low2 = pack2(uint32_t(-p.soft_decision_low) << sh);
    a.c_high2 = pack2((uint32_t(p.soft_decision_high) << sh) + 1u);
    const uint32_t span = uint32_t(p.soft_decision_high - p.soft_decision_low);
    a.c_inv2 = pack2(((p.soft_decision_max_error - span * uint32_t(p.R)) & emask) << sh);
    a.max_err2 = pack2((p.soft_decision_max_error & emask) << sh);
    a.thr2 = pack2((p.renormalisation_threshold & emask) << sh);
    a.init_start2 = pack2((p.initial_start_error & emask) << sh);
    a.init_other2 = pack2((p.initial_non_start_error & emask) << sh);
}

template <typename soft_t, int SH>
cudaError_t launch_ingest(const IngestParams& ip, unsigned n_blocks, cudaStream_t s) {
    dim3 grid((ip.n_sym + INGEST_TILE - 1) / INGEST_TILE, n_blocks);
    ingest_pairs_kernel<soft_t, SH><<<grid, INGEST_TILE, 0, s>>>(ip);
    return cudaGetLastError();
}

cudaError_t run_ingest(vitb_decoder* h, const IngestParams& ip, unsigned n_blocks, cudaStream_t s) {   // n_blocks: 64-frame tiles
    h->launches++;
    return (h->prm.soft_bytes == 1) ? launch_ingest<int8_t, 8>(ip, n_blocks, s) : launch_ingest<int16_t, 0>(ip, n_blocks, s);
}

size_t default_ws_limit() {
    if (const char* e = getenv("VITB_WORKSPACE_MB")) {
        const long long mb = atoll(e);
        if (mb > 0) return size_t(mb) << 20;
    }
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) return total_b / 10 * 6;   // 60 % of HBM (108 GB of 180 GB)
    cudaGetLastError();
    return size_t(24) << 30;
}

// workspace bytes per 64 frames for a frame of S steps (identical for every variant: decision rows are 2^(K-1) bits per frame-step)
size_t block_bytes(const vitb_decoder* h, size_t S, bool with_packed_stream = true, bool two_slots = false) {
    const size_t n_sym = S * size_t(h->prm.R);
    const bool windowed = h->window_bits && h->variants.front()->layout == LAYOUT_CTA && S > 2 * h->window_bits;
    if (windowed) two_slots = false;                         // sliding-window calls use one slot
    const size_t rows = windowed ? 2 * h->window_bits : S;   // sliding window: a ring of two windows of decision rows
    const size_t slot = rows * 64 * (size_t(h->n_states) / 8 < 8 ? 8 : size_t(h->n_states) / 8)   // decision rows
                      + size_t(64) * h->n_states * 2                                           // metrics
                      + 64 * 8;                                                                // accumulated error
    return (with_packed_stream ? n_sym * 32 * 4 : 0)        // packed symbols (not needed when the kernel reads the caller's rows itself)
         + slot * (two_slots ? 2 : 1);                      // consecutive chunks alternate between two slots (BatchSlot)
}

size_t dec_bytes_per_block64(const KernelEntry* e, size_t rows) {
    if (e->layout == LAYOUT_PAIR) return rows * 64 * 8;
    if (e->layout == LAYOUT_CTA) return size_t(32) * rows * (size_t(1) << e->logt) * size_t(e->dec_words) * 4;
    return (size_t(1) << e->logt) * rows * 32 * size_t(e->dec_words) * 4;
}

// bytes of one decision row of ONE warp block / CTA (the unit the streaming state keeps)
size_t dec_row_bytes_unit(const KernelEntry* e) {
    if (e->layout == LAYOUT_PAIR) return 64 * 8;
    if (e->layout == LAYOUT_CTA) return (size_t(1) << e->logt) * size_t(e->dec_words) * 4;
    return size_t(32) * size_t(e->dec_words) * 4;
}

cudaError_t launch_traceback(vitb_decoder* h, const KernelEntry* e, const void* dec, size_t dec_rows, size_t n_frames, size_t L,
                             size_t end_state, uint8_t* d_out, size_t out_stride, cudaStream_t s, const uint32_t* end_states = nullptr,
                             BatchSlot* sl = nullptr) {
    h->launches++;
    if (e->layout == LAYOUT_CTA) {
        TracebackCtaParams t{};
        t.dec = static_cast<const uint32_t*>(dec); t.dec_rows = uint32_t(dec_rows); t.n_frames = uint32_t(n_frames);
        t.total_bits = uint32_t(L); t.state_bits = uint32_t(h->prm.K - 1); t.logt = uint32_t(e->logt); t.end_state = uint32_t(end_state); t.end_states = end_states;
        t.out = d_out; t.out_stride = out_stride; t.words = uint32_t(e->dec_words);
        // the chain of a K = 15 frame is L / 5 dependent memory round trips (16 384 bits: 5.7 ms): walk it in concurrent segments
        // (warm-up from a guessed state, verified and repaired afterwards, traceback.cuh); the streaming state keeps the plain walk
        static const bool no_seg = getenv("VITB_NO_SEG_TRACEBACK") != nullptr;
        const size_t overlap = h->seg_overlap_forced >= 0 ? size_t(h->seg_overlap_forced) * 8 : size_t(16) * size_t(h->prm.K);     // rows
        size_t seg_bits = h->seg_records_forced > 0 ? size_t(h->seg_records_forced) * 8 : ((4 * overlap + 7) / 8 * 8);
        if (seg_bits < 8) seg_bits = 8;
        if ((L + seg_bits - 1) / seg_bits > 65535) seg_bits = ((L + 65534) / 65535 + 7) / 8 * 8;                    // gridDim.y
        const size_t n_seg = (no_seg || !sl) ? 1 : (L + seg_bits - 1) / seg_bits;      // no slot: the streaming state keeps the plain walk
        if (n_seg <= 1) {
            traceback_cta_kernel<5><<<unsigned((n_frames + 63) / 64), 64, 0, s>>>(t);
        } else {
            cudaError_t ce = sl->tb_spec.reserve(n_seg * n_frames * 4);
            if (ce == cudaSuccess) ce = sl->tb_fin.reserve(n_seg * n_frames * 4);
            if (ce != cudaSuccess) return ce;
            TracebackCtaSegParams sp{};
            sp.n_seg = uint32_t(n_seg); sp.seg_bits = uint32_t(seg_bits); sp.overlap = uint32_t(overlap);
            sp.spec = static_cast<uint32_t*>(sl->tb_spec.ptr); sp.fin = static_cast<uint32_t*>(sl->tb_fin.ptr);
            traceback_cta_seg_kernel<5><<<dim3(unsigned((n_frames + 63) / 64), unsigned(n_seg)), 64, 0, s>>>(t, sp);
            ce = cudaGetLastError();
            if (ce != cudaSuccess) return ce;
            h->launches++;
            traceback_cta_fix_kernel<5><<<unsigned((n_frames + 63) / 64), 64, 0, s>>>(t, sp);
        }
    } else if (e->layout == LAYOUT_PAIR) {
        TracebackParams t{};
        t.dec = static_cast<const uint64_t*>(dec); t.dec_rows = uint32_t(dec_rows); t.n_frames = uint32_t(n_frames);
        t.total_bits = uint32_t(L); t.state_bits = uint32_t(h->prm.K - 1); t.end_state = uint32_t(end_state); t.end_states = end_states;
        t.out = d_out; t.out_stride = out_stride; t.tag_layout = e->tagged_rows ? 1u : 0u;     // uint8_t catalogue pair kernels use the tagged butterfly
        traceback_u64_kernel<32><<<unsigned((n_frames + 127) / 128), 128, 0, s>>>(t);
    } else {
        TracebackGroupParams t{};
        t.dec = static_cast<const uint32_t*>(dec); t.dec_rows = uint32_t(dec_rows); t.n_frames = uint32_t(n_frames);
        t.total_bits = uint32_t(L); t.state_bits = uint32_t(h->prm.K - 1); t.logt = uint32_t(e->logt); t.end_state = uint32_t(end_state); t.end_states = end_states;
        t.out = d_out; t.out_stride = out_stride;
        // staged kernel: 32 frames per warp; needs the decision buffer padded to whole 64-frame blocks (it is) and PPW <= 16
        // two batches of ROWS rows are in flight per warp; with few warps per SM (large frames, small batches) 8 rows leave DRAM
        // under-subscribed, so 16 are used whenever the shared memory of the 4-warp CTA stays below 160 KB
        const size_t row_words = size_t(16) * (size_t(1) << e->logt) * size_t(e->dec_words);
        const bool deep = size_t(4) * 2 * 16 * row_words * 4 <= size_t(160) * 1024;
        const size_t smem = size_t(4) * 2 * (deep ? 16 : 8) * row_words * 4;
        const unsigned grid = unsigned((n_frames + 127) / 128);
        static const bool force_shuffle = getenv("VITB_TRACEBACK_SHUFFLE") != nullptr;
        // the streaming state holds a single warp block (not a padded 64-frame block): only the shuffle kernel stays inside it
        const bool use_shuffle = force_shuffle || dec == h->s_dec.ptr;
        auto staged = [&](auto kernel) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            kernel<<<grid, 128, smem, s>>>(t);
        };
        if (use_shuffle) {
            const size_t ppw = size_t(32) >> e->logt, n_pairs = (n_frames + 1) / 2, n_wblocks = (n_pairs + ppw - 1) / ppw;
            const unsigned g2 = unsigned((n_wblocks + 3) / 4);
            if (e->dec_words == 1) traceback_group_kernel<1, 16><<<g2, 128, 0, s>>>(t);
            else traceback_group_kernel<2, 16><<<g2, 128, 0, s>>>(t);
        } else if (e->dec_words == 1) {
            if (deep) staged(traceback_group_staged_kernel<1, 16>); else staged(traceback_group_staged_kernel<1, 8>);
        } else {
            if (deep) staged(traceback_group_staged_kernel<2, 16>); else staged(traceback_group_staged_kernel<2, 8>);
        }
    }
    return cudaGetLastError();
}

cudaError_t mark(vitb_decoder* h, cudaStream_t s) {
    if (!h->profiling) return cudaSuccess;
    if (h->ev_used == h->ev.size()) {
        cudaEvent_t e;
        const cudaError_t r = cudaEventCreate(&e);
        if (r != cudaSuccess) return r;
        h->ev.push_back(e);
    }
    return cudaEventRecord(h->ev[h->ev_used++], s);
}

// can the ACS kernel `e` read the caller's rows where they lie (no ingest pass, no packed stream)?
bool direct_fetch_ok(const vitb_decoder* h, const KernelEntry* e, const void* d_symbols, size_t row_stride) {
    static const bool no_direct = getenv("VITB_NO_DIRECT") != nullptr;
    const size_t row_bytes = row_stride * size_t(h->prm.soft_bytes);
    const bool hist = h->use_hist && e->launch_hist != nullptr;
    return !no_direct && (hist || e->launch_direct) && !h->n_depunctured && (row_bytes % 4 == 0) && row_bytes >= 4 &&
           (reinterpret_cast<uintptr_t>(d_symbols) % 4 == 0);
}

// order this call after the previous batch call of the handle (which may still be running on another stream), and publish its end
cudaError_t batch_begin(vitb_decoder* h, cudaStream_t s) {
    if (!h->batch_done) {
        const cudaError_t e = cudaEventCreateWithFlags(&h->batch_done, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return h->batch_pending ? cudaStreamWaitEvent(s, h->batch_done, 0) : cudaSuccess;
}
cudaError_t batch_end(vitb_decoder* h, cudaStream_t s) {
    h->batch_pending = true;
    return cudaEventRecord(h->batch_done, s);
}

// Slots and streams of the batch calls.  A chunk's ingest and ACS kernels run on the caller's stream `s`; its best-state, traceback,
// gather kernels (and, for the host-pointer path, its device-to-host copies) run on the handle's tb_stream behind the slot's
// acs_done event.  The traceback is bound by DRAM round trips and the ACS by instruction issue, so the two overlap almost for free:
// measured on config 2, a step is 0.83 ms with the stages in sequence and ~0.66 ms with the traceback of batch i next to the ACS of
// batch i+1 (profiles/r02_summary.md).  join_slots makes `s` wait for the results.
cudaError_t slot_events(BatchSlot& sl) {
    if (!sl.acs_done) {
        cudaError_t e = cudaEventCreateWithFlags(&sl.acs_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.tb_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.in_ready, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// `s` waits for the results of every slot (keep_newest: except the slot used last - a pipelined call leaves that one in flight).
// tb_pending stays set: a later call on ANOTHER stream has to wait again, and waiting for a completed event costs nothing.
cudaError_t join_slots(vitb_decoder* h, cudaStream_t s, bool keep_newest) {
    const int newest = h->slot_next ^ 1;
    for (int i = 0; i < 2; i++) {
        BatchSlot& sl = h->slots[i];
        if (!sl.tb_pending || (keep_newest && i == newest)) continue;
        const cudaError_t e = cudaStreamWaitEvent(s, sl.tb_done, 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// One chunk of frames, everything on device, asynchronous: ACS on `s`, traceback and result gather on h->tb_stream (joined by the
// caller).  d2h_*: optional host destinations of this chunk's results, copied on tb_stream behind the gather.
int decode_chunk_dev(vitb_decoder* h, const KernelEntry* e, const void* d_symbols, size_t row_stride, size_t n_frames, size_t L,
                     size_t start_state, size_t end_state, uint8_t* d_out, uint64_t* d_acc, uint32_t* d_final, cudaStream_t s,
                     uint8_t* h_out = nullptr, uint64_t* h_acc = nullptr, uint32_t* h_final = nullptr) {
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), S = L + K - 1, n_sym = S * R;
    const unsigned n_b64 = unsigned((n_frames + 63) / 64);
    BatchSlot& sl = h->slots[h->slot_next];
    h->slot_next ^= 1;
    VITB_CUDA(h, slot_events(sl));
    if (sl.tb_pending) VITB_CUDA(h, cudaStreamWaitEvent(s, sl.tb_done, 0));     // the slot's buffers are free once its last traceback is done
    cudaStream_t ts = h->tb_stream;
    // one-thread-per-pair entries: survivor-history records (acs_hist.cuh) instead of decision rows; same bytes per frame and step.
    // uint8_t metrics: warp block = 64 frames, 8-step records; uint16_t metrics: warp block = 32 frames, 16-step records.
    // K = 9 with uint16_t metrics: the frame-over-4-lanes history kernel (acs_hist_group.cuh) when the symbols can be fetched directly
    // and the batch fills the GPU with 8 frames per warp, or when that variant was pinned with vitb_set_variant(h, 4)
    // K = 15 with uint16_t metrics: the frame-per-CTA history kernel (acs_hist_cta.cuh) for unpunctured input (any alignment)
    const size_t row_bytes0 = row_stride * size_t(h->prm.soft_bytes);
    bool hg = false, hc = false;
    if (h->use_hist && h->hg_entry && getenv("VITB_NO_DIRECT") == nullptr && !h->n_depunctured) {
        const KernelEntry* a = h->hg_entry;
        const bool forced = h->forced_variant == entry_variant(a);
        if (a->layout == LAYOUT_HISTGROUP) {
            const bool direct_ok = (row_bytes0 % 4 == 0) && row_bytes0 >= 4 && (reinterpret_cast<uintptr_t>(d_symbols) % 4 == 0);
            // K = 9: the default once its 8 frames per warp fill the GPU.  K = 7: the one-lane history kernel (32 frames per warp,
            // 16 instructions per add-compare-select pair fewer) is the faster one as soon as every scheduler has a warp of it; below
            // that a frame over 4 lanes brings four times the warps (run_simple, 8192 frames: 256 warps on 592 schedulers)
            // (measured, B200, ACS only: Voyager soft16 4096 frames 0.173 -> 0.106 ms, 8192 frames 0.175 -> 0.145 ms, 12 288 frames
            // 0.175 -> 0.182 ms; DAB R = 1/4, whose 16-entry branch metric table is paid per lane: 8192 frames 1.53 -> 1.60 ms, 4096 frames
            // 1.53 -> 1.49 ms - pinned only)
            const size_t one_lane_warps = (n_frames + 31) / 32;
            const bool auto_ok = (h->prm.K == 9) ? ((n_frames + 7) / 8 >= size_t(h->n_sm) * 2)
                                                 : (h->prm.R <= 3 && one_lane_warps < size_t(h->n_sm) * 2);
            hg = direct_ok && (forced || (h->forced_variant == 0 && auto_ok));
        } else {
            // One frame per CTA: 1.5 times the instructions per frame of the packed decision-row kernel (acs_cta.cuh, a frame PAIR per
            // CTA: 14.0 ms per wave of 148 pairs against 10.4 ms per wave of 148 frames, config 5), but half the granularity: it takes
            // the batches (and, see cta_wave_split, the last partial wave of a large batch) that fit one wave of single-frame CTAs
            const bool direct_ok = (row_bytes0 % 2 == 0) && (reinterpret_cast<uintptr_t>(d_symbols) % 2 == 0);
            hc = direct_ok && (forced || (h->forced_variant == 0 && n_frames <= size_t(h->n_sm)));
        }
        if (hg || hc) { e = a; h->last_batch = e; }
    }
    const bool hist = h->use_hist && e->launch_hist != nullptr;
    const bool hist_wide = hist && e->sh == 0;
    h->last_batch_hist = hist;
    const size_t hist_bits = hist_wide ? 16 : 8, n_periods = (S + hist_bits - 1) / hist_bits;
    const unsigned ppw = hist_wide ? 16u : unsigned(e->ppw);
    const unsigned n_wblocks = hc ? unsigned(n_frames) : (hg ? unsigned((n_frames + 7) / 8) : n_b64 * (32u / ppw));
    // record bytes per block and period: a warp block of the one-lane kernels holds 64 states-bytes x 64 (or 32 x 2) frames, the
    // lane-group kernel 8 frames, the CTA kernel one frame: always 2^(K-1) bits per frame and step
    const size_t rec_block_bytes = hc ? size_t(2) * size_t(h->n_states) : ((64 * size_t(h->n_states)) >> (hg ? e->logt : 0));
    VITB_CUDA(h, sl.dec.reserve(hist ? size_t(n_wblocks) * n_periods * rec_block_bytes : size_t(n_b64) * dec_bytes_per_block64(e, S)));
    VITB_CUDA(h, sl.metrics.reserve(size_t(n_b64) * 64 * h->n_states * 2));
    VITB_CUDA(h, sl.acc.reserve(size_t(n_b64) * 64 * 8));

    // one-thread-per-pair kernels can read the caller's rows themselves when they are 4-byte aligned and not punctured
    const size_t row_bytes = row_stride * size_t(h->prm.soft_bytes);
    const bool direct = hc || hg || direct_fetch_ok(h, e, d_symbols, row_stride);
    VITB_CUDA(h, mark(h, s));
    if (!direct) {
        VITB_CUDA(h, sl.pk.reserve(size_t(n_b64) * n_sym * 32 * 4));      // packed stream (per slot): only the ingest path needs it
        IngestParams ip{};
        ip.symbols = d_symbols; ip.row_stride = row_stride; ip.n_frames = uint32_t(n_frames); ip.n_sym = uint32_t(n_sym);
        ip.depuncture_map = h->n_depunctured ? static_cast<const int32_t*>(h->map.ptr) : nullptr;
        ip.fill_value = h->unpunctured_value; ip.pk = static_cast<uint32_t*>(sl.pk.ptr); ip.ppw = ppw;
        VITB_CUDA(h, run_ingest(h, ip, n_b64, s));
    }
    VITB_CUDA(h, mark(h, s));
    // Pipelined calls run the ACS kernel on the handle's own stream, behind an event of the caller's stream (inputs and ingest
    // done, slot free): the caller's stream is then free to run the NEXT call's ingest pass (memory bound, into the other slot's
    // packed stream) next to this ACS kernel - a three-stage pipeline ingest | ACS | traceback (config 4: 0.66 -> 0.5 ms per batch)
    cudaStream_t as = h->pipelined ? h->acs_stream : s;
    if (h->pipelined) {
        VITB_CUDA(h, cudaEventRecord(sl.in_ready, s));
        VITB_CUDA(h, cudaStreamWaitEvent(as, sl.in_ready, 0));
    }

    AcsParams a{};
    fill_acs_params(h, a);
    a.sym = d_symbols; a.sym_row_bytes = row_bytes; a.sym_total_bytes = row_bytes * n_frames; a.n_frames = uint32_t(n_frames);
    a.pk = static_cast<const uint32_t*>(sl.pk.ptr); a.dec = sl.dec.ptr;
    a.metrics = static_cast<uint16_t*>(sl.metrics.ptr); a.acc = static_cast<uint64_t*>(sl.acc.ptr);
    a.n_blocks = n_wblocks; a.n_steps = uint32_t(S); a.dec_rows = uint32_t(S); a.dec_row0 = 0; a.resume = 0; a.start_state = uint32_t(start_state);
    h->launches++;
    if (hist) VITB_CUDA(h, (direct && !hg && !hc) ? e->launch_hist_direct(a, as) : e->launch_hist(a, as));    // the hist-group / hist-CTA launchers always fetch directly
    else if (e->generic) VITB_CUDA(h, e->launch_generic(a, h->gcode, as));
    else VITB_CUDA(h, direct ? e->launch_direct(a, as) : e->launch(a, as));
    VITB_CUDA(h, mark(h, as));
    VITB_CUDA(h, cudaEventRecord(sl.acs_done, as));
    VITB_CUDA(h, cudaStreamWaitEvent(ts, sl.acs_done, 0));

    const uint32_t* end_states = nullptr;
    if (end_state == VITB_END_STATE_BEST) {
        VITB_CUDA(h, sl.end_states.reserve(n_frames * 4));
        h->launches++;
        best_state_kernel<<<unsigned((n_frames + 3) / 4), 128, 0, ts>>>(a.metrics, uint32_t(h->n_states), uint32_t(n_frames),
                                                                         static_cast<uint32_t*>(sl.end_states.ptr));
        VITB_CUDA(h, cudaGetLastError());
        end_states = static_cast<const uint32_t*>(sl.end_states.ptr);
        end_state = 0;
    }
    if (d_out && hist) {
        h->launches++;
        TracebackHistParams t{};
        t.dec = static_cast<const uint8_t*>(sl.dec.ptr); t.n_periods = uint32_t(n_periods); t.n_frames = uint32_t(n_frames);
        t.total_bits = uint32_t(L); t.state_bits = uint32_t(K - 1); t.end_state = uint32_t(end_state); t.end_states = end_states; t.n_steps = uint32_t(S);
        t.hist_bits = uint32_t(hist_bits); t.logt = (hg || hc) ? uint32_t(e->logt) : 0u; t.out = d_out; t.out_stride = (L + 7) / 8;
        // A frame's chain is n_periods dependent memory round trips; small batches cannot hide them, so the chain is cut into
        // segments walked concurrently (warm-up over `overlap` records from a guessed state, verified and repaired afterwards:
        // traceback.cuh).  About 128 K threads saturate DRAM; the warm-up is kept below a quarter of the walk.
        static const bool no_seg = getenv("VITB_NO_SEG_TRACEBACK") != nullptr;
        // warm-up: 96 steps for K <= 9, 16 K steps for longer codes (the depth the K = 15 row walk uses)
        const size_t warm_steps = K <= 9 ? 96 : 16 * K;
        const size_t overlap = h->seg_overlap_forced >= 0 ? size_t(h->seg_overlap_forced) : (warm_steps + hist_bits - 1) / hist_bits;
        static const size_t seg_target = getenv("VITB_SEG_TARGET") ? size_t(atoll(getenv("VITB_SEG_TARGET"))) : size_t(131072);
        // A traceback that runs NEXT TO the ACS kernel of another chunk only has to finish before that kernel does, and a single chain
        // per frame always does (n_periods round trips of ~0.63 us against >= 150 clocks per step of the ACS); walking concurrent
        // segments would only take DRAM bandwidth and latency away from the ACS (config 2, pipelined: 0.749 ms per batch with two
        // segments per frame, 0.677 ms with one; profiles/r02_summary.md)
        // (Only behind the one-lane kernels, whose few CTAs leave room on every SM: the K = 9 kernel fills the register files, the
        // traceback then runs in its tail anyway and is better off segmented - config 3 pipelined: 4.83 ms single chain, 4.64 segmented.)
        // (The K = 7 lane-group kernel is a small-batch kernel: its CTAs leave room too.)
        const bool gentle = h->overlap_hint && !(hg && K == 9) && !hc;
        const size_t want_seg = gentle ? 1 : (seg_target + n_frames - 1) / n_frames;
        size_t seg_records = (n_periods + want_seg - 1) / want_seg;
        if (seg_records < 4 * overlap && !gentle) seg_records = 4 * overlap;
        if (h->seg_records_forced > 0) seg_records = size_t(h->seg_records_forced);
        if (seg_records == 0) seg_records = 1;
        if ((n_periods + seg_records - 1) / seg_records > 65535) seg_records = (n_periods + 65534) / 65535;      // gridDim.y
        const size_t n_seg = no_seg ? 1 : (n_periods + seg_records - 1) / seg_records;
        if (n_seg <= 1) {
            traceback_hist_kernel<<<unsigned((n_frames + 127) / 128), 128, 0, ts>>>(t);
        } else {
            VITB_CUDA(h, sl.tb_spec.reserve(n_seg * n_frames * 4));
            VITB_CUDA(h, sl.tb_fin.reserve(n_seg * n_frames * 4));
            TracebackSegParams sp{};
            sp.n_seg = uint32_t(n_seg); sp.seg_records = uint32_t(seg_records); sp.overlap = uint32_t(overlap);
            sp.spec = static_cast<uint32_t*>(sl.tb_spec.ptr); sp.fin = static_cast<uint32_t*>(sl.tb_fin.ptr);
            traceback_hist_seg_kernel<<<dim3(unsigned((n_frames + 127) / 128), unsigned(n_seg)), 128, 0, ts>>>(t, sp);
            VITB_CUDA(h, cudaGetLastError());
            h->launches++;
            traceback_hist_fix_kernel<<<unsigned((n_frames + 127) / 128), 128, 0, ts>>>(t, sp);
        }
        VITB_CUDA(h, cudaGetLastError());
    } else if (d_out) VITB_CUDA(h, launch_traceback(h, e, sl.dec.ptr, S, n_frames, L, end_state, d_out, (L + 7) / 8, ts, end_states, &sl));
    VITB_CUDA(h, mark(h, ts));
    if (d_acc || d_final) {
        h->launches++;
        gather_results_kernel<<<unsigned((n_frames + 255) / 256), 256, 0, ts>>>(a.acc, a.metrics, uint32_t(h->n_states), uint32_t(end_state),
                                                                               end_states, uint32_t(n_frames), d_acc, d_final);
        VITB_CUDA(h, cudaGetLastError());
    }
    VITB_CUDA(h, mark(h, ts));
    const size_t out_stride = (L + 7) / 8;
    if (h_out && d_out) VITB_CUDA(h, cudaMemcpyAsync(h_out, d_out, n_frames * out_stride, cudaMemcpyDeviceToHost, ts));
    if (h_acc && d_acc) VITB_CUDA(h, cudaMemcpyAsync(h_acc, d_acc, n_frames * 8, cudaMemcpyDeviceToHost, ts));
    if (h_final && d_final) VITB_CUDA(h, cudaMemcpyAsync(h_final, d_final, n_frames * 4, cudaMemcpyDeviceToHost, ts));
    VITB_CUDA(h, cudaEventRecord(sl.tb_done, ts));
    sl.tb_pending = true;
    return VITB_OK;
}

// Sliding-window decode of one chunk of frames (K = 15, decision-row kernel; vitb_set_traceback_window).  The reference keeps every
// decision row of a frame until chainback (core.h:180-186: 2 KB per step, 33.6 MB per config-5 frame).  Here the frame is cut
// into windows of W = window_bits steps; the ACS kernel runs one window per launch (resuming from the saved metrics, acs_cta.cuh) into
// a ring of 2 W rows, and as soon as window c+1 is in, the decoded bits of window c are walked out: the walk starts W - (K-1) rows
// above them from state 0, and survivor paths merge long before they reach the bits that are written (the same warm-up as the
// segmented walk of the exact mode, which additionally verifies and repairs; here the rows are gone by then, so mismatches are only
// COUNTED: vitb_get_window_mismatches == 0 certifies that the windowed bytes are the exact ones).  The last window starts from the
// true end state.  Metrics, accumulated errors and final errors are unaffected by the window.  Everything runs on `s`.
int decode_chunk_window(vitb_decoder* h, const KernelEntry* e, const void* d_symbols, size_t row_stride, size_t n_frames, size_t L,
                        size_t start_state, size_t end_state, uint8_t* d_out, uint64_t* d_acc, uint32_t* d_final, cudaStream_t s) {
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), SB = K - 1, S = L + SB, n_sym = S * R, W = h->window_bits, ring = 2 * W;
    const unsigned n_b64 = unsigned((n_frames + 63) / 64);
    BatchSlot& sl = h->slots[0];
    h->slot_next = 1;
    VITB_CUDA(h, slot_events(sl));
    if (sl.tb_pending) VITB_CUDA(h, cudaStreamWaitEvent(s, sl.tb_done, 0));
    h->last_batch = e;
    h->last_batch_hist = false;
    VITB_CUDA(h, sl.dec.reserve(size_t(n_b64) * dec_bytes_per_block64(e, ring)));
    VITB_CUDA(h, sl.metrics.reserve(size_t(n_b64) * 64 * h->n_states * 2));
    VITB_CUDA(h, sl.acc.reserve(size_t(n_b64) * 64 * 8));
    VITB_CUDA(h, h->pk.reserve(size_t(n_b64) * n_sym * 32 * 4));
    VITB_CUDA(h, h->win_count.reserve(8));
    VITB_CUDA(h, mark(h, s));
    IngestParams ip{};
    ip.symbols = d_symbols; ip.row_stride = row_stride; ip.n_frames = uint32_t(n_frames); ip.n_sym = uint32_t(n_sym);
    ip.depuncture_map = h->n_depunctured ? static_cast<const int32_t*>(h->map.ptr) : nullptr;
    ip.fill_value = h->unpunctured_value; ip.pk = static_cast<uint32_t*>(h->pk.ptr); ip.ppw = unsigned(e->ppw);
    VITB_CUDA(h, run_ingest(h, ip, n_b64, s));
    VITB_CUDA(h, mark(h, s));

    const size_t n_seg = (L + W - 1) / W, n_win = (S + W - 1) / W;
    if (d_out) {
        VITB_CUDA(h, sl.tb_spec.reserve(n_seg * n_frames * 4));
        VITB_CUDA(h, sl.tb_fin.reserve(n_seg * n_frames * 4));
    }
    TracebackCtaParams t{};
    t.dec = static_cast<const uint32_t*>(sl.dec.ptr); t.dec_rows = uint32_t(ring); t.ring_rows = uint32_t(ring); t.n_frames = uint32_t(n_frames);
    t.total_bits = uint32_t(L); t.state_bits = uint32_t(SB); t.logt = uint32_t(e->logt); t.end_state = uint32_t(end_state == VITB_END_STATE_BEST ? 0 : end_state);
    t.end_states = nullptr; t.out = d_out; t.out_stride = (L + 7) / 8; t.words = uint32_t(e->dec_words);
    TracebackCtaSegParams sp{};
    sp.n_seg = uint32_t(n_seg); sp.seg_bits = uint32_t(W); sp.overlap = uint32_t(W - SB);
    sp.spec = static_cast<uint32_t*>(sl.tb_spec.ptr); sp.fin = static_cast<uint32_t*>(sl.tb_fin.ptr);
    auto walk_segment = [&](size_t j) -> cudaError_t {
        sp.seg0 = uint32_t(j);
        h->launches++;
        traceback_cta_seg_kernel<5><<<dim3(unsigned((n_frames + 63) / 64), 1), 64, 0, s>>>(t, sp);
        return cudaGetLastError();
    };

    AcsParams a{};
    fill_acs_params(h, a);
    a.sym = d_symbols; a.n_frames = uint32_t(n_frames);
    a.pk = static_cast<const uint32_t*>(h->pk.ptr); a.dec = sl.dec.ptr;
    a.metrics = static_cast<uint16_t*>(sl.metrics.ptr); a.acc = static_cast<uint64_t*>(sl.acc.ptr);
    a.n_blocks = n_b64 * (32u / unsigned(e->ppw)); a.dec_rows = uint32_t(ring); a.start_state = uint32_t(start_state);
    a.pk_steps = uint32_t(S);
    size_t next_seg = 0;
    for (size_t c = 0; c < n_win; c++) {
        const size_t step0 = c * W, nst = (S - step0 < W) ? (S - step0) : W;
        a.n_steps = uint32_t(nst); a.pk_step0 = uint32_t(step0); a.dec_row0 = uint32_t((c & 1) * W); a.resume = c ? 1u : 0u;
        h->launches++;
        VITB_CUDA(h, e->launch(a, s));
        // window c-1 can be walked once rows up to (c+1) W - 1 exist (its warm-up ends there); the windows whose warm-up would
        // reach the end of the frame wait for the end state
        if (d_out && c >= 1 && next_seg == c - 1 && next_seg + 1 < n_seg && (c + 1) * W <= S) {
            VITB_CUDA(h, walk_segment(next_seg));
            next_seg++;
        }
    }
    VITB_CUDA(h, mark(h, s));
    const uint32_t* end_states = nullptr;
    if (end_state == VITB_END_STATE_BEST) {
        VITB_CUDA(h, sl.end_states.reserve(n_frames * 4));
        h->launches++;
        best_state_kernel<<<unsigned((n_frames + 3) / 4), 128, 0, s>>>(a.metrics, uint32_t(h->n_states), uint32_t(n_frames),
                                                                        static_cast<uint32_t*>(sl.end_states.ptr));
        VITB_CUDA(h, cudaGetLastError());
        end_states = static_cast<const uint32_t*>(sl.end_states.ptr);
        t.end_states = end_states;
        end_state = 0;
    }
    if (d_out) {
        for (; next_seg < n_seg; next_seg++) VITB_CUDA(h, walk_segment(next_seg));
        h->launches++;
        traceback_seg_count_mismatch_kernel<<<unsigned((n_frames + 127) / 128), 128, 0, s>>>(sp.spec, sp.fin, uint32_t(n_seg), uint32_t(n_frames),
                                                                                            static_cast<unsigned long long*>(h->win_count.ptr));
        VITB_CUDA(h, cudaGetLastError());
    }
    VITB_CUDA(h, mark(h, s));
    if (d_acc || d_final) {
        h->launches++;
        gather_results_kernel<<<unsigned((n_frames + 255) / 256), 256, 0, s>>>(a.acc, a.metrics, uint32_t(h->n_states), uint32_t(end_state),
                                                                              end_states, uint32_t(n_frames), d_acc, d_final);
        VITB_CUDA(h, cudaGetLastError());
    }
    VITB_CUDA(h, mark(h, s));
    VITB_CUDA(h, cudaEventRecord(sl.tb_done, s));
    sl.tb_pending = true;
    return VITB_OK;
}

// K = 15: frames of a batch that go to the decision-row kernel (whole waves of n_sm frame pairs); the rest, if it fits one wave of
// single-frame CTAs, is decoded as a second chunk by the history kernel (1024 frames on 148 SMs: 3 waves of pairs + 136 single
// frames = 3 x 14.0 + 10.4 ms instead of 4 x 14.0 ms).  Returns n_frames when there is nothing to split.
size_t cta_wave_split(const vitb_decoder* h, size_t n_frames) {
    if (!h->hg_entry || h->hg_entry->layout != LAYOUT_HISTCTA || h->forced_variant != 0 || !h->use_hist || h->n_depunctured) return n_frames;
    const size_t wave = size_t(h->n_sm) * 2, first = n_frames / wave * wave, rem = n_frames - first;
    return (first && rem && rem <= size_t(h->n_sm)) ? first : n_frames;
}

int check_batch_args(const vitb_decoder* h, size_t n_frames, size_t L, const vitb_batch_opts* o, size_t* row_stride, size_t* start, size_t* end) {
    if (!h) return VITB_ERR_ARG;
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), n_sym = (L + K - 1) * R;
    if (n_frames == 0) return VITB_OK;
    if (n_frames > 0x7fffffffu || L > 0x3fffffffu) return VITB_ERR_ARG;
    size_t used = n_sym;
    if (h->n_depunctured) {
        if (h->n_depunctured != n_sym) return VITB_ERR_ARG;
        used = h->n_received;
    }
    *row_stride = (o && o->row_stride) ? o->row_stride : used;
    if (*row_stride < used) return VITB_ERR_ARG;
    *start = o ? o->starting_state : 0;
    *end = o ? o->end_state : 0;
    if (*end >= size_t(h->n_states) && *end != VITB_END_STATE_BEST) return VITB_ERR_ARG;      // core.h:196, 218
    return VITB_OK;
}

size_t chunk_frames_for(const vitb_decoder* h, size_t L, bool with_packed_stream = true, bool two_slots = false) {
    const size_t S = L + size_t(h->prm.K) - 1;
    const size_t limit = h->ws_limit ? h->ws_limit : h->ws_default;
    size_t blocks = limit / block_bytes(h, S, with_packed_stream, two_slots);
    if (blocks < 1) blocks = 1;
    if (blocks > 65535) blocks = 65535;       // the ingest kernel and the segmented tracebacks index 64-frame blocks with gridDim.y
    return blocks * 64;
}

}  // namespace

extern "C" {

const char* vitb_version(void) { return "viterbi_b200 0.1 (sm_100a)"; }

int vitb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* vitb_status_string(int status) {
    switch (status) {
    case VITB_OK: return "ok";
    case VITB_ERR_ARG: return "invalid argument (violates a reference precondition)";
    case VITB_ERR_UNSUPPORTED: return "no kernel compiled for this (K, R, G, types) combination";
    case VITB_ERR_CUDA: return "CUDA runtime error";
    case VITB_ERR_STATE: return "call protocol violated";
    case VITB_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}

int vitb_is_supported(const vitb_params* p) { return (p && params_valid(*p) && !find_entries(*p).empty()) ? 1 : 0; }

int vitb_create(const vitb_params* p, vitb_decoder** out) {
    if (!p || !out) return VITB_ERR_ARG;
    *out = nullptr;
    if (!params_valid(*p)) return VITB_ERR_ARG;
    const std::vector<const KernelEntry*> found = find_entries(*p);
    if (found.empty()) return VITB_ERR_UNSUPPORTED;
    const KernelEntry* e = found.front();
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return VITB_ERR_CUDA; }
    if (p->device < 0 || p->device >= n_dev) return VITB_ERR_ARG;
    vitb_decoder* h = new (std::nothrow) vitb_decoder();
    if (!h) return VITB_ERR_NOMEM;
    h->prm = *p;
    h->variants = found;
    h->entry = e;
    h->hg_entry = find_hist_group(*p);
    if (const char* f = getenv("VITB_FORCE_LOGT")) h->forced_variant = 1 << atoi(f);
    h->n_states = 1 << (p->K - 1);
    h->sh = e->sh;
    if (e->generic) {
        h->gcode.R = uint32_t(p->R);
        for (uint32_t j = 0; j < uint32_t(h->n_states) / 2; j++) h->gcode.pat[j] = uint8_t(bfly_pattern_rt(p->G, p->R, j));
    }
    cudaError_t ce = cudaSetDevice(p->device);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, p->device);
    if (ce == cudaSuccess) h->ws_default = default_ws_limit();
    if (ce == cudaSuccess) {
        if (const char* g = getenv("VITB_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(atoi(g)));   // experiment knob
        cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->tb_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->acs_stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { delete h; return VITB_ERR_CUDA; }
    *out = h;
    // core.h:175-176: the constructor leaves the decoder reset with traceback length 0
    int rc = vitb_set_traceback_length(h, 0);
    if (rc == VITB_OK) rc = vitb_reset(h, 0);
    if (rc != VITB_OK) { vitb_destroy(h); *out = nullptr; }
    return rc;
}

int vitb_destroy(vitb_decoder* h) {
    if (!h) return VITB_OK;
    cudaSetDevice(h->prm.device);
    for (BatchSlot& sl : h->slots) {
        for (DeviceBuffer* b : {&sl.dec, &sl.metrics, &sl.acc, &sl.tb_spec, &sl.tb_fin, &sl.end_states, &sl.pk}) b->release();
        if (sl.acs_done) cudaEventDestroy(sl.acs_done);
        if (sl.tb_done) cudaEventDestroy(sl.tb_done);
        if (sl.in_ready) cudaEventDestroy(sl.in_ready);
    }
    if (h->tb_stream) cudaStreamDestroy(h->tb_stream);
    if (h->acs_stream) cudaStreamDestroy(h->acs_stream);
    for (DeviceBuffer* b : {&h->pk, &h->d_in, &h->d_out, &h->d_accout, &h->d_finout, &h->map, &h->end_states, &h->win_count,
                            &h->s_pk, &h->s_dec, &h->s_metrics, &h->s_acc, &h->s_in, &h->s_out, &h->g_tx, &h->g_sym, &h->g_cnt}) b->release();
    h->s_map.release();
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->copy_ev) cudaEventDestroy(e);
    if (h->batch_done) cudaEventDestroy(h->batch_done);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return VITB_OK;
}

int vitb_last_cuda_error(const vitb_decoder* h) { return h ? h->last_cuda : 0; }
const char* vitb_kernel_name(const vitb_decoder* h) {
    if (!h) return "";
    if (!h->name_override.empty()) return h->name_override.c_str();
    if (h->last_batch && h->last_batch_hist) {      // "acs<...>" -> "acs_hist<...>"
        vitb_decoder* m = const_cast<vitb_decoder*>(h);
        m->name_buf = std::string("acs_hist") + (h->last_batch->name + 3);
        return m->name_buf.c_str();
    }
    return h->last_batch ? h->last_batch->name : (h->entry ? h->entry->name : "");
}

int vitb_set_variant(vitb_decoder* h, int lanes_per_pair) {
    if (!h) return VITB_ERR_ARG;
    if (lanes_per_pair <= 0) { h->forced_variant = 0; return VITB_OK; }
    for (const KernelEntry* e : h->variants) {
        if (entry_variant(e) == lanes_per_pair) { h->forced_variant = lanes_per_pair; return VITB_OK; }
    }
    if (h->hg_entry && entry_variant(h->hg_entry) == lanes_per_pair) { h->forced_variant = lanes_per_pair; return VITB_OK; }
    return VITB_ERR_UNSUPPORTED;
}

int vitb_set_traceback_segments(vitb_decoder* h, int seg_records, int overlap_records) {
    if (!h || seg_records < 0 || overlap_records < -1) return VITB_ERR_ARG;
    h->seg_records_forced = seg_records;
    h->seg_overlap_forced = overlap_records;
    return VITB_OK;
}

int vitb_set_history_kernel(vitb_decoder* h, int enabled) {
    if (!h) return VITB_ERR_ARG;
    h->use_hist = enabled != 0;
    return VITB_OK;
}

int vitb_get_variants(const vitb_decoder* h, int* lanes_per_pair, int capacity) {
    if (!h) return VITB_ERR_ARG;
    int n = 0;
    if (h->hg_entry) {          // batch-only variant, listed first (fewest lanes)
        if (lanes_per_pair && n < capacity) lanes_per_pair[n] = entry_variant(h->hg_entry);
        n++;
    }
    for (const KernelEntry* e : h->variants) {
        if (lanes_per_pair && n < capacity) lanes_per_pair[n] = entry_variant(e);
        n++;
    }
    return n;
}
int vitb_kernel_launch_count(const vitb_decoder* h, uint64_t* count) {
    if (!h || !count) return VITB_ERR_ARG;
    *count = h->launches;
    return VITB_OK;
}

int vitb_set_workspace_limit(vitb_decoder* h, size_t bytes) {
    if (!h) return VITB_ERR_ARG;
    h->ws_limit = bytes;
    return VITB_OK;
}

int vitb_workspace_bytes(const vitb_decoder* h, size_t n_frames, size_t L, size_t* bytes) {
    if (!h || !bytes) return VITB_ERR_ARG;
    const size_t S = L + size_t(h->prm.K) - 1;
    *bytes = ((n_frames + 63) / 64) * block_bytes(h, S);
    return VITB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// single-frame streaming API
// ------------------------------------------------------------------------------------------------------------------
int vitb_set_traceback_length(vitb_decoder* h, size_t traceback_length) {
    if (!h) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const size_t rows = traceback_length + size_t(h->prm.K - 1);       // core.h:181
    const size_t old_rows = h->traceback_length + size_t(h->prm.K - 1);
    if (!h->s_dec.ptr || rows != old_rows) {
        // std::vector::resize keeps the leading rows (core.h:182): so do we.  The streaming state is one warp block (block 0).
        const size_t row_bytes = dec_row_bytes_unit(h->entry);
        DeviceBuffer nb;
        VITB_CUDA(h, nb.reserve(rows * row_bytes + 16));
        cudaError_t ce = cudaMemsetAsync(nb.ptr, 0, rows * row_bytes + 16, h->stream);
        if (ce == cudaSuccess && h->s_dec.ptr) {
            const size_t keep = rows < old_rows ? rows : old_rows;
            ce = cudaMemcpyAsync(nb.ptr, h->s_dec.ptr, keep * row_bytes, cudaMemcpyDeviceToDevice, h->stream);
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (ce != cudaSuccess) { nb.release(); return cuda_fail(h, ce); }      // nb is a plain handle: free it on every early return
        h->s_dec.release();
        h->s_dec = nb;
    }
    h->traceback_length = traceback_length;
    if (h->current_decoded_bit > rows) h->current_decoded_bit = rows;   // core.h:183-185
    return VITB_OK;
}

int vitb_get_traceback_length(const vitb_decoder* h, size_t* traceback_length) {
    if (!h || !traceback_length) return VITB_ERR_ARG;
    *traceback_length = h->traceback_length;
    return VITB_OK;
}

int vitb_get_current_decoded_bit(const vitb_decoder* h, size_t* bit) {
    if (!h || !bit) return VITB_ERR_ARG;
    *bit = h->current_decoded_bit;
    return VITB_OK;
}

int vitb_reset(vitb_decoder* h, size_t starting_state) {
    if (!h) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const size_t ns = size_t(h->n_states);
    const uint32_t emask = (h->prm.soft_bytes == 1) ? 0xffu : 0xffffu;
    std::vector<uint16_t> m(64 * ns, uint16_t(h->prm.initial_non_start_error & emask));     // core.h:205-208
    for (size_t f = 0; f < 64; f++) m[f * ns + (starting_state & (ns - 1))] = uint16_t(h->prm.initial_start_error & emask);  // core.h:209-210
    VITB_CUDA(h, h->s_metrics.reserve(m.size() * 2));
    VITB_CUDA(h, h->s_acc.reserve(64 * 8));
    VITB_CUDA(h, cudaMemcpyAsync(h->s_metrics.ptr, m.data(), m.size() * 2, cudaMemcpyHostToDevice, h->stream));
    VITB_CUDA(h, cudaMemsetAsync(h->s_acc.ptr, 0, 64 * 8, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->current_decoded_bit = 0;                                                             // core.h:203
    return VITB_OK;
}

int vitb_update(vitb_decoder* h, const void* symbols, size_t n_symbols, uint64_t* accumulated_error) {
    if (!h || (!symbols && n_symbols)) return VITB_ERR_ARG;
    const size_t R = size_t(h->prm.R);
    if (n_symbols % R) return VITB_ERR_ARG;                                                 // scalar.h:37
    const size_t steps = n_symbols / R;
    const size_t rows = h->traceback_length + size_t(h->prm.K - 1);
    if (steps + h->current_decoded_bit > rows) return VITB_ERR_ARG;                         // scalar.h:38-40
    if (accumulated_error) *accumulated_error = 0;
    if (steps == 0) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const size_t sb = size_t(h->prm.soft_bytes);
    VITB_CUDA(h, h->s_pk.reserve(n_symbols * 32 * 4));
    // Calls of up to 64 KB of symbols (the reference's programs call update once per trellis step or per puncture segment:
    // run_punctured_decoder.cpp, puncture_code_helpers.h:32-52) go through mapped pinned memory: the ingest kernel reads the symbols
    // from there, the ACS kernel leaves the call's accumulated minima there - two launches and one synchronize per call.
    constexpr size_t MAP_ACC_BYTES = 512, MAP_SYM_BYTES = 64 * 1024;
    const bool mapped = n_symbols * sb <= MAP_SYM_BYTES && !getenv("VITB_NO_MAPPED_STREAMING");
    const void* d_sym;
    uint64_t* d_acc;
    if (mapped) {
        VITB_CUDA(h, h->s_map.reserve(MAP_ACC_BYTES + MAP_SYM_BYTES));
        memcpy(static_cast<char*>(h->s_map.host) + MAP_ACC_BYTES, symbols, n_symbols * sb);
        memset(h->s_map.host, 0, MAP_ACC_BYTES);                                // update() returns the minima of THIS call (scalar.h:42)
        d_sym = static_cast<const char*>(h->s_map.dev) + MAP_ACC_BYTES;
        d_acc = static_cast<uint64_t*>(h->s_map.dev);
    } else {
        VITB_CUDA(h, h->s_in.reserve(n_symbols * sb));
        VITB_CUDA(h, cudaMemcpyAsync(h->s_in.ptr, symbols, n_symbols * sb, cudaMemcpyHostToDevice, h->stream));
        VITB_CUDA(h, cudaMemsetAsync(h->s_acc.ptr, 0, 64 * 8, h->stream));
        d_sym = h->s_in.ptr;
        d_acc = static_cast<uint64_t*>(h->s_acc.ptr);
    }

    IngestParams ip{};
    ip.symbols = d_sym; ip.row_stride = n_symbols; ip.n_frames = 1; ip.n_sym = uint32_t(n_symbols);
    ip.depuncture_map = nullptr; ip.fill_value = 0; ip.pk = static_cast<uint32_t*>(h->s_pk.ptr); ip.ppw = uint32_t(h->entry->ppw);
    VITB_CUDA(h, run_ingest(h, ip, 1, h->stream));

    AcsParams a{};
    fill_acs_params(h, a);
    a.pk = static_cast<const uint32_t*>(h->s_pk.ptr); a.dec = h->s_dec.ptr;
    a.metrics = static_cast<uint16_t*>(h->s_metrics.ptr); a.acc = d_acc;
    a.n_blocks = 1;      // warp block 0 holds the user's frame (frame 0); its other frames decode zeros and are ignored
    a.n_frames = 1;
    a.n_steps = uint32_t(steps); a.dec_rows = uint32_t(rows); a.dec_row0 = uint32_t(h->current_decoded_bit);
    a.resume = 1; a.start_state = 0;
    h->launches++;
    if (h->entry->generic) VITB_CUDA(h, h->entry->launch_generic(a, h->gcode, h->stream));
    else VITB_CUDA(h, h->entry->launch(a, h->stream));
    uint64_t acc = 0;
    if (!mapped) VITB_CUDA(h, cudaMemcpyAsync(&acc, h->s_acc.ptr, 8, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    if (mapped) acc = *static_cast<volatile uint64_t*>(h->s_map.host);
    h->current_decoded_bit += steps;                                                        // scalar.h:52
    if (accumulated_error) *accumulated_error = acc;
    return VITB_OK;
}

// best final state of the streaming frame (frame 0 of the streaming block), on the handle's stream
static int streaming_best_state(vitb_decoder* h, uint32_t* best) {
    VITB_CUDA(h, h->end_states.reserve(4));
    h->launches++;
    best_state_kernel<<<1, 128, 0, h->stream>>>(static_cast<const uint16_t*>(h->s_metrics.ptr), uint32_t(h->n_states), 1u,
                                                static_cast<uint32_t*>(h->end_states.ptr));
    VITB_CUDA(h, cudaGetLastError());
    if (best) {
        VITB_CUDA(h, cudaMemcpyAsync(best, h->end_states.ptr, 4, cudaMemcpyDeviceToHost, h->stream));
        VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return VITB_OK;
}

int vitb_get_error(vitb_decoder* h, size_t end_state, uint32_t* error) {
    if (!h || !error) return VITB_ERR_ARG;
    if (end_state >= size_t(h->n_states) && end_state != VITB_END_STATE_BEST) return VITB_ERR_ARG;     // core.h:196
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    if (end_state == VITB_END_STATE_BEST) {
        uint32_t best = 0;
        const int r = streaming_best_state(h, &best);
        if (r != VITB_OK) return r;
        end_state = best;
    }
    uint16_t v = 0;
    VITB_CUDA(h, cudaMemcpyAsync(&v, static_cast<uint16_t*>(h->s_metrics.ptr) + end_state, 2, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    *error = v;
    return VITB_OK;
}

int vitb_get_metrics(vitb_decoder* h, uint32_t* metrics_out) {
    if (!h || !metrics_out) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    std::vector<uint16_t> m(size_t(h->n_states));
    VITB_CUDA(h, cudaMemcpyAsync(m.data(), h->s_metrics.ptr, m.size() * 2, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < m.size(); i++) metrics_out[i] = m[i];
    return VITB_OK;
}

// the reference's m_metrics and m_current_decoded_bit are public members (core.h:238-242): callers may assign them
int vitb_set_metrics(vitb_decoder* h, const uint32_t* metrics_in) {
    if (!h || !metrics_in) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    if (!h->s_metrics.ptr) return VITB_ERR_STATE;
    const uint32_t emask = (h->prm.soft_bytes == 1) ? 0xffu : 0xffffu;
    std::vector<uint16_t> m(size_t(h->n_states));
    for (size_t i = 0; i < m.size(); i++) m[i] = uint16_t(metrics_in[i] & emask);
    VITB_CUDA(h, cudaMemcpyAsync(h->s_metrics.ptr, m.data(), m.size() * 2, cudaMemcpyHostToDevice, h->stream));      // frame 0 of the streaming block
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    return VITB_OK;
}

int vitb_set_current_decoded_bit(vitb_decoder* h, size_t bit) {
    if (!h) return VITB_ERR_ARG;
    if (bit > h->traceback_length + size_t(h->prm.K - 1)) return VITB_ERR_ARG;       // core.h:183-185 keeps it inside the decision rows
    h->current_decoded_bit = bit;
    return VITB_OK;
}

int vitb_get_decisions(vitb_decoder* h, size_t first_row, size_t n_rows, uint64_t* rows_out) {
    if (!h || (!rows_out && n_rows)) return VITB_ERR_ARG;
    const size_t rows = h->traceback_length + size_t(h->prm.K - 1);
    if (first_row + n_rows > rows) return VITB_ERR_ARG;
    if (!n_rows) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const KernelEntry* e = h->entry;
    if (e->layout == LAYOUT_PAIR) {
        // frame 0 of the block: one uint64 every 64 words, already in the reference bit order
        VITB_CUDA(h, cudaMemcpy2DAsync(rows_out, 8, static_cast<uint64_t*>(h->s_dec.ptr) + first_row * 64, 64 * 8, 8, n_rows,
                                       cudaMemcpyDeviceToHost, h->stream));
        VITB_CUDA(h, cudaStreamSynchronize(h->stream));
        if (e->tagged_rows) {      // tagged row layout -> reference bit order
            for (size_t r = 0; r < n_rows; r++) {
                const uint64_t w = rows_out[r];
                uint64_t o = 0;
                for (uint32_t st = 0; st < uint32_t(h->n_states); st++) {
                    const uint32_t J = st >> 1, idx = (((J >> 3) * 2 + (st & 1u)) << 3) + 7u - (J & 7u);
                    o |= ((w >> idx) & 1ull) << st;
                }
                rows_out[r] = o;
            }
        }
        return VITB_OK;
    }
    // group layout (acs_group.cuh): lane t, bit q of row r holds the decision of state rotl^(n+1)((q << logt) | t), n = r % LB
    const uint32_t g = uint32_t(e->logt), T = 1u << g, SB = uint32_t(h->prm.K - 1), LB = SB - g, NL = 1u << LB, W = uint32_t(e->dec_words);
    const size_t words_per_row = dec_row_bytes_unit(e) / 4, out_words = (size_t(h->n_states) + 63) / 64;
    std::vector<uint32_t> raw(n_rows * words_per_row);
    VITB_CUDA(h, cudaMemcpyAsync(raw.data(), static_cast<uint32_t*>(h->s_dec.ptr) + first_row * words_per_row, raw.size() * 4,
                                 cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (size_t r = 0; r < n_rows; r++) {
        uint64_t* out = rows_out + r * out_words;
        for (size_t w = 0; w < out_words; w++) out[w] = 0;
        const uint32_t n = uint32_t((first_row + r) % LB);
        for (uint32_t t = 0; t < T; t++) {
            const uint32_t* lane_words = &raw[r * words_per_row + size_t(t) * W];      // pair 0 = lanes 0..T-1, frame A
            for (uint32_t q = 0; q < NL; q++) {
                const uint32_t bit = (lane_words[0] >> q) & 1u;
                if (!bit) continue;
                const uint32_t s2 = rotl_bits((q << g) | t, int(n + 1), int(SB));
                out[s2 / 64] |= uint64_t(1) << (s2 % 64);
            }
        }
    }
    return VITB_OK;
}

int vitb_chainback(vitb_decoder* h, uint8_t* bytes_out, size_t total_bits, size_t end_state) {
    if (!h || (!bytes_out && total_bits)) return VITB_ERR_ARG;
    if (h->traceback_length < total_bits) return VITB_ERR_ARG;                              // core.h:216
    if (h->current_decoded_bit < size_t(h->prm.K - 1) + total_bits) return VITB_ERR_STATE;  // core.h:217
    if (end_state >= size_t(h->n_states) && end_state != VITB_END_STATE_BEST) return VITB_ERR_ARG;     // core.h:218
    if (!total_bits) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const size_t nbytes = (total_bits + 7) / 8;
    VITB_CUDA(h, h->s_out.reserve(nbytes));
    const uint32_t* end_states = nullptr;
    if (end_state == VITB_END_STATE_BEST) {
        const int r = streaming_best_state(h, nullptr);
        if (r != VITB_OK) return r;
        end_states = static_cast<const uint32_t*>(h->end_states.ptr);
        end_state = 0;
    }
    VITB_CUDA(h, launch_traceback(h, h->entry, h->s_dec.ptr, h->traceback_length + size_t(h->prm.K - 1), 1, total_bits, end_state,
                                  static_cast<uint8_t*>(h->s_out.ptr), nbytes, h->stream, end_states));
    VITB_CUDA(h, cudaMemcpyAsync(bytes_out, h->s_out.ptr, nbytes, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    return VITB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// batched API
// ------------------------------------------------------------------------------------------------------------------
int vitb_set_puncture_schedule(vitb_decoder* h, const uint8_t* keep, size_t n_depunctured, int32_t unpunctured_value) {
    if (!h) return VITB_ERR_ARG;
    if (n_depunctured == 0) { h->n_depunctured = 0; h->n_received = 0; return VITB_OK; }
    if (!keep || n_depunctured % size_t(h->prm.R) || n_depunctured > 0x7fffffffu) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    std::vector<int32_t> map(n_depunctured);
    int32_t next = 0;
    for (size_t i = 0; i < n_depunctured; i++) map[i] = keep[i] ? next++ : -1;               // puncture_code_helpers.h:29-45
    // a batch call of this handle may still be reading the old map on a non-blocking stream: order the copy after it
    VITB_CUDA(h, batch_begin(h, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    VITB_CUDA(h, h->map.reserve(n_depunctured * 4));
    VITB_CUDA(h, cudaMemcpyAsync(h->map.ptr, map.data(), n_depunctured * 4, cudaMemcpyHostToDevice, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->n_depunctured = n_depunctured;
    h->n_received = size_t(next);
    h->unpunctured_value = unpunctured_value;
    return VITB_OK;
}

int vitb_decode_batch_dev(vitb_decoder* h, const void* d_symbols, size_t n_frames, size_t L, const vitb_batch_opts* opts,
                          uint8_t* d_out, uint64_t* d_acc, uint32_t* d_final, void* stream) {
    size_t row_stride = 0, start = 0, end = 0;
    const int rc = check_batch_args(h, n_frames, L, opts, &row_stride, &start, &end);
    if (rc != VITB_OK || n_frames == 0) return rc;
    if (!d_symbols) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    h->ev_used = 0;
    VITB_CUDA(h, batch_begin(h, s));
    // a batch that fits the workspace is one chunk in one slot; pipelined calls and multi-chunk batches keep two slots alive
    size_t chunk = chunk_frames_for(h, L, true, h->pipelined);
    const size_t out_stride = (L + 7) / 8, sb = size_t(h->prm.soft_bytes);
    const KernelEntry* e = choose_variant(h, n_frames < chunk ? n_frames : chunk);
    h->last_batch = e;
    const bool no_pk = direct_fetch_ok(h, e, d_symbols, row_stride);                                               // no packed stream to pay for
    if (n_frames > chunk && no_pk) chunk = chunk_frames_for(h, L, false, h->pipelined);
    if (n_frames > chunk) chunk = chunk_frames_for(h, L, !no_pk, true);
    h->overlap_hint = h->pipelined || n_frames > chunk;
    const bool windowed = h->window_bits && e->layout == LAYOUT_CTA && !e->generic && L + size_t(h->prm.K) - 1 > 2 * h->window_bits;
    if (windowed) {
        VITB_CUDA(h, h->win_count.reserve(8));
        VITB_CUDA(h, cudaMemsetAsync(h->win_count.ptr, 0, 8, s));
        h->name_override.clear();
        for (size_t f0 = 0; f0 < n_frames; f0 += chunk) {
            const size_t nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
            const int r = decode_chunk_window(h, e, static_cast<const uint8_t*>(d_symbols) + f0 * row_stride * sb, row_stride, nf, L, start, end,
                                              d_out ? d_out + f0 * out_stride : nullptr, d_acc ? d_acc + f0 : nullptr,
                                              d_final ? d_final + f0 : nullptr, s);
            if (r != VITB_OK) return r;
        }
        VITB_CUDA(h, batch_end(h, s));
        return VITB_OK;
    }
    const size_t split = (n_frames <= chunk) ? cta_wave_split(h, n_frames) : n_frames;
    if (split < n_frames) chunk = split;                  // K = 15: whole waves of frame pairs, then the rest one frame per CTA
    if (!h->pipelined && n_frames <= chunk) h->slot_next = 0;      // isolated single-chunk calls stay in one slot (half the workspace)
    h->name_override.clear();
    for (size_t f0 = 0; f0 < n_frames; f0 += chunk) {
        const size_t nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        const int r = decode_chunk_dev(h, e, static_cast<const uint8_t*>(d_symbols) + f0 * row_stride * sb, row_stride, nf, L, start, end,
                                       d_out ? d_out + f0 * out_stride : nullptr, d_acc ? d_acc + f0 : nullptr,
                                       d_final ? d_final + f0 : nullptr, s);
        if (r != VITB_OK) return r;
    }
    if (split < n_frames) h->name_override = std::string(e->name) + " for " + std::to_string(split) + " frames + acs_hist" + (h->hg_entry->name + 3) + " for the rest";
    // results in stream order at return - or, for pipelined calls, everything but the newest chunk (see vitb_set_pipelining)
    VITB_CUDA(h, join_slots(h, s, h->pipelined));
    VITB_CUDA(h, batch_end(h, s));
    return VITB_OK;
}

int vitb_set_traceback_window(vitb_decoder* h, size_t window_bits) {
    if (!h) return VITB_ERR_ARG;
    if (window_bits == 0) { h->window_bits = 0; return VITB_OK; }
    if (h->variants.front()->layout != LAYOUT_CTA) return VITB_ERR_UNSUPPORTED;      // K <= 9 keeps every row: a frame's rows are small there
    // whole bytes of output and whole exchange periods of the kernel per window; the warm-up (window - (K-1) rows) must be positive
    const size_t unit = 8 * (size_t(h->prm.K) - 1 - size_t(h->variants.front()->logt));
    size_t w = (window_bits + unit - 1) / unit * unit;
    if (w < 2 * unit) w = 2 * unit;
    h->window_bits = w;
    return VITB_OK;
}

int vitb_get_window_mismatches(vitb_decoder* h, uint64_t* count) {
    if (!h || !count) return VITB_ERR_ARG;
    *count = 0;
    if (!h->win_count.ptr) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    VITB_CUDA(h, cudaDeviceSynchronize());
    unsigned long long c = 0;
    VITB_CUDA(h, cudaMemcpy(&c, h->win_count.ptr, 8, cudaMemcpyDeviceToHost));
    *count = c;
    return VITB_OK;
}

int vitb_set_pipelining(vitb_decoder* h, int enabled) {
    if (!h) return VITB_ERR_ARG;
    h->pipelined = enabled != 0;
    return VITB_OK;
}

int vitb_batch_flush(vitb_decoder* h, void* stream) {
    if (!h) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    VITB_CUDA(h, join_slots(h, static_cast<cudaStream_t>(stream), false));
    return VITB_OK;
}

// Host-pointer path, pipelined: the batch is cut into a few chunks of frames; chunk c+1 is copied host->device on the handle's copy
// stream while chunk c is decoded on the caller's stream, and each chunk's results go back as soon as they exist.  End to end the
// call then costs about max(PCIe copy, kernels) instead of their sum (config 2: 269 MB over PCIe ~ 5 ms vs 1.6 ms of kernels).
int vitb_decode_batch_async(vitb_decoder* h, const void* symbols, size_t n_frames, size_t L, const vitb_batch_opts* opts,
                            uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error, void* stream) {
    size_t row_stride = 0, start = 0, end = 0;
    const int rc = check_batch_args(h, n_frames, L, opts, &row_stride, &start, &end);
    if (rc != VITB_OK || n_frames == 0) return rc;
    if (!symbols) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t sb = size_t(h->prm.soft_bytes), out_stride = (L + 7) / 8;
    const size_t row_bytes = row_stride * sb, in_bytes = n_frames * row_bytes;
    VITB_CUDA(h, batch_begin(h, s));       // also keeps d_in / d_out of a call still in flight on another stream intact
    VITB_CUDA(h, h->d_in.reserve(in_bytes));
    if (out_bytes) VITB_CUDA(h, h->d_out.reserve(n_frames * out_stride));
    if (acc_error) VITB_CUDA(h, h->d_accout.reserve(n_frames * 8));
    if (final_error) VITB_CUDA(h, h->d_finout.reserve(n_frames * 4));
    if (h->window_bits && h->variants.front()->layout == LAYOUT_CTA && L + size_t(h->prm.K) - 1 > 2 * h->window_bits) {
        // sliding-window mode: copy in, the device-pointer call, copy out, all on `s`
        VITB_CUDA(h, cudaMemcpyAsync(h->d_in.ptr, symbols, in_bytes, cudaMemcpyHostToDevice, s));
        vitb_batch_opts o{}; o.row_stride = row_stride; o.starting_state = start; o.end_state = end;
        const int r = vitb_decode_batch_dev(h, h->d_in.ptr, n_frames, L, &o, out_bytes ? static_cast<uint8_t*>(h->d_out.ptr) : nullptr,
                                            acc_error ? static_cast<uint64_t*>(h->d_accout.ptr) : nullptr,
                                            final_error ? static_cast<uint32_t*>(h->d_finout.ptr) : nullptr, s);
        if (r != VITB_OK) return r;
        if (out_bytes) VITB_CUDA(h, cudaMemcpyAsync(out_bytes, h->d_out.ptr, n_frames * out_stride, cudaMemcpyDeviceToHost, s));
        if (acc_error) VITB_CUDA(h, cudaMemcpyAsync(acc_error, h->d_accout.ptr, n_frames * 8, cudaMemcpyDeviceToHost, s));
        if (final_error) VITB_CUDA(h, cudaMemcpyAsync(final_error, h->d_finout.ptr, n_frames * 4, cudaMemcpyDeviceToHost, s));
        VITB_CUDA(h, batch_end(h, s));
        return VITB_OK;
    }

    // pipeline chunks: at least ~16 MB of input each (PCIe efficiency), 4 chunks (measured, VITB_E2E_CHUNKS = 4 / 8 / 16: config 2
    // 5.54 / 5.55 / 8.03 ms, config 3 11.5 / 13.5 ms - smaller chunks underfill the GPU), whole 64-frame blocks, and never larger than what the workspace allows
    static const size_t want_chunks = getenv("VITB_E2E_CHUNKS") ? size_t(atoi(getenv("VITB_E2E_CHUNKS"))) : size_t(4);
    size_t chunk = (n_frames + want_chunks - 1) / (want_chunks ? want_chunks : 1);
    const size_t min_chunk = (size_t(16) << 20) / (row_bytes ? row_bytes : 1) + 1;
    if (chunk < min_chunk) chunk = min_chunk;
    const size_t ws_chunk = chunk_frames_for(h, L, true, true);
    if (chunk > ws_chunk) chunk = ws_chunk;
    // one-CTA-per-pair kernels (K = 15) run a few waves of CTAs per batch: splitting only makes the wave quantisation worse
    if (h->variants.front()->layout == LAYOUT_CTA && n_frames < size_t(h->n_sm) * 2 * 16) chunk = ws_chunk;
    chunk = (chunk + 63) / 64 * 64;
    const size_t n_chunks = (n_frames + chunk - 1) / chunk;
    while (h->copy_ev.size() < n_chunks + 1) {
        cudaEvent_t e;
        VITB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->copy_ev.push_back(e);
    }
    // fork: the copies must not start before the caller's earlier work on `s` (it may still be using the buffers of a previous call)
    VITB_CUDA(h, cudaEventRecord(h->copy_ev[n_chunks], s));
    VITB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->copy_ev[n_chunks], 0));
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t f0 = c * chunk, nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        VITB_CUDA(h, cudaMemcpyAsync(static_cast<uint8_t*>(h->d_in.ptr) + f0 * row_bytes, static_cast<const uint8_t*>(symbols) + f0 * row_bytes,
                                     nf * row_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        VITB_CUDA(h, cudaEventRecord(h->copy_ev[c], h->copy_stream));
    }
    h->ev_used = 0;
    const KernelEntry* e = choose_variant(h, n_frames < chunk ? n_frames : chunk);
    h->last_batch = e;
    h->overlap_hint = n_chunks > 1;
    h->name_override.clear();
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t f0 = c * chunk, nf = (n_frames - f0 < chunk) ? (n_frames - f0) : chunk;
        VITB_CUDA(h, cudaStreamWaitEvent(s, h->copy_ev[c], 0));
        uint8_t* d_out = out_bytes ? static_cast<uint8_t*>(h->d_out.ptr) + f0 * out_stride : nullptr;
        uint64_t* d_acc = acc_error ? static_cast<uint64_t*>(h->d_accout.ptr) + f0 : nullptr;
        uint32_t* d_fin = final_error ? static_cast<uint32_t*>(h->d_finout.ptr) + f0 : nullptr;
        // the chunk's results go back on the traceback stream, behind its gather and next to the ACS of the next chunk
        const int r = decode_chunk_dev(h, e, static_cast<const uint8_t*>(h->d_in.ptr) + f0 * row_bytes, row_stride, nf, L, start, end,
                                       d_out, d_acc, d_fin, s, out_bytes ? out_bytes + f0 * out_stride : nullptr,
                                       acc_error ? acc_error + f0 : nullptr, final_error ? final_error + f0 : nullptr);
        if (r != VITB_OK) return r;
    }
    VITB_CUDA(h, join_slots(h, s, false));
    VITB_CUDA(h, batch_end(h, s));
    return VITB_OK;
}

int vitb_decode_batch(vitb_decoder* h, const void* symbols, size_t n_frames, size_t L, const vitb_batch_opts* opts,
                      uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error) {
    if (!h) return VITB_ERR_ARG;
    const int r = vitb_decode_batch_async(h, symbols, n_frames, L, opts, out_bytes, acc_error, final_error, h->stream);
    if (r != VITB_OK) return r;
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    return VITB_OK;
}

int vitb_set_profiling(vitb_decoder* h, int enabled) {
    if (!h) return VITB_ERR_ARG;
    h->profiling = enabled != 0;
    h->ev_used = 0;
    return VITB_OK;
}

int vitb_get_stage_ms(vitb_decoder* h, float ms[4]) {
    if (!h || !ms) return VITB_ERR_ARG;
    for (int k = 0; k < 4; k++) ms[k] = 0.f;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    for (size_t c = 0; c + 5 <= h->ev_used; c += 5) {
        for (int k = 0; k < 4; k++) {
            float t = 0.f;
            VITB_CUDA(h, cudaEventElapsedTime(&t, h->ev[c + k], h->ev[c + k + 1]));
            ms[k] += t;
        }
    }
    return VITB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// device-side front end (frontend.cuh)
// ------------------------------------------------------------------------------------------------------------------
int vitb_synth_frames_dev(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed,
                          uint8_t* d_tx_bytes, void* d_symbols, size_t row_stride, void* stream) {
    if (!h || !d_tx_bytes || !d_symbols || (total_bits % 8) || n_frames > 0x7fffffffu || total_bits > 0x3fffffffu) return VITB_ERR_ARG;
    if (n_frames == 0 || total_bits == 0) return VITB_OK;
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), steps = total_bits + K - 1, n_sym = steps * R;
    size_t used = n_sym;
    if (h->n_depunctured) {
        if (h->n_depunctured != n_sym) return VITB_ERR_ARG;
        used = h->n_received;
    }
    if (row_stride == 0) row_stride = used;
    if (row_stride < used) return VITB_ERR_ARG;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SynthParams sp{};
    sp.K = h->prm.K; sp.R = h->prm.R;
    for (size_t i = 0; i < R; i++) sp.G[i] = h->prm.G[i];
    sp.high = h->prm.soft_decision_high; sp.low = h->prm.soft_decision_low;
    sp.n_frames = uint32_t(n_frames); sp.total_bits = uint32_t(total_bits);
    const float mag = (float(sp.high) - float(sp.low)) / 2.0f;                       // run_snr_ber.cpp:311-312
    sp.mean = (float(sp.high) + float(sp.low)) / 2.0f;
    if (std::isnan(EbNo_dB)) {
        sp.sigma = -1.f;
        sp.scale = mag;
    } else {
        const float EsNo_dB = EbNo_dB - 10.0f * std::log10(float(R));               // run_snr_ber.cpp:320-325
        const float noise_variance = std::pow(10.0f, -(EsNo_dB + 3.0f) / 10.0f);
        sp.sigma = std::sqrt(noise_variance);
        sp.scale = mag * (1.0f / std::sqrt(1.0f + noise_variance));
    }
    sp.seed = seed;
    sp.depuncture_map = h->n_depunctured ? static_cast<const int32_t*>(h->map.ptr) : nullptr;
    sp.tx_bytes = d_tx_bytes; sp.symbols = d_symbols; sp.row_stride = row_stride;
    const size_t nb = n_frames * (total_bits / 8), ns = n_frames * steps;
    h->launches += 2;
    synth_bytes_kernel<<<unsigned((nb + 255) / 256), 256, 0, s>>>(sp);
    VITB_CUDA(h, cudaGetLastError());
    if (h->prm.soft_bytes == 1) synth_symbols_kernel<int8_t><<<unsigned((ns + 255) / 256), 256, 0, s>>>(sp);
    else synth_symbols_kernel<int16_t><<<unsigned((ns + 255) / 256), 256, 0, s>>>(sp);
    VITB_CUDA(h, cudaGetLastError());
    return VITB_OK;
}

int vitb_synth_frames(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed,
                      uint8_t* tx_bytes, void* symbols, size_t row_stride) {
    if (!h || !tx_bytes || !symbols) return VITB_ERR_ARG;
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), n_sym = (total_bits + K - 1) * R;
    const size_t used = h->n_depunctured ? h->n_received : n_sym;
    if (row_stride == 0) row_stride = used;
    const size_t sb = size_t(h->prm.soft_bytes), nb = n_frames * (total_bits / 8), sbytes = n_frames * row_stride * sb;
    if (n_frames == 0) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    VITB_CUDA(h, h->g_tx.reserve(nb ? nb : 1));
    VITB_CUDA(h, h->g_sym.reserve(sbytes ? sbytes : 1));
    VITB_CUDA(h, cudaMemsetAsync(h->g_sym.ptr, 0, sbytes, h->stream));
    const int r = vitb_synth_frames_dev(h, n_frames, total_bits, EbNo_dB, seed, static_cast<uint8_t*>(h->g_tx.ptr), h->g_sym.ptr, row_stride, h->stream);
    if (r != VITB_OK) return r;
    VITB_CUDA(h, cudaMemcpyAsync(tx_bytes, h->g_tx.ptr, nb, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaMemcpyAsync(symbols, h->g_sym.ptr, sbytes, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    return VITB_OK;
}

int vitb_quantise(vitb_decoder* h, const float* x, size_t n, float scale, float mean, void* soft_out) {
    if (!h || (!x && n) || (!soft_out && n)) return VITB_ERR_ARG;
    if (!n) return VITB_OK;
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    const size_t sb = size_t(h->prm.soft_bytes);
    VITB_CUDA(h, h->g_tx.reserve(n * 4));
    VITB_CUDA(h, h->g_sym.reserve(n * sb));
    VITB_CUDA(h, cudaMemcpyAsync(h->g_tx.ptr, x, n * 4, cudaMemcpyHostToDevice, h->stream));
    h->launches++;
    if (sb == 1) quantise_kernel<int8_t><<<unsigned((n + 255) / 256), 256, 0, h->stream>>>(static_cast<const float*>(h->g_tx.ptr), n, scale, mean,
                                        h->prm.soft_decision_low, h->prm.soft_decision_high, static_cast<int8_t*>(h->g_sym.ptr));
    else quantise_kernel<int16_t><<<unsigned((n + 255) / 256), 256, 0, h->stream>>>(static_cast<const float*>(h->g_tx.ptr), n, scale, mean,
                                        h->prm.soft_decision_low, h->prm.soft_decision_high, static_cast<int16_t*>(h->g_sym.ptr));
    VITB_CUDA(h, cudaGetLastError());
    VITB_CUDA(h, cudaMemcpyAsync(soft_out, h->g_sym.ptr, n * sb, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    return VITB_OK;
}

// Generate, decode and count on the device: returns the number of bit errors of n_frames frames at EbNo_dB.  Nothing but the
// count crosses PCIe (the inner loop of examples/run_snr_ber.cpp:336-372 for one batch).
int vitb_ber_trial(vitb_decoder* h, size_t n_frames, size_t total_bits, float EbNo_dB, uint64_t seed, uint64_t* bit_errors) {
    if (!h || !bit_errors || (total_bits % 8)) return VITB_ERR_ARG;
    *bit_errors = 0;
    if (n_frames == 0 || total_bits == 0) return VITB_OK;
    const size_t K = size_t(h->prm.K), R = size_t(h->prm.R), n_sym = (total_bits + K - 1) * R;
    const size_t used = h->n_depunctured ? h->n_received : n_sym;
    const size_t row = (used * size_t(h->prm.soft_bytes) + 3) / 4 * 4 / size_t(h->prm.soft_bytes);      // 4-byte aligned rows: direct fetch
    const size_t sb = size_t(h->prm.soft_bytes), nb = n_frames * (total_bits / 8);
    VITB_CUDA(h, cudaSetDevice(h->prm.device));
    VITB_CUDA(h, h->g_tx.reserve(nb));
    VITB_CUDA(h, h->g_sym.reserve(n_frames * row * sb));
    VITB_CUDA(h, h->g_cnt.reserve(8));
    VITB_CUDA(h, h->d_out.reserve(nb));
    VITB_CUDA(h, cudaMemsetAsync(h->g_cnt.ptr, 0, 8, h->stream));
    VITB_CUDA(h, cudaMemsetAsync(h->g_sym.ptr, 0, n_frames * row * sb, h->stream));
    int r = vitb_synth_frames_dev(h, n_frames, total_bits, EbNo_dB, seed, static_cast<uint8_t*>(h->g_tx.ptr), h->g_sym.ptr, row, h->stream);
    if (r != VITB_OK) return r;
    vitb_batch_opts o{}; o.row_stride = row;
    r = vitb_decode_batch_dev(h, h->g_sym.ptr, n_frames, total_bits, &o, static_cast<uint8_t*>(h->d_out.ptr), nullptr, nullptr, h->stream);
    if (r != VITB_OK) return r;
    VITB_CUDA(h, join_slots(h, h->stream, false));
    h->launches++;
    bit_errors_kernel<<<296, 256, 0, h->stream>>>(static_cast<const uint8_t*>(h->g_tx.ptr), static_cast<const uint8_t*>(h->d_out.ptr), nb,
                                                  static_cast<unsigned long long*>(h->g_cnt.ptr));
    VITB_CUDA(h, cudaGetLastError());
    unsigned long long c = 0;
    VITB_CUDA(h, cudaMemcpyAsync(&c, h->g_cnt.ptr, 8, cudaMemcpyDeviceToHost, h->stream));
    VITB_CUDA(h, cudaStreamSynchronize(h->stream));
    *bit_errors = c;
    return VITB_OK;
}

int vitb_decode_batch_multi(vitb_decoder* const* handles, int n_handles, const void* symbols, size_t n_frames, size_t L,
                            const vitb_batch_opts* opts, uint8_t* out_bytes, uint64_t* acc_error, uint32_t* final_error) {
    if (!handles || n_handles < 1) return VITB_ERR_ARG;
    for (int i = 0; i < n_handles; i++) if (!handles[i]) return VITB_ERR_ARG;
    size_t row_stride = 0, start = 0, end = 0;
    const int rc = check_batch_args(handles[0], n_frames, L, opts, &row_stride, &start, &end);
    if (rc != VITB_OK || n_frames == 0) return rc;
    const size_t sb = size_t(handles[0]->prm.soft_bytes), out_stride = (L + 7) / 8;
    vitb_batch_opts o{}; o.row_stride = row_stride; o.starting_state = start; o.end_state = end;
    // contiguous range of frames per device, no exchange between devices
    auto range = [&](int i, size_t& f0, size_t& f1) { f0 = n_frames * size_t(i) / size_t(n_handles); f1 = n_frames * size_t(i + 1) / size_t(n_handles); };
    auto call = [&](int i, bool async) -> int {
        size_t f0, f1;
        range(i, f0, f1);
        if (f1 == f0) return VITB_OK;
        const void* in = static_cast<const uint8_t*>(symbols) + f0 * row_stride * sb;
        uint8_t* ob = out_bytes ? out_bytes + f0 * out_stride : nullptr;
        uint64_t* oa = acc_error ? acc_error + f0 : nullptr;
        uint32_t* of = final_error ? final_error + f0 : nullptr;
        return async ? vitb_decode_batch_async(handles[i], in, f1 - f0, L, &o, ob, oa, of, handles[i]->stream)
                     : vitb_decode_batch(handles[i], in, f1 - f0, L, &o, ob, oa, of);
    };
    // Pinned host memory: every copy is asynchronous, so ONE host thread enqueues the whole pipeline (copies in, kernels, copies out)
    // of every device on that device's streams and then waits for them: the devices run concurrently with no thread per call.
    // Pageable memory makes cudaMemcpyAsync block the calling thread, so there the ranges are decoded by one host thread per device.
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, symbols) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                        (!out_bytes || (cudaPointerGetAttributes(&pa, out_bytes) == cudaSuccess && pa.type == cudaMemoryTypeHost));
    cudaGetLastError();
    if (pinned) {
        int rc2 = VITB_OK;
        for (int i = 0; i < n_handles && rc2 == VITB_OK; i++) rc2 = call(i, true);
        for (int i = 0; i < n_handles; i++) {           // join every device, also after an error
            if (cudaSetDevice(handles[i]->prm.device) != cudaSuccess || cudaStreamSynchronize(handles[i]->stream) != cudaSuccess) {
                if (rc2 == VITB_OK) rc2 = cuda_fail(handles[i], cudaGetLastError());
            }
        }
        return rc2;
    }
    std::vector<int> results(size_t(n_handles), VITB_OK);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_handles; i++) threads.emplace_back([&, i] { results[size_t(i)] = call(i, false); });
    for (auto& t : threads) t.join();
    for (int r : results) if (r != VITB_OK) return r;
    return VITB_OK;
}

}  // extern "C"
