// Micro-benchmark v5 (round 2): what bounds the survivor-history butterfly (2 VIADD.16x2 + 2 VIADDMNMX.U16x2)?
//   * every number is timed with clock64 as the MAX over the warps of an SM (r01 took the median warp, which the hi-wid-first
//     arbiter inflates at 8 warps per scheduler) and cross-checked with CUDA events over the whole launch;
//   * step variants: ping-pong (the shipped form), in place (64 registers: more warps per scheduler fit), unfused (4 add + 2 min,
//     all 2-operand), adds as 32-bit IADD (timing only), adds on the ALU pipe (VIADDMNMX with an all-ones third operand),
//     ping-pong with the butterflies grouped by branch pattern (operand reuse);
//   * a one-byte-per-64-byte gather over 1 GiB at cudaLimitMaxL2FetchGranularity 64 / 32 (the traceback's access pattern).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITER = 2048;      // steps per launch

template <int MODE>
__device__ __forceinline__ void bfly(uint32_t a, uint32_t b, uint32_t t, uint32_t ti, uint32_t tt, uint32_t tti, uint32_t& y0, uint32_t& y1) {
    if (MODE == 0 || MODE == 1 || MODE == 5) {
        const uint32_t b0 = __vadd2(b, tti), b1 = __vadd2(b, tt);
        y0 = __viaddmin_u16x2(a, t, b0);
        y1 = __viaddmin_u16x2(a, ti, b1);
    } else if (MODE == 2) {
        const uint32_t b0 = __vadd2(b, tti), b1 = __vadd2(b, tt), a0 = __vadd2(a, t), a1 = __vadd2(a, ti);
        y0 = __vminu2(a0, b0);
        y1 = __vminu2(a1, b1);
    } else if (MODE == 3) {
        const uint32_t b0 = b + tti, b1 = b + tt;
        y0 = __viaddmin_u16x2(a, t, b0);
        y1 = __viaddmin_u16x2(a, ti, b1);
    } else {
        const uint32_t b0 = __viaddmin_u16x2(b, tti, 0xffffffffu), b1 = __viaddmin_u16x2(b, tt, 0xffffffffu);
        y0 = __viaddmin_u16x2(a, t, b0);
        y1 = __viaddmin_u16x2(a, ti, b1);
    }
}

template <int MODE>
__device__ __forceinline__ void step_pp(const uint32_t (&x)[64], uint32_t (&y)[64], const uint32_t (&T)[4], const uint32_t (&TT)[4]) {
#pragma unroll
    for (int J = 0; J < 32; J++) {
        const int pi = (MODE == 5) ? (J >> 3) : ((J * 5 + (J >> 3)) & 3);
        bfly<MODE>(x[J], x[J + 32], T[pi], T[pi ^ 3], TT[pi], TT[pi ^ 3], y[2 * J], y[2 * J + 1]);
    }
}

template <int MODE>
__device__ __forceinline__ void step_inplace(uint32_t (&x)[64], const uint32_t (&T)[4], const uint32_t (&TT)[4]) {
#pragma unroll
    for (int J = 0; J < 32; J++) {
        const int pi = (J * 5 + (J >> 3)) & 3;
        uint32_t y0, y1;
        bfly<MODE>(x[J], x[J + 32], T[pi], T[pi ^ 3], TT[pi], TT[pi ^ 3], y0, y1);
        x[J] = y0;
        x[J + 32] = y1;
    }
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) hist(uint32_t* out, long long* cycles, const uint32_t* in) {
    uint32_t x[64], y[64], T[4], TT[4];
#pragma unroll
    for (int i = 0; i < 64; i++) x[i] = (in[i & 31] + threadIdx.x * (i + 1)) & 0xff00ff00u;
#pragma unroll
    for (int i = 0; i < 4; i++) T[i] = in[32 + i] & 0x0f000f00u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; it++) {
        uint32_t tag = 0x00010001u;
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int i = 0; i < 4; i++) { T[i] = __vadd2(T[i], 0x01000100u); TT[i] = T[i] + tag; }
            if (MODE == 1) step_inplace<MODE>(x, T, TT); else step_pp<MODE>(x, y, T, TT);
#pragma unroll
            for (int i = 0; i < 4; i++) TT[i] = T[i] + (tag << 1);
            if (MODE == 1) step_inplace<MODE>(x, T, TT); else step_pp<MODE>(y, x, T, TT);
            tag <<= 2;
        }
#pragma unroll
        for (int i = 0; i < 64; i++) x[i] &= 0xff00ff00u;
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int MODE, int THREADS>
void run_one(const char* name, int wps, int nsm, const uint32_t* din, double clk_hz) {
    if (32 * 4 * wps != THREADS) return;
    uint32_t* out; long long* cyc;
    const int nw = nsm * 4 * wps;
    CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * THREADS)); CK(cudaMalloc(&cyc, sizeof(long long) * nw));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    hist<MODE, THREADS><<<nsm, THREADS>>>(out, cyc, din); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    hist<MODE, THREADS><<<nsm, THREADS>>>(out, cyc, din);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(nw);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost));
    // per SM: max over its warps; then the median SM
    std::vector<long long> sm(nsm);
    for (int s = 0; s < nsm; s++) sm[s] = *std::max_element(h.begin() + s * 4 * wps, h.begin() + (s + 1) * 4 * wps);
    std::sort(sm.begin(), sm.end());
    const double clk = double(sm[nsm / 2]);
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, hist<MODE, THREADS>));
    printf("%-46s w%d: %5.2f clk/bfly/SMSP (max-warp clock64)  %5.2f (events at %.0f MHz)  regs %d  spill %zu B\n", name, wps,
           clk / (double(ITER) * 32 * wps), double(ms) * 1e-3 * clk_hz / (double(ITER) * 32 * wps), clk_hz / 1e6, fa.numRegs, fa.localSizeBytes);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

template <int MODE>
void run_mode(const char* name, int nsm, const uint32_t* din, double clk_hz, bool many) {
    run_one<MODE, 128>(name, 1, nsm, din, clk_hz);
    run_one<MODE, 256>(name, 2, nsm, din, clk_hz);
    run_one<MODE, 384>(name, 3, nsm, din, clk_hz);
    if (many) {
        run_one<MODE, 512>(name, 4, nsm, din, clk_hz);
        run_one<MODE, 768>(name, 6, nsm, din, clk_hz);
    }
}

// one byte out of every 64-byte record, record chosen per thread, byte chosen pseudo-randomly: the traceback's access pattern
__global__ void gather(const uint8_t* buf, size_t n_rec, uint32_t* out, int per_thread) {
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nthr = size_t(gridDim.x) * blockDim.x;
    uint32_t acc = 0, s = uint32_t(tid) * 2654435761u;
    for (int i = 0; i < per_thread; i++) {
        const size_t rec = (tid + size_t(i) * nthr) % n_rec;
        s = s * 1664525u + 1013904223u;
        acc += buf[rec * 64 + ((s >> 24) & 63u) + (acc & 0u)];
    }
    out[tid] = acc;
}

void run_gather(int gran) {
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(gran)));
    size_t got = 0; CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
    const size_t bytes = size_t(1) << 30, n_rec = bytes / 64;
    uint8_t* buf; uint32_t* out;
    const int threads = 256, blocks = 148 * 16, per = int(n_rec / (size_t(threads) * blocks));
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, sizeof(uint32_t) * threads * blocks));
    CK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    gather<<<blocks, threads>>>(buf, n_rec, out, per); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    gather<<<blocks, threads>>>(buf, n_rec, out, per);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double lookups = double(per) * threads * blocks;
    printf("gather 1 B per 64 B record over 1 GiB, L2 fetch granularity %zu (asked %d): %.3f ms, %.1f G lookups/s = %.0f GB/s at 64 B, %.0f GB/s at 32 B per lookup\n",
           got, gran, ms, lookups / ms / 1e6, lookups * 64 / ms / 1e6, lookups * 32 / ms / 1e6);
    CK(cudaFree(buf)); CK(cudaFree(out));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double clk_hz = double(khz) * 1e3;
    uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x01230457u * (i + 3) | 0x00010001u;
    uint32_t* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("device %s, %d SMs, clock attribute %.0f MHz. K=7 survivor-history step, 32 butterflies per warp and step\n", p.name, nsm, clk_hz / 1e6);
    run_mode<0>("ping-pong 2 VIADD.16x2 + 2 VIADDMNMX.U16x2", nsm, din, clk_hz, false);
    run_mode<5>("ping-pong, butterflies grouped by pattern", nsm, din, clk_hz, false);
    run_mode<1>("in place   2 VIADD.16x2 + 2 VIADDMNMX.U16x2", nsm, din, clk_hz, true);
    run_mode<2>("ping-pong 4 VIADD.16x2 + 2 VIMNMX.U16x2", nsm, din, clk_hz, false);
    run_mode<3>("ping-pong 2 IADD(32) + 2 VIADDMNMX.U16x2", nsm, din, clk_hz, false);
    run_mode<4>("ping-pong 4 VIADDMNMX.U16x2 (ALU pipe only)", nsm, din, clk_hz, false);
    run_gather(64);
    run_gather(32);
    run_gather(128);
    return 0;
}
