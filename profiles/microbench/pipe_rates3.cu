// Micro-benchmark v7 (round 2): pipe cost of the packed-integer instructions of the survivor-history butterfly, measured the way the
// kernel uses them: 32 independent registers per lane (no dependent chains inside a pass), 4 warps per sub-partition, the whole
// launch timed with CUDA events (no per-warp clock64: the hi-wid-first arbiter lets some warps finish early and inflated the r01
// pair rates above 1 instruction per clock).
// Reports clocks per warp-instruction per sub-partition for pure streams and for mixes; a mix of two instructions that use
// DIFFERENT pipes costs max(a, b) per pair, the SAME pipe a + b.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int NR = 32, ITER = 4096;
enum { ADD2, MIN2, ADDMIN2, LOP, PRM, IADD, IMADK, MIN3, ADD2I, NONE, ADD2R, ADDMIN2R, MIN2R };

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c, uint32_t g_t0 = 0, uint32_t g_t1 = 0, uint32_t g_t2 = 0) {
    if (OP == ADD2) return __vadd2(a, b);
    if (OP == MIN2) return __vminu2(a, b);
    if (OP == ADDMIN2) return __viaddmin_u16x2(a, b, c);
    if (OP == LOP) return (a & b) ^ c;
    if (OP == PRM) return __byte_perm(a, b, 0x6420);
    if (OP == IADD) return a + b;
    if (OP == IMADK) return a * 3u + b;
    if (OP == MIN3) return __vimin3_u16x2(a, b, c);
    if (OP == ADD2I) return __vadd2(a, 0x01000100u);
    // "R" forms: every operand but the first is the SAME register in every instruction of the stream (operand reuse cache: one
    // register-file read per instruction)
    if (OP == ADD2R) return __vadd2(a, g_t0);
    if (OP == ADDMIN2R) return __viaddmin_u16x2(a, g_t1, g_t2);
    if (OP == MIN2R) return __vminu2(a, g_t1);
    return a;
}

// per pass: NA instructions of A and NB of B for every group of (NA + NB) registers
template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(128) mix(uint32_t* out, const uint32_t* in) {
    uint32_t x[NR], y[NR], T[4];
#pragma unroll
    for (int i = 0; i < NR; i++) { x[i] = in[i] + threadIdx.x * (i + 1); y[i] = in[(i + 7) & 31] ^ threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 4; i++) T[i] = in[32 + i];
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int k = i % (NA + NB);
            if (k < NA) x[i] = op<A>(x[i], T[i & 3], y[i], T[0], T[1], T[2]);
            else x[i] = op<B>(x[i], T[i & 3], y[i], T[0], T[1], T[2]);
        }
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int k = i % (NA + NB);
            if (k < NA) y[i] = op<A>(y[i], T[(i + 1) & 3], x[(i + 5) & 31], T[0], T[1], T[2]);
            else y[i] = op<B>(y[i], T[(i + 1) & 3], x[(i + 5) & 31], T[0], T[1], T[2]);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NR; i++) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int A, int NA, int B, int NB>
void run(const char* name, int nsm, const uint32_t* din, double clk_hz) {
    for (int c : {2, 4}) {
        uint32_t* out;
        const int grid = nsm * c;
        CK(cudaMalloc(&out, sizeof(uint32_t) * size_t(grid) * 128));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        mix<A, NA, B, NB><<<grid, 128>>>(out, din); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        mix<A, NA, B, NB><<<grid, 128>>>(out, din);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double instr = double(ITER) * 2 * NR * c;      // warp-instructions per sub-partition (c warps each)
        printf("%-44s %d warps/SMSP: %5.2f clk per warp-instruction per SMSP\n", name, c, double(ms) * 1e-3 * clk_hz / instr);
        CK(cudaFree(out));
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double clk_hz = double(khz) * 1e3;
    uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x01230457u * (i + 3) | 0x00010001u;
    uint32_t* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("device %s, %d SMs, %.0f MHz\n", p.name, nsm, clk_hz / 1e6);
    run<ADD2, 1, NONE, 0>("VIADD.16x2", nsm, din, clk_hz);
    run<ADD2I, 1, NONE, 0>("VIADD.16x2 immediate", nsm, din, clk_hz);
    run<MIN2, 1, NONE, 0>("VIMNMX.U16x2", nsm, din, clk_hz);
    run<ADDMIN2, 1, NONE, 0>("VIADDMNMX.U16x2", nsm, din, clk_hz);
    run<MIN3, 1, NONE, 0>("VIMNMX3.U16x2", nsm, din, clk_hz);
    run<LOP, 1, NONE, 0>("LOP3", nsm, din, clk_hz);
    run<PRM, 1, NONE, 0>("PRMT", nsm, din, clk_hz);
    run<IADD, 1, NONE, 0>("IADD3", nsm, din, clk_hz);
    run<IMADK, 1, NONE, 0>("IMAD", nsm, din, clk_hz);
    run<ADD2, 1, MIN2, 1>("VIADD.16x2 : VIMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<ADD2, 2, MIN2, 1>("VIADD.16x2 : VIMNMX.U16x2 = 2:1", nsm, din, clk_hz);
    run<ADD2, 1, ADDMIN2, 1>("VIADD.16x2 : VIADDMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<ADD2, 1, ADDMIN2, 3>("VIADD.16x2 : VIADDMNMX.U16x2 = 1:3", nsm, din, clk_hz);
    run<ADD2, 3, ADDMIN2, 1>("VIADD.16x2 : VIADDMNMX.U16x2 = 3:1", nsm, din, clk_hz);
    run<LOP, 1, ADDMIN2, 1>("LOP3 : VIADDMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<LOP, 1, ADD2, 1>("LOP3 : VIADD.16x2 = 1:1", nsm, din, clk_hz);
    run<IADD, 1, ADDMIN2, 1>("IADD3 : VIADDMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<IMADK, 1, ADDMIN2, 1>("IMAD : VIADDMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<IMADK, 1, ADD2, 1>("IMAD : VIADD.16x2 = 1:1", nsm, din, clk_hz);
    run<IMADK, 1, MIN2, 1>("IMAD : VIMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<PRM, 1, ADD2, 1>("PRMT : VIADD.16x2 = 1:1", nsm, din, clk_hz);
    run<PRM, 1, ADDMIN2, 1>("PRMT : VIADDMNMX.U16x2 = 1:1", nsm, din, clk_hz);
    run<MIN3, 1, ADD2, 1>("VIMNMX3.U16x2 : VIADD.16x2 = 1:1", nsm, din, clk_hz);
    run<MIN3, 1, ADD2, 2>("VIMNMX3.U16x2 : VIADD.16x2 = 1:2", nsm, din, clk_hz);
    printf("-- one register-file read per instruction (all other operands repeat: operand reuse cache) --\n");
    run<ADD2R, 1, MIN2R, 1>("VIADD.16x2 : VIMNMX.U16x2 = 1:1, 1 read each", nsm, din, clk_hz);
    run<ADD2R, 1, ADDMIN2R, 1>("VIADD.16x2 : VIADDMNMX.U16x2 = 1:1, 1 read each", nsm, din, clk_hz);
    run<ADD2R, 1, NONE, 0>("VIADD.16x2, 1 read", nsm, din, clk_hz);
    run<ADDMIN2R, 1, NONE, 0>("VIADDMNMX.U16x2, 1 read", nsm, din, clk_hz);
    return 0;
}
