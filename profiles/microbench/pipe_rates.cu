// Micro-benchmark: issue/pipe throughput of the sm_100a instructions the ACS kernels rely on.
// Reports warp-instructions per clock per SM sub-partition (SMSP) for each op mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int CH = 16;     // independent chains per thread
constexpr int ITER = 2048; // loop iterations

enum Op { VIADD2, VMIN2, VMIN2P_FADD, IADD, LOP, IMAD, FADD, PFADD, PIADD, MIX_ACS, MIX_ACS_PFADD, MIX_ACS_PIADD, MIX_ACS_PLOP,
          VIADDMIN, SHFL, VOTE, PRMT, ABSDIFF, MIX_VIADD_FADD, MIX_VIADD_IMAD, VMIN3, SEL, MIX_ACS_PIMAD, NOPS };
const char* names[] = {"VIADD.16x2", "VIMNMX.U16x2", "VIMNMX.U16x2+2pred -> 2x@P FADD (3 instr)", "IADD3", "LOP3", "IMAD", "FADD", "ISETP + @P FADD (2 instr)", "ISETP + @P IADD (2 instr)",
    "ACS 2xVIADD+1xVIMNMX (3 instr)", "ACS + 2x@P FADD (5 instr)", "ACS + 2x@P IADD (5 instr)", "ACS + 2x@P LOP3.OR (5 instr)",
    "VIADDMNMX.U16x2", "SHFL.BFLY", "VOTE.ANY", "PRMT", "VABSDIFF", "VIADD.16x2 + FADD (2 instr)", "VIADD.16x2 + IMAD (2 instr)", "VIMNMX3.U16x2", "ISETP + SEL (2 instr)",
    "ACS + 2x@P IMAD (5 instr)"};
const int instr_per_body[] = {1,1,3,1,1,1,1,2,2,3,5,5,5,1,1,1,1,1,2,2,1,2,5};

template<int OP>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t x[CH], y[CH];
    float f[CH];
    #pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = seed * (threadIdx.x + i * 77 + 1); y[i] = x[i] ^ 0x5a5a1234u; f[i] = 8388608.f; }
    uint32_t k1 = seed | 0x00010001u, k2 = seed ^ 0x00030002u;
    uint32_t one = (seed >> 20) | 1;  // opaque 1
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == VIADD2) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(k1));
            if (OP == VMIN2)  asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == VMIN2P_FADD) {
                bool ph, pl; x[i] = __vibmin_u16x2(x[i] ^ k1, y[i], &ph, &pl);
                if (!pl) f[i] += 1.0f; if (!ph) f[(i+1)%CH] += 2.0f;
            }
            if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(k1));
            if (OP == LOP)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(k1), "r"(k2));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(k1), "r"(k2));
            if (OP == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0f));
            if (OP == PFADD) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; @p add.f32 %0, %0, 0f3F800000;}" : "+f"(f[i]) : "r"(k1));
            if (OP == PIADD) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; @p add.u32 %0, %0, 5;}" : "+r"(x[i]) : "r"(k1));
            if (OP == MIX_ACS || OP == MIX_ACS_PFADD || OP == MIX_ACS_PIADD || OP == MIX_ACS_PLOP || OP == MIX_ACS_PIMAD) {
                uint32_t a = __vadd2(x[i], k1), b = __vadd2(y[i], k2);
                bool ph, pl; uint32_t m = __vibmin_u16x2(a, b, &ph, &pl);
                y[i] = x[i]; x[i] = m;
                if (OP == MIX_ACS_PFADD) { if (!pl) f[i] += 1.0f; if (!ph) f[(i + CH/2) % CH] += 2.0f; }
                if (OP == MIX_ACS_PIADD) { if (!pl) asm volatile("add.u32 %0, %0, 4;" : "+r"(out[0]) :: ); }
            }
            if (OP == VIADDMIN) x[i] = __viaddmin_u16x2(x[i], k1, y[i]);
            if (OP == SHFL) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
            if (OP == VOTE) x[i] = __ballot_sync(0xffffffffu, x[i] & (1u << (threadIdx.x & 31)));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == ABSDIFF) x[i] = __sad((int)x[i], (int)k1, 0u);
            if (OP == MIX_VIADD_FADD) { asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(k1)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(1.0f)); }
            if (OP == MIX_VIADD_IMAD) { asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(k1)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(k1), "r"(k2)); }
            if (OP == VMIN3) x[i] = __vimin3_u16x2(x[i], y[i], k1);
            if (OP == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(x[i]) : "r"(y[i]), "r"(k1));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0; float facc = 0;
    #pragma unroll
    for (int i = 0; i < CH; i++) { acc ^= x[i] ^ y[i]; facc += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(facc);
    if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

// Separate kernels for the predicated-integer variants where the decision word is a per-thread register
template<int MODE>
__global__ void __launch_bounds__(1024) bench_acs_dec(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t x[CH], y[CH];
    uint32_t d[4] = {0, 0, 0, 0};
    float f[4] = {8388608.f, 8388608.f, 8388608.f, 8388608.f};
    #pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = seed * (threadIdx.x + i * 77 + 1); y[i] = x[i] ^ 0x5a5a1234u; }
    uint32_t k1 = seed | 0x00010001u, k2 = seed ^ 0x00030002u;
    uint32_t one = (seed >> 20) | 1;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int i = 0; i < CH; i++) {
            uint32_t a = __vadd2(x[i], k1), b = __vadd2(y[i], k2);
            bool ph, pl; uint32_t m = __vibmin_u16x2(a, b, &ph, &pl);
            y[i] = x[i]; x[i] = m;
            if (MODE == 0) { if (!pl) d[i & 1] += (1u << i); if (!ph) d[2 + (i & 1)] += (1u << i); }           // compiler's choice of int add
            if (MODE == 1) { if (!pl) d[i & 1] |= (1u << i); if (!ph) d[2 + (i & 1)] |= (1u << i); }           // or
            if (MODE == 2) { if (!pl) f[i & 1] += float(1u << i); if (!ph) f[2 + (i & 1)] += float(1u << i); } // fadd
            if (MODE == 3) {  // forced IMAD: d = one * imm + d
                if (!pl) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[i & 1]) : "r"(one), "r"(1u << i));
                if (!ph) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[2 + (i & 1)]) : "r"(one), "r"(1u << i));
            }
        }
        if (MODE == 2) { d[0] ^= __float_as_uint(f[0]); d[1] ^= __float_as_uint(f[1]); d[2] ^= __float_as_uint(f[2]); d[3] ^= __float_as_uint(f[3]);
                         f[0] = f[1] = f[2] = f[3] = 8388608.f; }
    }
    long long t1 = clock64();
    uint32_t acc = d[0] ^ d[1] ^ d[2] ^ d[3];
    #pragma unroll
    for (int i = 0; i < CH; i++) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template<typename F>
void run(const char* name, int ipb, F launch, int nsm) {
    const int warps_per_smsp_list[] = {1, 2, 4, 8};
    printf("%-46s", name);
    for (int wps : warps_per_smsp_list) {
        int threads = 32 * 4 * wps;
        uint32_t* out; long long* cyc;
        int nw = nsm * 4 * wps;
        CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * threads));
        CK(cudaMalloc(&cyc, sizeof(long long) * nw));
        launch(nsm, threads, out, cyc);  // warm
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        launch(nsm, threads, out, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(nw);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        double med = (double)h[nw / 2];
        double rate = (double)wps * ITER * CH * ipb / med;  // warp-instr / clk / SMSP
        printf("  w%d: %.3f", wps, rate);
        if (wps == 8) printf("   (%.0f MHz eff)", med / (ms * 1e3));
        CK(cudaFree(out)); CK(cudaFree(cyc));
    }
    printf("\n");
}

#define RUN(OP) run(names[OP], instr_per_body[OP], [](int g, int t, uint32_t* o, long long* c) { bench<OP><<<g, t>>>(o, c, 12345u); }, nsm)
#define RUND(MODE, NAME) run(NAME, 5, [](int g, int t, uint32_t* o, long long* c) { bench_acs_dec<MODE><<<g, t>>>(o, c, 12345u); }, nsm)

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
    printf("columns: warp-instructions / clock / SMSP at 1,2,4,8 warps per SMSP (16 independent chains per thread)\n");
    RUN(VIADD2); RUN(VMIN2); RUN(VMIN3); RUN(VIADDMIN); RUN(IADD); RUN(LOP); RUN(IMAD); RUN(FADD); RUN(PFADD); RUN(PIADD); RUN(SEL); RUN(PRMT); RUN(ABSDIFF);
    RUN(SHFL); RUN(VOTE);
    RUN(MIX_VIADD_FADD); RUN(MIX_VIADD_IMAD);
    RUN(VMIN2P_FADD); RUN(MIX_ACS); RUN(MIX_ACS_PFADD);
    RUND(0, "ACS + 2x@P int add (compiler choice, 5 instr)");
    RUND(1, "ACS + 2x@P int or  (compiler choice, 5 instr)");
    RUND(2, "ACS + 2x@P FADD into 4 accumulators (5 instr)");
    RUND(3, "ACS + 2x@P IMAD forced (5 instr)");
    return 0;
}
