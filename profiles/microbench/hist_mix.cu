// Micro-benchmark v4: (1) issue rates of the fused add-min instructions (VIADDMNMX.U16x2 / .U32) alone and next to the adds they are
// paired with in the survivor-history butterfly (acs_hist.cuh); (2) sustained clocks per butterfly of the ping-pong K=7 step
// (64 metric registers -> 64 metric registers, 32 butterflies) in its packed u16x2 form (two frames per register, 8-bit history) and
// its 32-bit form (one frame per register, 16-bit history).  One CTA per SM, w warps per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int CH = 8;
constexpr int ITER = 4096;

enum Op { O_NONE, O_VIADD2, O_IADD3, O_AM16, O_AM32, O_MIN32, O_MIN16, O_LOP3 };

template<int OP> __device__ __forceinline__ void doop(uint32_t& x, uint32_t y, uint32_t c) {
    if (OP == O_VIADD2) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == O_IADD3)  asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == O_AM16)   x = __viaddmin_u16x2(x, c, y);
    if (OP == O_AM32)   x = __viaddmin_u32(x, c, y);
    if (OP == O_MIN32)  asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == O_MIN16)  asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == O_LOP3)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x1e;" : "+r"(x) : "r"(y), "r"(0x5a5a5a5au));
}
template<int OP> constexpr int icount() { return OP == O_NONE ? 0 : 1; }

template<int A, int B>
__global__ void __launch_bounds__(1024) pair(uint32_t* out, long long* cycles, const uint32_t* in) {
    uint32_t x[CH], y[CH], u[CH], v[CH];
    const uint32_t c = in[40];
    #pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = in[i] * (threadIdx.x + 1); y[i] = in[i + 8] ^ threadIdx.x; u[i] = in[16 + i] + threadIdx.x; v[i] = in[24 + i]; }
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 2
    for (int it = 0; it < ITER / 2; it++) {
        #pragma unroll
        for (int i = 0; i < CH; i++) { doop<A>(x[i], y[i], c); doop<B>(u[i], v[i], c); }
        #pragma unroll
        for (int i = 0; i < CH; i++) { doop<A>(y[i], x[i], c); doop<B>(v[i], u[i], c); }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int i = 0; i < CH; i++) acc ^= x[i] ^ y[i] ^ u[i] ^ v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cycles[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

// ping-pong K=7 step.  MODE 0: packed u16x2 (VIADD.16x2 + VIADDMNMX.U16x2); MODE 1: 32-bit (IADD + VIADDMNMX.U32);
// MODE 2: packed, min(x0 + d, x1) + v form (the fused instruction first, then the add)
template<int MODE>
__device__ __forceinline__ void step(const uint32_t (&x)[64], uint32_t (&y)[64], const uint32_t (&T)[4], const uint32_t (&TT)[4]) {
    #pragma unroll
    for (int J = 0; J < 32; J++) {
        const int pi = (J * 5 + (J >> 3)) & 3;
        if (MODE == 0) {
            const uint32_t b0 = __vadd2(x[J + 32], TT[pi ^ 3]), b1 = __vadd2(x[J + 32], TT[pi]);
            y[2 * J] = __viaddmin_u16x2(x[J], T[pi], b0);
            y[2 * J + 1] = __viaddmin_u16x2(x[J], T[pi ^ 3], b1);
        } else if (MODE == 1) {
            const uint32_t b0 = x[J + 32] + TT[pi ^ 3], b1 = x[J + 32] + TT[pi];
            y[2 * J] = __viaddmin_u32(x[J], T[pi], b0);
            y[2 * J + 1] = __viaddmin_u32(x[J], T[pi ^ 3], b1);
        } else {
            y[2 * J] = __vadd2(__viaddmin_u16x2(x[J], T[pi], x[J + 32]), TT[pi ^ 3]);
            y[2 * J + 1] = __vadd2(__viaddmin_u16x2(x[J], T[pi ^ 3], x[J + 32]), TT[pi]);
        }
    }
}

template<int MODE>
__global__ void __launch_bounds__(384) hist(uint32_t* out, long long* cycles, const uint32_t* in) {
    uint32_t x[64], y[64], T[4], TT[4];
    #pragma unroll
    for (int i = 0; i < 64; i++) x[i] = (in[i & 31] + threadIdx.x * (i + 1)) & (MODE == 1 ? 0xffff0000u : 0xff00ff00u);
    #pragma unroll
    for (int i = 0; i < 4; i++) { T[i] = in[32 + i] & (MODE == 1 ? 0x0fff0000u : 0x0f000f00u); }
    const uint32_t inc = MODE == 1 ? 0x00010000u : 0x01000100u;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITER / 8; it++) {
        uint32_t tag = MODE == 1 ? 1u : 0x00010001u;
        #pragma unroll 1
        for (int k = 0; k < 4; k++) {
            #pragma unroll
            for (int i = 0; i < 4; i++) { T[i] = MODE == 1 ? T[i] + inc : __vadd2(T[i], inc); TT[i] = T[i] + tag; }
            step<MODE>(x, y, T, TT);
            #pragma unroll
            for (int i = 0; i < 4; i++) { TT[i] = T[i] + (tag << 1); }
            step<MODE>(y, x, T, TT);
            tag <<= 2;
        }
        #pragma unroll
        for (int i = 0; i < 64; i++) x[i] &= (MODE == 1 ? 0xffffff00u : 0xff00ff00u);      // clear the (8 bits of) history
    }
    __syncthreads();
    long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int i = 0; i < 64; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<typename F>
void run_pair(const char* name, double instr_per_warp, F launch, int nsm, const uint32_t* din) {
    const int wl[] = {1, 2, 4, 8};
    printf("%-52s", name);
    for (int wps : wl) {
        int threads = 32 * 4 * wps, nw = nsm * 4 * wps;
        uint32_t* out; long long* cyc;
        CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * threads)); CK(cudaMalloc(&cyc, sizeof(long long) * nw));
        launch(nsm, threads, out, cyc, din); CK(cudaDeviceSynchronize());
        launch(nsm, threads, out, cyc, din); CK(cudaDeviceSynchronize());
        std::vector<long long> h(nw);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        double med = (double)h[nw / 2];
        printf("  w%d: %.3f", wps, wps * instr_per_warp / med);
        CK(cudaFree(out)); CK(cudaFree(cyc));
    }
    printf("\n");
}

template<typename F>
void run_hist(const char* name, F launch, int nsm, const uint32_t* din) {
    const int wl[] = {1, 2, 3};
    printf("%-60s", name);
    for (int wps : wl) {
        int threads = 32 * 4 * wps;
        uint32_t* out; long long* cyc;
        CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * threads)); CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
        launch(nsm, threads, out, cyc, din); CK(cudaDeviceSynchronize());
        launch(nsm, threads, out, cyc, din); CK(cudaDeviceSynchronize());
        std::vector<long long> h(nsm);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        double med = (double)h[nsm / 2];
        printf("  w%d: %5.2f", wps, med / (double(ITER) * 32 * wps));     // clocks per butterfly per SMSP
        CK(cudaFree(out)); CK(cudaFree(cyc));
    }
    printf("\n");
}

#define PAIR(A, B, NAME) run_pair(NAME, double(ITER) * CH * (icount<A>() + icount<B>()), [](int g, int t, uint32_t* o, long long* c, const uint32_t* in) { pair<A, B><<<g, t>>>(o, c, in); }, nsm, din)
#define HIST(M, NAME) run_hist(NAME, [](int g, int t, uint32_t* o, long long* c, const uint32_t* in) { hist<M><<<g, t>>>(o, c, in); }, nsm, din)

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x01230457u * (i + 3) | 0x00010001u;
    uint32_t* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("device %s, %d SMs. value = warp-instr/clk/SMSP\n", p.name, nsm);
    PAIR(O_AM16, O_NONE, "VIADDMNMX.U16x2");
    PAIR(O_AM32, O_NONE, "VIADDMNMX.U32");
    PAIR(O_MIN32, O_NONE, "VIMNMX.U32");
    PAIR(O_MIN16, O_NONE, "VIMNMX.U16x2");
    PAIR(O_AM16, O_VIADD2, "VIADDMNMX.U16x2 | VIADD.16x2");
    PAIR(O_AM16, O_IADD3, "VIADDMNMX.U16x2 | IADD3");
    PAIR(O_AM16, O_LOP3, "VIADDMNMX.U16x2 | LOP3");
    PAIR(O_AM16, O_MIN16, "VIADDMNMX.U16x2 | VIMNMX.U16x2");
    PAIR(O_AM32, O_IADD3, "VIADDMNMX.U32 | IADD3");
    PAIR(O_AM32, O_VIADD2, "VIADDMNMX.U32 | VIADD.16x2");
    PAIR(O_AM32, O_LOP3, "VIADDMNMX.U32 | LOP3");
    PAIR(O_MIN32, O_IADD3, "VIMNMX.U32 | IADD3");
    printf("ping-pong K=7 survivor-history step: clocks per butterfly per SMSP at w warps per SMSP (4 instructions per butterfly)\n");
    HIST(0, "packed u16x2: 2 VIADD.16x2 + 2 VIADDMNMX.U16x2");
    HIST(2, "packed u16x2: 2 VIADDMNMX.U16x2 (x0+d vs x1) + 2 VIADD.16x2");
    HIST(1, "32-bit: 2 IADD + 2 VIADDMNMX.U32 (one frame per register)");
    return 0;
}
