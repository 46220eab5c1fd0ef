// Micro-benchmark (round 2): how many shared-memory wavefronts does one LDS.128 cost when several lanes of a warp read the SAME 16-byte
// slot?  16 warps per SM (one 512-thread CTA, like acs_hist_cta_kernel), 8 independent LDS.128 per iteration, slot pattern per lane as
// listed in main() (ld.volatile: ptxas merges plain loads of one address, which made the first version of this test measure its ALU work).  Reports clocks per warp-level LDS.128 per SM (1 wavefront = 1 clock at best).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096, NLD = 8;

__global__ void __launch_bounds__(512, 1) k(uint32_t* out, const int* slot_of_lane, int vec) {
    __shared__ uint4 tbl[2][128];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) (&tbl[0][0])[i] = make_uint4(i, i * 3, i * 5, i * 7);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t addr = uint32_t(__cvta_generic_to_shared(&tbl[0][0])) + uint32_t(slot_of_lane[lane]) * 16u;
    uint32_t acc = 0;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
        uint32_t a[NLD], b[NLD], c[NLD], d[NLD];
        if (vec == 4) {
#pragma unroll
            for (int i = 0; i < NLD; i++)
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[i]), "=r"(b[i]), "=r"(c[i]), "=r"(d[i]) : "r"(addr ^ (uint32_t(i & 1) << 11)));
#pragma unroll
            for (int i = 0; i < NLD; i++) acc ^= a[i] ^ d[i];
        } else {
#pragma unroll
            for (int i = 0; i < NLD; i++)
                asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(a[i]), "=r"(b[i]) : "r"(addr ^ (uint32_t(i & 1) << 11)));
#pragma unroll
            for (int i = 0; i < NLD; i++) acc ^= a[i] ^ b[i];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const int nsm = p.multiProcessorCount;
    uint32_t* out; CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * 512));
    int* dslot; CK(cudaMalloc(&dslot, 32 * sizeof(int)));
    struct Pat { const char* name; int (*f)(int); };
    const Pat pats[] = {
        {"32 distinct slots, slot = lane", [](int l) { return l; }},
        {"16 distinct, slot = lane & 15", [](int l) { return l & 15; }},
        {"16 distinct, slot = lane >> 1", [](int l) { return l >> 1; }},
        {"16 distinct, scattered: slot = (lane*7 + lane/16) & 15", [](int l) { return (l * 7 + l / 16) & 15; }},
        {"8 distinct, slot = lane & 7 (each quarter warp: 8 slots)", [](int l) { return l & 7; }},
        {"8 distinct, slot = lane >> 2 (each quarter warp: 2 slots)", [](int l) { return l >> 2; }},
        {"8 distinct, scattered: slot = (lane*5 + lane/8) & 7", [](int l) { return (l * 5 + l / 8) & 7; }},
        {"8 distinct in ONE bank group: slot = (lane & 7) * 8", [](int l) { return (l & 7) * 8; }},
        {"8 distinct, slot = (lane & 7) + 8*(lane>>3 & 1)... 16 slots in 8 groups", [](int l) { return (l & 7) + 8 * ((l >> 3) & 1); }},
        {"4 distinct, slot = lane & 3", [](int l) { return l & 3; }},
        {"1 slot (broadcast)", [](int) { return 5; }},
    };
    printf("device %s, %d SMs, %.0f MHz\n", p.name, nsm, khz / 1e3);
    for (int vec : {4, 2})
        for (const Pat& pt : pats) {
            int h[32]; for (int l = 0; l < 32; l++) h[l] = pt.f(l);
            CK(cudaMemcpy(dslot, h, sizeof(h), cudaMemcpyHostToDevice));
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            k<<<nsm, 512>>>(out, dslot, vec); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); k<<<nsm, 512>>>(out, dslot, vec); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double loads = double(ITER) * NLD * 16;     // warp-level loads per SM
            printf("LDS.%d  %-72s %5.2f clk per warp load per SM\n", vec * 32, pt.name, double(ms) * 1e-3 * khz * 1e3 / loads);
        }
    return 0;
}
