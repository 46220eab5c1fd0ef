// Micro-benchmark v6 (round 2): does the survivor-history butterfly speed up with MORE WARPS when a frame pair is spread over
// T lanes (NQ = 64 / T registers per lane, in place)?  hist_mix2 showed 6.4 clk per butterfly per sub-partition with 64 registers per
// lane at 1 - 6 warps per scheduler (register-file bound?); here the lanes are thin (16 or 32 registers) and the warps many.
//   * local step: NQ/2 in-place butterflies (2 VIADD.16x2 + 2 VIADDMNMX.U16x2 each) + 4 tag adds
//   * cross step: per register VIADD.16x2 + SHFL.BFLY + VIADDMNMX.U16x2
//   * mix: the real cycle of K = 7: (6 - LT) local steps, LT cross steps
//   * table from shared memory (one LDS.128 per step) instead of registers that stay
// value = clocks per butterfly-equivalent (4 add-compare-selects of two frames) per sub-partition, from CUDA events at 1965 MHz.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int ITER = 2048 * 3;      // steps per launch (multiple of 6)

template <int NQ, int BIT>
__device__ __forceinline__ void local_step(uint32_t (&x)[NQ], const uint32_t (&T)[4], const uint32_t (&TT)[4]) {
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        if ((q & BIT) == 0) {
            const int q0 = q, q1 = q | BIT, pi = (q * 5 + (q >> 3)) & 3;
            const uint32_t b0 = __vadd2(x[q1], TT[pi ^ 3]), b1 = __vadd2(x[q1], TT[pi]);
            const uint32_t y0 = __viaddmin_u16x2(x[q0], T[pi], b0), y1 = __viaddmin_u16x2(x[q0], T[pi ^ 3], b1);
            x[q0] = y0; x[q1] = y1;
        }
    }
}

template <int NQ>
__device__ __forceinline__ void cross_step(uint32_t (&x)[NQ], const uint32_t (&S)[4], int lanemask) {
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int pi = (q * 5 + (q >> 3)) & 3;
        const uint32_t s = __vadd2(x[q], S[pi ^ 3]);
        const uint32_t r = __shfl_xor_sync(0xffffffffu, s, lanemask);
        x[q] = __viaddmin_u16x2(x[q], S[pi], r);
    }
}

// MODE 0: local steps only; 1: real cycle with cross steps; 2: real cycle, table via LDS.128 from shared memory
template <int NQ, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) lanes(uint32_t* out, const uint32_t* in) {
    constexpr int LT = (NQ == 16) ? 2 : (NQ == 32 ? 1 : 0);
    __shared__ uint4 tab[THREADS / 32][8][8];
    uint32_t x[NQ], T[4], TT[4];
#pragma unroll
    for (int i = 0; i < NQ; i++) x[i] = (in[i & 31] + threadIdx.x * (i + 1)) & 0xff00ff00u;
#pragma unroll
    for (int i = 0; i < 4; i++) T[i] = in[32 + i] & 0x0f000f00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) tab[warp][k][lane] = make_uint4(T[0] + k, T[1] + lane, T[2], T[3]);
    }
    __syncthreads();
    uint32_t tag = 0x00010001u;
#pragma unroll 1
    for (int it = 0; it < ITER / 6; it++) {
        auto table = [&](int k) {
            if (MODE == 2) {
                const uint4 v = tab[warp][(it + k) & 7][lane >> 2];
                T[0] = v.x; T[1] = v.y; T[2] = v.z; T[3] = v.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) T[i] = __vadd2(T[i], 0x01000100u);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) TT[i] = __vadd2(T[i], tag);
            tag = (tag << 1) | (tag >> 7);
            tag &= 0x00ff00ffu;
        };
        if constexpr (MODE == 0) {
            table(0); local_step<NQ, NQ / 2>(x, T, TT);
            table(1); local_step<NQ, NQ / 4>(x, T, TT);
            table(2); local_step<NQ, NQ / 8>(x, T, TT);
            table(3); local_step<NQ, NQ / 16>(x, T, TT);
            table(4); local_step<NQ, NQ / 2>(x, T, TT);
            table(5); local_step<NQ, NQ / 4>(x, T, TT);
        } else {
            table(0); local_step<NQ, NQ / 2>(x, T, TT);
            table(1); local_step<NQ, NQ / 4>(x, T, TT);
            table(2); local_step<NQ, NQ / 8>(x, T, TT);
            table(3); local_step<NQ, NQ / 16>(x, T, TT);
            if constexpr (LT == 2) {
                table(4); cross_step<NQ>(x, TT, 2);
                table(5); cross_step<NQ>(x, TT, 1);
            } else if constexpr (LT == 1) {
                table(4); local_step<NQ, NQ / 32 ? NQ / 32 : 1>(x, T, TT);
                table(5); cross_step<NQ>(x, TT, 1);
            } else {
                table(4); local_step<NQ, NQ / 32>(x, T, TT);
                table(5); local_step<NQ, NQ / 64>(x, T, TT);
            }
        }
        if ((it & 3) == 3) {
#pragma unroll
            for (int i = 0; i < NQ; i++) x[i] &= 0xff00ff00u;
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NQ; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NQ, int MODE, int THREADS>
void run_one(const char* name, int ctas_per_sm, int nsm, const uint32_t* din, double clk_hz) {
    uint32_t* out;
    const int grid = nsm * ctas_per_sm;
    CK(cudaMalloc(&out, sizeof(uint32_t) * size_t(grid) * THREADS));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    lanes<NQ, MODE, THREADS><<<grid, THREADS>>>(out, din); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    lanes<NQ, MODE, THREADS><<<grid, THREADS>>>(out, din);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, lanes<NQ, MODE, THREADS>));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lanes<NQ, MODE, THREADS>, THREADS, 0));
    const double wps = double(ctas_per_sm) * (THREADS / 32) / 4.0;               // warps per sub-partition
    const double bfly = double(ITER) * (NQ / 2) * wps;                            // butterfly-equivalents per sub-partition
    printf("%-44s NQ %2d  %5.2f warps/SMSP (occ limit %d CTAs): %5.2f clk/bfly/SMSP  regs %d spill %zu B\n", name, NQ, wps, occ,
           double(ms) * 1e-3 * clk_hz / bfly, fa.numRegs, fa.localSizeBytes);
    CK(cudaFree(out));
}

template <int NQ, int MODE>
void sweep(const char* name, int nsm, const uint32_t* din, double clk_hz) {
    for (int c : {1, 2, 3, 4, 6, 7, 8, 10, 12}) {
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lanes<NQ, MODE, 128>, 128, 0));
        if (c <= occ) run_one<NQ, MODE, 128>(name, c, nsm, din, clk_hz);
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double clk_hz = double(khz) * 1e3;
    uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x01230457u * (i + 3) | 0x00010001u;
    uint32_t* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("device %s, %d SMs, %.0f MHz; CTAs of 4 warps, c CTAs per SM = c warps per sub-partition\n", p.name, nsm, clk_hz / 1e6);
    sweep<64, 0>("T1 (64 regs) local steps only", nsm, din, clk_hz);
    sweep<32, 0>("T2 (32 regs) local steps only", nsm, din, clk_hz);
    sweep<16, 0>("T4 (16 regs) local steps only", nsm, din, clk_hz);
    sweep<32, 1>("T2 real cycle (5 local + 1 cross)", nsm, din, clk_hz);
    sweep<16, 1>("T4 real cycle (4 local + 2 cross)", nsm, din, clk_hz);
    sweep<16, 2>("T4 real cycle, table by LDS.128", nsm, din, clk_hz);
    sweep<32, 2>("T2 real cycle, table by LDS.128", nsm, din, clk_hz);
    return 0;
}
