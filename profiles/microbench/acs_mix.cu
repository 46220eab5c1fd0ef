// Micro-benchmark v3: sustained throughput of the in-place ACS pass (16 packed registers = 8 butterflies per step, 4 phases)
// with different ways of turning the VIMNMX predicates into decision bits.  Timing = whole-CTA duration (one CTA per SM,
// clock64 between two __syncthreads), so warp-scheduling order cannot inflate the number.
// Reports clocks per butterfly per SM sub-partition (lower is better; 4 ACS of two frames = 1 butterfly) and instr/clk.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
constexpr int ITER = 2048;   // steps

// DEC: 0 none, 1 @P FADD imm, 2 @P int add (compiler's choice of IADD3/VIADD/IMAD.IADD), 3 @P LOP3 or, 4 @P HADD2 (half2 accumulators),
//      5 half FADD / half int add, 6 @P FADD with register operand
// min + both decision bits deposited by predicated DADDs, in one PTX block (ptxas folds min+setp into VIMNMX with predicate outputs)
__device__ __forceinline__ uint32_t min_dadd2(uint32_t a, uint32_t b, double& dlo, double& dhi, double w) {
    uint32_t m;
    asm("{ .reg .pred pu, pv; .reg .u16 r0, r1, r2, r3;\n\t"
        "min.u16x2 %0, %3, %4;\n\t"
        "mov.b32 {r0, r1}, %0;\n\t"
        "mov.b32 {r2, r3}, %3;\n\t"
        "setp.ne.u16 pv, r0, r2;\n\t"
        "setp.ne.u16 pu, r1, r3;\n\t"
        "@pv add.f64 %1, %1, %5;\n\t"
        "@pu add.f64 %2, %2, %5; }"
        : "=r"(m), "+d"(dlo), "+d"(dhi) : "r"(a), "r"(b), "d"(w));
    return m;
}
__device__ __forceinline__ uint32_t min_dadd_fadd(uint32_t a, uint32_t b, double& dlo, float& fhi, double w, float wf) {
    uint32_t m;
    asm("{ .reg .pred pu, pv; .reg .u16 r0, r1, r2, r3;\n\t"
        "min.u16x2 %0, %3, %4;\n\t"
        "mov.b32 {r0, r1}, %0;\n\t"
        "mov.b32 {r2, r3}, %3;\n\t"
        "setp.ne.u16 pv, r0, r2;\n\t"
        "setp.ne.u16 pu, r1, r3;\n\t"
        "@pv add.f64 %1, %1, %5;\n\t"
        "@pu add.f32 %2, %2, %6; }"
        : "=r"(m), "+d"(dlo), "+f"(fhi) : "r"(a), "r"(b), "d"(w), "f"(wf));
    return m;
}

template<int DEC>
__global__ void __launch_bounds__(1024) acs(uint32_t* out, long long* cycles, const uint32_t* in) {
    uint32_t x[16]; uint32_t T[4], I[4];
    float fa[4] = {8388608.f, 8388608.f, 8388608.f, 8388608.f};
    uint32_t da[4] = {0, 0, 0, 0};
    uint32_t ha[4] = {0, 0, 0, 0};
    float fw[4];
    double dd[4] = {4503599627370496.0, 4503599627370496.0, 4503599627370496.0, 4503599627370496.0};   // 2^52
    #pragma unroll
    for (int i = 0; i < 16; i++) x[i] = in[i] + threadIdx.x;
    #pragma unroll
    for (int i = 0; i < 4; i++) { T[i] = in[16 + i]; I[i] = in[20 + i]; fw[i] = __uint_as_float(in[24 + i]); }
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITER / 4; it++) {
        #pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            const int bit = 8 >> ph;
            #pragma unroll
            for (int q = 0; q < 16; q++) {
                if (q & bit) continue;
                uint32_t &x0 = x[q], &x1 = x[q | bit];
                const uint32_t t = T[(q * 5 + ph) & 3], iv = I[(q * 5 + ph) & 3];
                const uint32_t a0 = __vadd2(x0, t), b0 = __vadd2(x1, iv), a1 = __vadd2(x0, iv), b1 = __vadd2(x1, t);
                bool p0h = false, p0l = false, p1h = false, p1l = false;
                const int bp = (q & 7) * 2;
                if (DEC == 16) { x0 = min_dadd2(a0, b0, dd[0], dd[1], double(1u << bp)); x1 = min_dadd2(a1, b1, dd[2], dd[3], double(2u << bp)); }
                else if (DEC == 17) { x0 = min_dadd_fadd(a0, b0, dd[0], fa[1], double(1u << bp), float(1u << bp));
                                      x1 = min_dadd_fadd(a1, b1, dd[2], fa[3], double(2u << bp), float(2u << bp)); }
                else {
                x0 = __vibmin_u16x2(a0, b0, &p0h, &p0l);
                x1 = __vibmin_u16x2(a1, b1, &p1h, &p1l);
                }
                if (DEC == 1) { if (!p0l) fa[0] += float(1u << bp); if (!p0h) fa[1] += float(1u << bp);
                                if (!p1l) fa[2] += float(2u << bp); if (!p1h) fa[3] += float(2u << bp); }
                if (DEC == 2) { if (!p0l) da[0] += (1u << bp); if (!p0h) da[1] += (1u << bp);
                                if (!p1l) da[2] += (2u << bp); if (!p1h) da[3] += (2u << bp); }
                if (DEC == 3) { if (!p0l) da[0] |= (1u << bp); if (!p0h) da[1] |= (1u << bp);
                                if (!p1l) da[2] |= (2u << bp); if (!p1h) da[3] |= (2u << bp); }
                if (DEC == 4) { // half2 accumulators: one HADD2 per predicate, constant = (2^k, 0) or (0, 2^k) pattern in fp16 (k < 11)
                                if (!p0l) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(ha[0]) : "r"(0x00003c00u + (bp << 10)));
                                if (!p0h) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(ha[1]) : "r"(0x00003c00u + (bp << 10)));
                                if (!p1l) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(ha[2]) : "r"(0x3c000000u + (bp << 26)));
                                if (!p1h) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(ha[3]) : "r"(0x3c000000u + (bp << 26))); }
                if (DEC == 5) { if (!p0l) fa[0] += float(1u << bp); if (!p0h) da[1] += (1u << bp);
                                if (!p1l) fa[2] += float(2u << bp); if (!p1h) da[3] += (2u << bp); }
                if (DEC == 7) { if (!p0l) fa[0] += float(1u << bp); }
                if (DEC == 8) { if (!p0l) fa[0] += float(1u << bp); if (!p1l) fa[2] += float(2u << bp); }
                if (DEC == 9) { if (!p0l) fa[0] += float(1u << bp); if (!p0h) fa[1] += float(1u << bp); }
                if (DEC == 10) { // SEL + 3-input add: 2 SEL + 1 IADD3 per VIMNMX
                                da[0] += (p0l ? 0u : (1u << bp)) + (p0h ? 0u : (0x10000u << bp));
                                da[1] += (p1l ? 0u : (2u << bp)) + (p1h ? 0u : (0x20000u << bp)); }
                if (DEC == 11) { if (!p0l) dd[0] += double(1u << bp); if (!p0h) dd[1] += double(1u << bp);
                                 if (!p1l) dd[2] += double(2u << bp); if (!p1h) dd[3] += double(2u << bp); }
                if (DEC == 12) { if (!p0l) dd[0] += double(1u << bp); if (!p0h) dd[1] += double(1u << bp);
                                 if (!p1l) dd[2] += double(2u << bp); if (!p1h) fa[3] += float(2u << bp); }
                if (DEC == 13) { if (!p0l) dd[0] += double(1u << bp); if (!p0h) fa[1] += float(1u << bp);
                                 if (!p1l) dd[2] += double(2u << bp); if (!p1h) fa[3] += float(2u << bp); }
                if (DEC == 14) { if (!p0l) dd[0] += double(1u << bp); if (!p0h) dd[1] += double(1u << bp);
                                 if (!p1l) dd[2] += double(2u << bp); if (!p1h) da[3] += (2u << bp); }
                if (DEC == 15) { if (!p0l) dd[0] += double(1u << bp); if (!p0h) da[1] += (1u << bp);
                                 if (!p1l) dd[2] += double(2u << bp); if (!p1h) fa[3] += float(2u << bp); }
                if (DEC == 6) { if (!p0l) fa[0] += fw[0]; if (!p0h) fa[1] += fw[1];
                                if (!p1l) fa[2] += fw[2]; if (!p1h) fa[3] += fw[3]; }
            }
            if (DEC == 1 || (DEC >= 5 && DEC <= 9) || DEC >= 12) { da[0] ^= __float_as_uint(fa[0]) ^ __float_as_uint(fa[1]); da[1] += __float_as_uint(fa[2]) ^ __float_as_uint(fa[3]);
                            fa[0] = fa[1] = fa[2] = fa[3] = 8388608.f; }
            if (DEC >= 11) { da[2] ^= uint32_t(__double_as_longlong(dd[0])) ^ uint32_t(__double_as_longlong(dd[1]));
                             da[3] += uint32_t(__double_as_longlong(dd[2])) ^ uint32_t(__double_as_longlong(dd[3]));
                             dd[0] = dd[1] = dd[2] = dd[3] = 4503599627370496.0; }
            if (DEC == 4) { da[0] ^= ha[0] ^ ha[1]; da[1] += ha[2] ^ ha[3]; ha[0] = ha[1] = ha[2] = ha[3] = 0; }
            T[ph] = __vadd2(T[ph], 0x00010001u);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    uint32_t acc = da[0] ^ da[1] ^ da[2] ^ da[3];
    #pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// tagged butterfly (uint8 metrics << 8): tag byte on the path-1 operand, VIADDMNMX, PRMT + shift-accumulate, LOP3 mask
// ACCMODE: 0 = acc*2 + t (compiler's choice), 1 = shift+or, 2 = no accumulate (only PRMT), 3 = no decisions at all (no mask either)
template<int ACCMODE>
__global__ void __launch_bounds__(1024) acs_tag(uint32_t* out, long long* cycles, const uint32_t* in) {
    uint32_t x[16]; uint32_t T[4], TT[4];
    uint32_t da[2] = {0, 0};
    #pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (in[i] + threadIdx.x) & 0xff00ff00u;
    #pragma unroll
    for (int i = 0; i < 4; i++) { T[i] = in[16 + i] & 0xff00ff00u; TT[i] = T[i] + 0x00010001u; }
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITER / 4; it++) {
        #pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            const int bit = 8 >> ph;
            #pragma unroll
            for (int q = 0; q < 16; q++) {
                if (q & bit) continue;
                uint32_t &x0 = x[q], &x1 = x[q | bit];
                const int pi = (q * 5 + ph) & 3;
                const uint32_t b0 = __vadd2(x1, TT[pi ^ 3]), b1 = __vadd2(x1, TT[pi]);
                const uint32_t m0 = __viaddmin_u16x2(x0, T[pi], b0), m1 = __viaddmin_u16x2(x0, T[pi ^ 3], b1);
                if (ACCMODE == 0) da[q & 1] = da[q & 1] + da[q & 1] + __byte_perm(m0, m1, 0x6420);
                if (ACCMODE == 1) da[q & 1] = (da[q & 1] << 1) | __byte_perm(m0, m1, 0x6420);
                if (ACCMODE == 2) da[q & 1] ^= __byte_perm(m0, m1, 0x6420);
                if (ACCMODE == 3) { x0 = m0; x1 = m1; } else { x0 = m0 & 0xff00ff00u; x1 = m1 & 0xff00ff00u; }
            }
            T[ph] = __vadd2(T[ph], 0x01000100u); TT[ph] = __vadd2(TT[ph], 0x01000100u);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    uint32_t acc = da[0] ^ da[1];
    #pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<typename F>
void run(const char* name, int instr_per_bfly, F launch, int nsm, const uint32_t* din) {
    const int wl[] = {1, 2, 3, 4, 6, 8};
    printf("%-44s", name);
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
        double clk_per_bfly = med / (double(ITER) * 8 * wps);     // per SMSP: wps warps, 8 butterflies per step each
        printf("  w%d: %5.2f (%.2f)", wps, clk_per_bfly, instr_per_bfly / clk_per_bfly);
        CK(cudaFree(out)); CK(cudaFree(cyc));
    }
    printf("\n");
}
#define ACS(DEC, IPB, NAME) run(NAME, IPB, [](int g, int t, uint32_t* o, long long* c, const uint32_t* in) { acs<DEC><<<g, t>>>(o, c, in); }, nsm, din)

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    uint32_t h[64]; for (int i = 0; i < 64; i++) h[i] = 0x01230457u * (i + 3) | 0x00010001u;
    for (int i = 24; i < 32; i++) { float v = float(1 << (i - 24)); h[i] = *reinterpret_cast<uint32_t*>(&v); }
    uint32_t* din; CK(cudaMalloc(&din, sizeof(h))); CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
    printf("device %s, %d SMs.  clocks per butterfly per SMSP (instr/clk in brackets) at w warps per SMSP\n", p.name, nsm);
    ACS(0, 6, "no decisions (4 VIADD.16x2 + 2 VIMNMX)");
    ACS(1, 10, "+ 4 @P FADD imm");
    ACS(6, 10, "+ 4 @P FADD reg");
    ACS(2, 10, "+ 4 @P int add (compiler mix)");
    ACS(3, 10, "+ 4 @P LOP3 or");
    ACS(4, 10, "+ 4 @P HADD2");
    ACS(5, 10, "+ 2 @P FADD + 2 @P int add");
    ACS(11, 10, "+ 4 @P DADD (fp64 pipe)");
    ACS(12, 10, "+ 3 @P DADD + 1 @P FADD");
    ACS(13, 10, "+ 2 @P DADD + 2 @P FADD");
    ACS(14, 10, "+ 3 @P DADD + 1 @P int add");
    ACS(15, 10, "+ 2 @P DADD + 1 FADD + 1 int add");
    ACS(16, 10, "+ 4 @P DADD via PTX (fp64 pipe)");
    ACS(17, 10, "+ 2 @P DADD + 2 @P FADD via PTX");
    ACS(7, 7, "+ 1 @P FADD (1 of 4 predicates used)");
    ACS(8, 8, "+ 2 @P FADD (lo pred of each VIMNMX)");
    ACS(9, 8, "+ 2 @P FADD (both preds of one VIMNMX)");
    ACS(10, 12, "+ 4 SEL + 2 IADD3 (3-input)");
#define TAG(M, IPB, NAME) run(NAME, IPB, [](int g, int t, uint32_t* o, long long* c, const uint32_t* in) { acs_tag<M><<<g, t>>>(o, c, in); }, nsm, din)
    TAG(3, 4, "tagged: 2 VIADD + 2 VIADDMNMX only");
    TAG(2, 7, "tagged: + 2 LOP3 mask + PRMT");
    TAG(0, 8, "tagged: + acc*2+t (compiler)");
    TAG(1, 9, "tagged: + shift | or");
    return 0;
}
