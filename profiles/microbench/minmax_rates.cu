// Microbenchmark (B200): per-SM issue throughput of the integer min/max forms a median network can use:
// scalar IMNMX (min.s32), packed 16-bit min/max (DPX: __vimax_s16x2 / __vimin_u16x2), 3-input forms
// (__vimax3_s32, __vimax3_s16x2), and FMNMX for comparison.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ unsigned pmin_s(unsigned a, unsigned b) { unsigned d; asm("min.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned pmax_s(unsigned a, unsigned b) { unsigned d; asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned pmin_u(unsigned a, unsigned b) { unsigned d; asm("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned pmax_u(unsigned a, unsigned b) { unsigned d; asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

constexpr int ITERS = 2048;
constexpr int U = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(unsigned* out, long long* cyc, unsigned m0, unsigned m1) {
    unsigned u[U]; int s[U]; float f[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { u[i] = threadIdx.x * 2654435761u + i * 40503u; s[i] = (int)u[i]; f[i] = (float)u[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) s[i] = min(s[i], s[(i + 1) % U] + it);                 // IADD + IMNMX
            else if (MODE == 1) s[i] = max(min(s[i], s[(i + 1) % U]), (int)m0);   // 2 x IMNMX
            else if (MODE == 2) u[i] = pmax_s(pmin_s(u[i], u[(i + 1) % U]), m0);
            else if (MODE == 3) u[i] = pmax_u(pmin_u(u[i], u[(i + 1) % U]), m0);
            else if (MODE == 4) s[i] = __vimax3_s32(s[i], s[(i + 1) % U], (int)m0);
            else if (MODE == 5) u[i] = __vimin3_s16x2(u[i], u[(i + 1) % U], m1);
            else if (MODE == 6) f[i] = fmaxf(fminf(f[i], f[(i + 1) % U]), 3.0f);
            else if (MODE == 7) u[i] = __vmaxs2(__vmins2(u[i], u[(i + 1) % U]), m0);
            else if (MODE == 8) {   // half2 min / max on the same bit patterns (HMNMX2): which pipe, what rate?
                __half2 a = *reinterpret_cast<__half2*>(&u[i]), b = *reinterpret_cast<__half2*>(&u[(i + 1) % U]);
                __half2 c = *reinterpret_cast<const __half2*>(&m0);
                __half2 r = __hmax2(__hmin2(a, b), c);
                u[i] = *reinterpret_cast<unsigned*>(&r);
            } else if (MODE == 9) {  // mixed: one packed integer min + one half2 max per step (do the pipes overlap?)
                __half2 a = *reinterpret_cast<__half2*>(&u[i]), c = *reinterpret_cast<const __half2*>(&m0);
                __half2 r = __hmax2(a, c);
                u[i] = pmin_u(*reinterpret_cast<unsigned*>(&r), u[(i + 1) % U]);
            }
        }
    }
    long long t1 = clock64();
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) r += u[i] + (unsigned)s[i] + (unsigned)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 1024>>>(out, cyc, 0x00010002u, 0x7fff7fffu);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    double warp_instr = (double)ITERS * U * 32 * instr_per_iter;
    printf("%-44s %.2f warp-instr/clk/SM\n", name, warp_instr / avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IADD + IMNMX (2 instr)", 2);
    run<1>("IMNMX x2 (min.s32, max.s32)", 2);
    run<2>("min.s16x2 / max.s16x2 (PTX) x2", 2);
    run<3>("min/max .u16x2 x2", 2);
    run<4>("__vimax3_s32 (1 instr)", 1);
    run<5>("__vimin3_s16x2 (1 instr)", 1);
    run<6>("FMNMX x2", 2);
    run<7>("__vmins2/__vmaxs2 x2", 2);
    run<8>("__hmin2 / __hmax2 (HMNMX2) x2", 2);
    run<9>("min.u16x2 + __hmax2 (mixed) x2", 2);
    return 0;
}
