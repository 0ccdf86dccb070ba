// Microbenchmark (B200): per-SM issue throughput of the instruction mix the enhancement kernels use:
// FFMA (register and constant-operand forms), FFMA2 (fma.rn.f32x2), FADD, FMNMX, PRMT, LOP3, IADD3,
// IMAD, FADD.RM, SEL, LDS.128.  Reports lane-ops per clock per SM (128 = one warp instruction per
// scheduler per clock).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int U = 8;

template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float c0, float c1, unsigned m) {
    float a[U]; float2 p[U]; unsigned u[U];
#pragma unroll
    for (int i = 0; i < U; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = make_float2(a[i], a[i] + 1.f); u[i] = threadIdx.x * 2654435761u + i; }
    __shared__ float4 sm[1024];
    sm[threadIdx.x] = make_float4(c0, c1, c0, c1);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            if (MODE == 0) a[i] = __fmaf_rn(a[i], c0, c1);                       // FFMA reg,reg(uniform),reg
            else if (MODE == 1) a[i] = __fmaf_rn(a[i], a[(i + 1) % U], a[(i + 2) % U]);  // 3 distinct registers
            else if (MODE == 2) p[i] = __ffma2_rn(p[i], make_float2(c0, c0), make_float2(c1, c1));
            else if (MODE == 3) a[i] = __fadd_rn(a[i], c0);
            else if (MODE == 4) a[i] = fminf(a[i], c0 + i);
            else if (MODE == 5) u[i] = __byte_perm(u[i], m, 0x7610 + i);
            else if (MODE == 6) u[i] = (u[i] & m) ^ (u[(i + 1) % U]);
            else if (MODE == 7) u[i] = u[i] + m + u[(i + 1) % U];
            else if (MODE == 8) u[i] = u[i] * m + u[(i + 1) % U];
            else if (MODE == 9) a[i] = __fadd_rd(a[i], c0);
            else if (MODE == 10) a[i] = a[i] > c0 ? a[(i + 1) % U] : c1;
            else if (MODE == 11) { float4 v = sm[(threadIdx.x + i * 32 + it) & 1023]; a[i] += v.x; }
            else if (MODE == 12) p[i] = __fadd2_rn(p[i], make_float2(c0, c1));
            else if (MODE == 13) p[i] = __ffma2_rn(p[i], p[(i + 1) % U], p[(i + 2) % U]);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < U; ++i) s += a[i] + p[i].x + p[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int lanes_per_instr) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, 1024>>>(out, cyc, 1.0001f, 0.5f, 0x43000000u);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    double warp_instr = (double)ITERS * U * 32;  // 32 warps per block
    printf("%-34s %.2f warp-instr/clk/SM  (%.0f result lanes/clk/SM)\n", name, warp_instr / avg, warp_instr / avg * lanes_per_instr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA  r, uniform, uniform", 32);
    run<1>("FFMA  r, r, r (3 registers)", 32);
    run<2>("FFMA2 r2, uniform2, uniform2", 64);
    run<13>("FFMA2 r2, r2, r2", 64);
    run<3>("FADD", 32);
    run<12>("FADD2", 64);
    run<4>("FMNMX", 32);
    run<5>("PRMT", 32);
    run<6>("LOP3", 32);
    run<7>("IADD3", 32);
    run<8>("IMAD", 32);
    run<9>("FADD.RM", 32);
    run<10>("FSETP+FSEL", 32);
    run<11>("LDS.128 (+FADD)", 32);
    return 0;
}
