// Microbenchmark (B200): per-SM throughput of the primitives a shared-memory histogram can be
// built from — ATOMS.ADD at k-way same-address multiplicity, MATCH.ANY, REDUX, VOTE, SHFL, POPC.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_prims hist_prims.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE>
__global__ void k(int* out, long long* cyc, int kway) {
    __shared__ int h[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    // lanes are spread over (32 / kway) distinct addresses: multiplicity = kway
    int bin = (lane / kway) * 7 + 3;
    int acc = 0;
    const unsigned lt = (1u << lane) - 1u;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        int b = (bin + it) & 255;
        if (MODE == 0) {
            atomicAdd(&h[warp][b], 1);
        } else if (MODE == 1) {
            unsigned m = __match_any_sync(0xffffffffu, b);
            acc += m;
        } else if (MODE == 2) {
            unsigned m = __match_any_sync(0xffffffffu, b);
            if ((m & lt) == 0) atomicAdd(&h[warp][b], __popc(m));
        } else if (MODE == 3) {
            acc += __reduce_min_sync(0xffffffffu, b);
        } else if (MODE == 4) {
            acc += __ballot_sync(0xffffffffu, b & 1);
        } else if (MODE == 5) {
            acc += __shfl_sync(0xffffffffu, b, 0);
        } else if (MODE == 6) {
            acc += __popc(b * 2654435761u);
        } else if (MODE == 7) {  // byte-counter read-modify-write, lane-private column
            unsigned char* p = reinterpret_cast<unsigned char*>(&h[0][0]) + (warp * 256 + b) * 4 * 0 + ((b * 32 + lane) & 8191);
            *p = *p + 1;
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + h[warp][lane];
}

template <int MODE>
void run(const char* name, int kway, int warps) {
    int* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<MODE><<<148, warps * 32>>>(out, cyc, kway);
    k<MODE><<<148, warps * 32>>>(out, cyc, kway);
    cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    printf("%-28s kway=%2d warps=%d : %.2f cycles per warp-instr per SM (%.1f cyc/iter/warp)\n", name, kway, warps,
           avg / ITERS / warps, avg / ITERS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 8}) {
        for (int kw : {1, 2, 4, 8, 16, 32}) run<0>("ATOMS.ADD", kw, w);
        for (int kw : {1, 4, 32}) run<1>("MATCH.ANY", kw, w);
        for (int kw : {1, 4, 32}) run<2>("MATCH.ANY+leader ATOMS", kw, w);
        run<3>("REDUX.MIN", 1, w);
        run<4>("VOTE.BALLOT", 1, w);
        run<5>("SHFL.IDX", 1, w);
        run<6>("POPC", 1, w);
        for (int kw : {1, 32}) run<7>("byte RMW (LDS.U8+STS.U8)", kw, w);
    }
    return 0;
}
