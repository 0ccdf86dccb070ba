// equalize.cu — global histogram equalisation, one histogram per plane.
// Replaces kornia.enhance.equalize -> _scale_channel, the rule torchvision's
// equalize uses as well (reference pyproject.toml:8,16; SURVEY.md §8(a) A2,
// site-packages/torchvision/transforms/_functional_tensor.py:863-881):
//   v = x01 * 255; hist = histc(v, bins=256, min=0, max=255);
//   step = (sum(nonzero) - last_nonzero) // 255;
//   lut = [0, ((cumsum(hist) + step//2) // step)[:-1]] clamped to [0,255];
//   out = lut[trunc(v)] / 255, or v / 255 when step == 0.
// Three launches: per-plane histogram (shared-memory sub-histograms, warp-voted
// adds, one global atomic per non-empty bin per block), LUT (one block per plane),
// apply.
#include "chain_fast.cuh"

namespace mie {

struct EqPlaneState {  // per plane, in the caller's workspace
    unsigned int hist[kBins];
    float lut[kBins];
    int step_nonzero;
    int pad[3];
};

__device__ __forceinline__ int eq_bin(float v) {  // torch.histc(v, 256, 0, 255); -1 = ignored
    if (!(v >= 0.0f && v <= 255.0f)) return -1;
    const int b = (int)__fmul_rn(div255(v), 256.0f);
    return b > 255 ? 255 : b;
}

template <typename SrcT>
__global__ void __launch_bounds__(256)
equalize_hist_kernel(const SrcT* __restrict__ src, int64_t ssn, int64_t ssh, int h, int w, int rows_per_block,
                     float lo, float rg, EqPlaneState* __restrict__ state) {
    __shared__ int s_hist[8 * kBins];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) s_hist[i * kBins + tid] = 0;
    __syncthreads();
    const int64_t n = blockIdx.y;
    const int y0 = blockIdx.x * rows_per_block, y1 = min(y0 + rows_per_block, h);
    const SrcT* plane = src + n * ssn;
    for (int y = y0 + warp; y < y1; y += 8) {
        const SrcT* row = plane + (int64_t)y * ssh;
        for (int x0 = 0; x0 < w; x0 += 32) {
            const int x = x0 + lane;
            int bin = -1;
            if (x < w) bin = eq_bin(__fmul_rn(Px<SrcT>::to01(row[x], lo, rg), 255.0f));
            hist_add(s_hist + warp * kBins, bin);
        }
    }
    __syncthreads();
    int hv = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) hv += s_hist[i * kBins + tid];
    if (hv) atomicAdd(&state[n].hist[tid], (unsigned int)hv);
}

__global__ void __launch_bounds__(256) equalize_lut_kernel(EqPlaneState* __restrict__ state) {
    __shared__ int s_red[8];
    __shared__ int s_last;
    EqPlaneState& st = state[blockIdx.x];
    const int b = threadIdx.x;
    const int hv = (int)st.hist[b];
    const int total = block_sum_256(hv, s_red);
    // index of the last non-empty bin
    int cand = hv ? b : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = max(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    __syncthreads();
    if ((b & 31) == 0) s_red[b >> 5] = cand;
    __syncthreads();
    if (b == 0) {
        int m = -1;
        for (int i = 0; i < 8; ++i) m = max(m, s_red[i]);
        s_last = m;
    }
    __syncthreads();
    const int last = s_last >= 0 ? (int)st.hist[s_last] : 0;
    const int step = (total - last) / 255;
    const int cum = block_scan_256(hv, s_red);  // inclusive
    if (step > 0) {
        // lut[b+1] = (cum[b] + step/2) / step, lut[0] = 0
        const long long q = ((long long)cum + step / 2) / step;
        if (b < 255) st.lut[b + 1] = (float)(q > 255 ? 255 : q);
        if (b == 0) st.lut[0] = 0.0f;
    }
    if (b == 0) st.step_nonzero = step > 0 ? 1 : 0;
}

template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(256)
equalize_apply_kernel(const SrcT* __restrict__ src, DstT* __restrict__ dst, int64_t ssn, int64_t ssh, int64_t dsn,
                      int64_t dsh, int h, int w, float lo, float rg, const EqPlaneState* __restrict__ state) {
    __shared__ float s_lut[kBins];
    const int64_t n = blockIdx.z;
    const EqPlaneState& st = state[n];
    s_lut[threadIdx.x] = st.lut[threadIdx.x];
    const int active = st.step_nonzero;
    __syncthreads();
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    const float v = __fmul_rn(Px<SrcT>::to01(src[n * ssn + (int64_t)y * ssh + x], lo, rg), 255.0f);
    float r = v;
    if (active) {
        const float c = fminf(fmaxf(v, 0.0f), 255.0f);
        r = s_lut[__float2int_rz(c)];
    }
    dst[n * dsn + (int64_t)y * dsh + x] = Px<DstT>::from01(div255(r), lo, rg);
}

}  // namespace mie

using namespace mie;

extern "C" {

size_t mie_equalize_workspace_bytes(int64_t n) { return n > 0 ? (size_t)n * sizeof(EqPlaneState) : 0; }

int mie_equalize(const void* src, void* dst, int src_dtype, int dst_dtype, int64_t n, int h, int w,
                 int64_t src_stride_n, int64_t src_stride_h, int64_t dst_stride_n, int64_t dst_stride_h, float lo,
                 float hi, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_planes(src, dst, n, h, w, src_stride_n, src_stride_h, dst_stride_n, dst_stride_h);
    if (rc) return rc;
    rc = check_dtypes(src_dtype, dst_dtype, lo, hi);
    if (rc) return rc;
    if ((int64_t)h * w >= (1LL << 31)) return MIE_E_SHAPE;
    if (n == 0) return MIE_OK;
    if (n > 65535) return MIE_E_SHAPE;
    if (!workspace) return MIE_E_NULL;
    if (workspace_bytes < mie_equalize_workspace_bytes(n)) return MIE_E_WORKSPACE;
    EqPlaneState* state = (EqPlaneState*)workspace;
    cudaError_t e = cudaMemsetAsync(state, 0, (size_t)n * sizeof(EqPlaneState), st);
    if (e != cudaSuccess) return (int)e;
    const float rg = hi - lo;
    // enough blocks per plane to fill the machine when n is small, few enough to keep global atomics rare
    int blocks_per_plane = (int)((4 * 148 + n - 1) / n);
    if (blocks_per_plane < 1) blocks_per_plane = 1;
    int rows_per_block = ceil_div(h, blocks_per_plane);
    if (rows_per_block < 8) rows_per_block = 8;
    dim3 hgrid((unsigned)ceil_div(h, rows_per_block), (unsigned)n);
    MIE_DISPATCH_SRC(src_dtype, (equalize_hist_kernel<SrcT><<<hgrid, 256, 0, st>>>(
                                    (const SrcT*)src, src_stride_n, src_stride_h, h, w, rows_per_block, lo, rg, state)));
    rc = check_launch();
    if (rc) return rc;
    equalize_lut_kernel<<<(unsigned)n, 256, 0, st>>>(state);
    rc = check_launch();
    if (rc) return rc;
    dim3 agrid((unsigned)ceil_div(w, 64), (unsigned)ceil_div(h, 4), (unsigned)n);
    MIE_DISPATCH_SRC_DST(src_dtype, dst_dtype, (equalize_apply_kernel<SrcT, DstT><<<agrid, 256, 0, st>>>(
                                                   (const SrcT*)src, (DstT*)dst, src_stride_n, src_stride_h,
                                                   dst_stride_n, dst_stride_h, h, w, lo, rg, state)));
    return check_launch();
}

}  // extern "C"
