// bilateral.cu — bilateral filter on single-channel planes.
// Replaces kornia.filters.bilateral_blur (reference pyproject.toml:8; SURVEY.md
// §8(a) A7, Appendix B2):  w = space[dy,dx] * exp(-0.5/sigma_color^2 * (v-c)^2),
// out = sum(w*v) / sum(w), full ky x kx window, padding by border_type.
//
// The colour weight uses mie_exp(): range reduction by ln2 (two-constant
// Cody-Waite) and a degree-6 polynomial, all explicit fp32 fma — the same
// sequence as oracle/mie_oracle.c:mie_exp, so the result is reproducible bit for
// bit and stays within 2 ulp of exp().  This op is FMA-pipe bound (81 taps x ~20
// instructions), not HBM bound.
#include "mie_common.cuh"

#define MIE_HAVE_BILATERAL 1

namespace mie {

constexpr int kBilMaxK = 15;

struct SpaceW {
    float w[kBilMaxK * kBilMaxK];
};

__device__ __forceinline__ float mie_exp(float a) {
    a = fminf(fmaxf(a, -87.0f), 88.0f);
    const float n = rintf(__fmul_rn(a, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693145751953125f, a);
    r = __fmaf_rn(n, -1.42860682030941723e-6f, r);
    float p = 1.3888889225e-3f;
    p = __fmaf_rn(p, r, 8.3333337680e-3f);
    p = __fmaf_rn(p, r, 4.1666667908e-2f);
    p = __fmaf_rn(p, r, 1.6666667163e-1f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    const float s = __int_as_float(((int)n + 127) << 23);
    return __fmul_rn(p, s);
}

// 32x32 output tile per block, 4 pixels per thread (rows ly, ly+8, ly+16, ly+24).
template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(256)
bilateral_kernel(const SrcT* __restrict__ src, DstT* __restrict__ dst, int64_t ssn, int64_t ssh, int64_t dsn,
                 int64_t dsh, int h, int w, int tiles_x, int tiles_y, int ky, int kx, float coef, int border, float lo,
                 float rg, SpaceW sw) {
    extern __shared__ __align__(16) float smem[];
    constexpr int T = 32;
    const int ry = ky / 2, rx = kx / 2;
    const int ew = T + 2 * rx, eh = T + 2 * ry, pitch = ew | 1;
    const int64_t tile = blockIdx.x;
    const int tx0 = (int)(tile % tiles_x) * T, ty0 = (int)((tile / tiles_x) % tiles_y) * T;
    const int64_t n = tile / ((int64_t)tiles_x * tiles_y);
    const SrcT* plane = src + n * ssn;
    for (int i = threadIdx.x; i < eh * ew; i += 256) {
        const int r = i / ew, c = i - r * ew;
        const int sy = border_index(ty0 - ry + r, h, border), sx = border_index(tx0 - rx + c, w, border);
        smem[r * pitch + c] = (sy < 0 || sx < 0) ? 0.0f : Px<SrcT>::to01(plane[(int64_t)sy * ssh + sx], lo, rg);
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    const int x = tx0 + lx;
    float ctr[4], num[4], den[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ctr[k] = smem[(ly0 + 8 * k + ry) * pitch + lx + rx];
        num[k] = 0.0f; den[k] = 0.0f;
    }
    for (int dy = 0; dy < ky; ++dy) {
        for (int dx = 0; dx < kx; ++dx) {
            const float ws = sw.w[dy * kx + dx];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v = smem[(ly0 + 8 * k + dy) * pitch + lx + dx];
                const float dv = __fsub_rn(v, ctr[k]);
                const float wgt = __fmul_rn(ws, mie_exp(__fmul_rn(coef, __fmul_rn(dv, dv))));
                num[k] = __fmaf_rn(wgt, v, num[k]);
                den[k] = __fadd_rn(den[k], wgt);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y = ty0 + ly0 + 8 * k;
        if (y < h && x < w)
            dst[n * dsn + (int64_t)y * dsh + x] = Px<DstT>::from01(__fdiv_rn(num[k], den[k]), lo, rg);
    }
}

int bilateral_impl(const void* src, void* dst, int sd, int dd, int64_t n, int h, int w, int64_t ssn, int64_t ssh,
                   int64_t dsn, int64_t dsh, const float* wspace, int ky, int kx, float sigma_color, int border,
                   float lo, float hi, cudaStream_t st) {
    int rc = check_planes(src, dst, n, h, w, ssn, ssh, dsn, dsh);
    if (rc) return rc;
    rc = check_dtypes(sd, dd, lo, hi);
    if (rc) return rc;
    if (!wspace) return MIE_E_NULL;
    if (ky <= 0 || kx <= 0 || !(ky & 1) || !(kx & 1) || ky > kBilMaxK || kx > kBilMaxK) return MIE_E_KERNEL;
    if (border < MIE_BORDER_CONSTANT || border > MIE_BORDER_CIRCULAR) return MIE_E_BORDER;
    if (border == MIE_BORDER_REFLECT && (kx / 2 >= w || ky / 2 >= h)) return MIE_E_BORDER;
    if (border == MIE_BORDER_CIRCULAR && (kx / 2 > w || ky / 2 > h)) return MIE_E_BORDER;
    if (!(sigma_color > 0.0f)) return MIE_E_RANGE;
    if (n == 0) return MIE_OK;
    SpaceW sw;
    for (int i = 0; i < kBilMaxK * kBilMaxK; ++i) sw.w[i] = i < ky * kx ? wspace[i] : 0.0f;
    const float coef = (float)(-0.5 / ((double)sigma_color * (double)sigma_color));
    const int tiles_x = ceil_div(w, 32), tiles_y = ceil_div(h, 32);
    const int64_t blocks = n * tiles_x * tiles_y;
    if (blocks > 2147483647LL) return MIE_E_SHAPE;
    const int ew = 32 + 2 * (kx / 2), eh = 32 + 2 * (ky / 2);
    const size_t smem = (size_t)eh * (ew | 1) * 4;
    MIE_DISPATCH_SRC_DST(sd, dd, (bilateral_kernel<SrcT, DstT><<<(unsigned)blocks, 256, smem, st>>>(
                                     (const SrcT*)src, (DstT*)dst, ssn, ssh, dsn, dsh, h, w, tiles_x, tiles_y, ky, kx,
                                     coef, border, lo, hi - lo, sw)));
    return check_launch();
}

}  // namespace mie

using namespace mie;

extern "C" int mie_bilateral(const void* src, void* dst, int src_dtype, int dst_dtype, int64_t n, int h, int w,
                             int64_t src_stride_n, int64_t src_stride_h, int64_t dst_stride_n, int64_t dst_stride_h,
                             const float* wspace, int ky, int kx, float sigma_color, int border, float lo, float hi,
                             void* stream) {
    return bilateral_impl(src, dst, src_dtype, dst_dtype, n, h, w, src_stride_n, src_stride_h, dst_stride_n,
                          dst_stride_h, wspace, ky, kx, sigma_color, border, lo, hi, (cudaStream_t)stream);
}
