// clahe.cuh — device building blocks shared by the standalone CLAHE kernels
// (clahe.cu) and the fused chain kernels (chain.cu).
//
// Semantics restated from kornia.enhance.equalize_clahe (kornia 0.8.2,
// reference pyproject.toml:8 / uv.lock:219-230)
// and from cv::CLAHE (SURVEY.md §8(a) A1', Appendix A).
#pragma once

#include "mie_common.cuh"

namespace mie {

struct LutParams {
    int clip;         // per-bin ceiling (0 = clipping disabled)
    int pixels;       // th*tw
    float lut_scale;  // kornia: fp32(255.0/pixels) ; opencv: 255.f/float(pixels)
};

inline LutParams make_lut_params(const ClaheGeom& g, double clip_limit, int semantics) {
    LutParams p;
    p.pixels = g.th * g.tw;
    if (semantics == MIE_CLAHE_KORNIA) {
        // max_val = max(clip * pixels // num_bins, 1)  (python float floor-division)
        p.clip = 0;
        if (clip_limit > 0.0) {
            double q = floor(clip_limit * (double)p.pixels / 256.0);
            if (q < 1.0) q = 1.0;
            p.clip = q > 2147483647.0 ? 2147483647 : (int)q;
        }
        p.lut_scale = (float)(255.0 / (double)p.pixels);
    } else {
        p.clip = 0;
        if (clip_limit > 0.0) {
            double q = clip_limit * (double)p.pixels / 256.0;
            int c = q > 2147483647.0 ? 2147483647 : (int)q;
            p.clip = c < 1 ? 1 : c;
        }
        p.lut_scale = 255.0f / (float)p.pixels;
    }
    return p;
}

// ---------------------------------------------------------------- block primitives (256 threads)
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the 256 threads of the block; s_red: >= 8 ints of shared scratch.
__device__ __forceinline__ int block_sum_256(int v, int* s_red) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_red[i];
    return t;
}

// Inclusive prefix sum over the 256 threads (thread b owns bin b).
__device__ __forceinline__ int block_scan_256(int v, int* s_red) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) s_red[warp] = v;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) off += (i < warp) ? s_red[i] : 0;
    return v + off;
}

// Clip + redistribute + cumulate -> one LUT entry per thread (bin = threadIdx.x).
template <int SEM>
__device__ __forceinline__ uint8_t lut_entry_from_count(int hv, const LutParams& p, int* s_red) {
    const int b = threadIdx.x;
    if (p.clip > 0) {
        if (SEM == MIE_CLAHE_KORNIA) {
            int c = min(hv, p.clip);
            int clipped = p.pixels - block_sum_256(c, s_red);
            int resid = clipped & 255;
            int redist = (clipped - resid) >> 8;
            hv = c + redist + (b < resid ? 1 : 0);
        } else {
            int clipped = block_sum_256(max(hv - p.clip, 0), s_red);
            hv = min(hv, p.clip);
            int rb = clipped >> 8;
            int res = clipped - (rb << 8);
            hv += rb;
            if (res) {
                int step = max(256 / res, 1);
                if (b % step == 0 && b / step < res) hv += 1;
            }
        }
    }
    int cum = block_scan_256(hv, s_red);
    float f = __fmul_rn((float)cum, p.lut_scale);
    if (SEM == MIE_CLAHE_KORNIA) {
        f = floorf(fminf(fmaxf(f, 0.0f), 255.0f));
    } else {
        f = fminf(fmaxf(rintf(f), 0.0f), 255.0f);
    }
    return (uint8_t)(int)f;
}

// ---------------------------------------------------------------- per-pixel index rules
// kornia histogram bin: torch.histc(bins=256, min=0, max=1): values outside [0,1]
// (and NaN) are not counted; x == 1 falls in the last bin.  Returns -1 if ignored.
__device__ __forceinline__ int kornia_bin(float v) {
    if (!(v >= 0.0f && v <= 1.0f)) return -1;
    int b = (int)__fmul_rn(v, 256.0f);
    return b > 255 ? 255 : b;
}
// kornia lookup index: (x*255).long(), clamped to the table (kornia itself would
// raise on out-of-range input).
__device__ __forceinline__ int kornia_idx(float v) {
    float f = __fmul_rn(v, 255.0f);
    f = fminf(fmaxf(f, 0.0f), 255.0f);  // NaN -> 0
    return __float2int_rz(f);
}

// kornia interpolation geometry along one axis: tile size T (even), G tiles.
// Weight `wgt` belongs to tile j0 (top / left); j1 gets the rest.
__device__ __forceinline__ void kornia_axis(int y, int T, int G, int& j0, int& j1, float& wgt) {
    int hh = T >> 1;
    if (y < hh) {
        j0 = j1 = 0; wgt = 0.0f;
    } else if (y >= T * G - hh) {
        j0 = j1 = G - 1; wgt = 0.0f;
    } else {
        int rel = y - hh;
        j0 = rel / T;
        int r = rel - j0 * T;
        j1 = j0 + 1;
        wgt = __fdiv_rn((float)(T - 1 - r), (float)(T - 1));
    }
}

// t = tr + wx*(tl-tr); b = br + wx*(bl-br); out = b + wy*(t-b), each line one fma.
__device__ __forceinline__ float kornia_blend(float tl, float tr, float bl, float br, float wx, float wy) {
    float t = __fmaf_rn(wx, __fsub_rn(tl, tr), tr);
    float b = __fmaf_rn(wx, __fsub_rn(bl, br), br);
    return __fmaf_rn(wy, __fsub_rn(t, b), b);
}

// OpenCV interpolation geometry along one axis.
__device__ __forceinline__ void opencv_axis(int x, float inv_t, int G, int& t1, int& t2, float& a, float& a1) {
    float f = __fsub_rn(__fmul_rn((float)x, inv_t), 0.5f);
    int i1 = (int)floorf(f);
    a = __fsub_rn(f, (float)i1);
    a1 = __fsub_rn(1.0f, a);
    t1 = max(i1, 0);
    t2 = min(i1 + 1, G - 1);
}

__device__ __forceinline__ float opencv_blend(float l11, float l12, float l21, float l22, float xa, float xa1,
                                              float ya, float ya1) {
    float top = __fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa));
    float bot = __fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa));
    return __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
}

}  // namespace mie
