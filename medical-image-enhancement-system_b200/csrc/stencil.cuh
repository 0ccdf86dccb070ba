// stencil.cuh — separable 2-D correlation on a 64x64 output tile staged in shared
// memory, shared by the standalone Gaussian / unsharp kernels (gauss.cu) and the
// fused chain kernels (chain.cu).
//
// Operation order (restated by oracle/mie_oracle.c: sep_row / sep_col): horizontal
// pass first, vertical second; each pass  acc = w[0]*x[0]; acc = fma(w[i], x[i], acc)
// for i = 1..K-1, leftmost / topmost tap first.  Follows
// kornia.filters.gaussian_blur2d(separable=True) -> filter2d_separable
// (SURVEY.md §8(a) A3, Appendix B2).
#pragma once

#include "mie_common.cuh"

namespace mie {

constexpr int kTile = 64;  // output tile edge

// Row pitch of the haloed input tile: >= 64+2R, == 4 (mod 8) words so that a
// quarter-warp of LDS.128 issued by lanes on consecutive rows is conflict-free.
__host__ __device__ constexpr int pin_for(int R) {
    int p = kTile + 2 * R;
    while (p % 8 != 4) ++p;
    return p;
}
constexpr int kPMid = kTile + 4;  // 68 == 4 (mod 32): conflict-free STS.128 by row-lanes

template <int R>
struct TileSmem {
    static constexpr int rows = kTile + 2 * R;
    static constexpr int pin = pin_for(R);
    static constexpr int in_words = rows * pin;
    static constexpr int mid_words = rows * kPMid;
    static constexpr int bytes = (in_words + mid_words) * 4;
};

// Fill the haloed tile: f(gy, gx) -> float for raw (unmapped) image coordinates.
template <int R, typename LoadF>
__device__ __forceinline__ void tile_load(float* __restrict__ s_in, int ty0, int tx0, LoadF f) {
    constexpr int E = kTile + 2 * R;
    for (int i = threadIdx.x; i < E * E; i += 256) {
        const int r = i / E, c = i - r * E;
        s_in[r * TileSmem<R>::pin + c] = f(ty0 - R + r, tx0 - R + c);
    }
}

// Horizontal pass over all 64+2R rows: 8 outputs per work item, window in registers.
template <int R>
__device__ __forceinline__ void tile_row_pass(const float* __restrict__ s_in, float* __restrict__ s_mid,
                                              const Taps& wx) {
    constexpr int ROWS = kTile + 2 * R;
    constexpr int PIN = TileSmem<R>::pin;
    constexpr int NV = (8 + 2 * R + 3) / 4;
    for (int i = threadIdx.x; i < ROWS * 8; i += 256) {
        const int r = i % ROWS, s = i / ROWS;
        const float4* p = reinterpret_cast<const float4*>(s_in + r * PIN + s * 8);
        float win[NV * 4];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            float4 t = p[v];
            win[4 * v] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float acc = __fmul_rn(wx.w[0], win[j]);
#pragma unroll
            for (int t = 1; t <= 2 * R; ++t) acc = __fmaf_rn(wx.w[t], win[j + t], acc);
            o[j] = acc;
        }
        float4* q = reinterpret_cast<float4*>(s_mid + r * kPMid + s * 8);
        q[0] = make_float4(o[0], o[1], o[2], o[3]);
        q[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Vertical pass: each thread owns 4 columns x 4 rows; epi(row, col, float4) gets
// the four horizontally adjacent results of tile row `row`, columns col..col+3.
template <int R, typename EpiF>
__device__ __forceinline__ void tile_col_pass(const float* __restrict__ s_mid, const Taps& wy, EpiF epi) {
    const int q = threadIdx.x & 15, rb = threadIdx.x >> 4;
    float4 win[4 + 2 * R];
#pragma unroll
    for (int k = 0; k < 4 + 2 * R; ++k)
        win[k] = *reinterpret_cast<const float4*>(s_mid + (rb * 4 + k) * kPMid + q * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 a;
        a.x = __fmul_rn(wy.w[0], win[j].x); a.y = __fmul_rn(wy.w[0], win[j].y);
        a.z = __fmul_rn(wy.w[0], win[j].z); a.w = __fmul_rn(wy.w[0], win[j].w);
#pragma unroll
        for (int t = 1; t <= 2 * R; ++t) {
            a.x = __fmaf_rn(wy.w[t], win[j + t].x, a.x); a.y = __fmaf_rn(wy.w[t], win[j + t].y, a.y);
            a.z = __fmaf_rn(wy.w[t], win[j + t].z, a.z); a.w = __fmaf_rn(wy.w[t], win[j + t].w, a.w);
        }
        epi(rb * 4 + j, q * 4, a);
    }
}

}  // namespace mie
