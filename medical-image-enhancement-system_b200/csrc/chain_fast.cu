// chain_fast.cu — tuned kernels of the fused Gaussian -> CLAHE -> unsharp chain for
// 64x64-pixel CLAHE tiles (see chain_fast.cuh for the instruction-level tricks and
// chain.cu for the algorithm and the generic kernels these two must equal bit for bit).
#include "chain_fast.cuh"

namespace mie {

// 24 converted pixels x[0..23] = image columns c0-4 .. c0+19 of one row (16 outputs + halo 4);
// columns outside the image (left edge: c0 == 0, right edge: c0 + 16 == w) are filled from the
// registers already loaded, so edge tiles cost no extra loads.
template <typename SrcT>
__device__ __forceinline__ void load_row24(const SrcT* row, int c0, int w, int border, float* x) {
    Fast<SrcT>::load8(row + c0, x + 4);
    Fast<SrcT>::load8(row + c0 + 8, x + 12);
    if (c0 != 0) {
        Fast<SrcT>::load4(row + c0 - 4, x);
    } else if (border == MIE_BORDER_REFLECT) {
        x[0] = x[8]; x[1] = x[7]; x[2] = x[6]; x[3] = x[5];
    } else {
        const float e = border == MIE_BORDER_REPLICATE ? x[4] : 0.0f;
        x[0] = e; x[1] = e; x[2] = e; x[3] = e;
    }
    if (c0 + 16 != w) {
        Fast<SrcT>::load4(row + c0 + 16, x + 20);
    } else if (border == MIE_BORDER_REFLECT) {
        x[20] = x[18]; x[21] = x[17]; x[22] = x[16]; x[23] = x[15];
    } else {
        const float e = border == MIE_BORDER_REPLICATE ? x[19] : 0.0f;
        x[20] = e; x[21] = e; x[22] = e; x[23] = e;
    }
}

// s_mid row layout of the tuned chain_a kernel: the 16 column quads of a row are stored even quads
// first (quad q at word 4*(q/2)), odd quads from word 48 (4*(12 + q/2)), rows 84 words apart.  With
// it, the row pass (4 lanes of a row x 2 rows per quarter-warp, STS.128 each) and the column pass
// (8 lanes of a row per quarter-warp, LDS.128 each) both touch 32 distinct banks.
constexpr int kPMa = 84;
__device__ __forceinline__ int quad_off(int q) { return 4 * ((q >> 1) + ((q & 1) ? 12 : 0)); }

// ================================================================ chain_a (fast)
// One block (9 warps) per CLAHE tile:
//   row pass   : 72 rows x 4 segments = one item per thread; a thread loads 24 pixels of one row
//                with 64/128-bit loads, converts them, forms 16 horizontal sums -> s_mid;
//   col pass   : warps 0-7, 4 columns x 4 rows per thread out of s_mid;
//   epilogue   : lookup index -> 32-bit stores into the index plane; histogram bin ->
//                ATOMS.POPC.INC into the block histogram;
//   LUT        : warp 8 clips / redistributes / scans (8 bins per lane).
template <typename SrcT, int R>
__global__ void __launch_bounds__(kFastThreads)
chain_a_fast_kernel(ChainAArgs a, Taps wx, Taps wy) {
    constexpr int ROWS = kTile + 2 * R;
    constexpr bool NN = !(sizeof(SrcT) == 4);  // integer pixels: blurred values are >= 0 and finite
    __shared__ __align__(16) float s_mid[ROWS * kPMa];
    __shared__ __align__(16) int s_hist[kBins + 8];  // [256] = dummy slot for ignored pixels

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kBins + 8) s_hist[tid] = 0;

    const int64_t tile = blockIdx.x;
    const int tx = (int)(tile % a.g.gw), ty = (int)((tile / a.g.gw) % a.g.gh);
    const int64_t n = tile / ((int64_t)a.g.gw * a.g.gh);
    const int tx0 = tx * kTile, ty0 = ty * kTile;
    const int h = a.g.h, w = a.g.w;
    const SrcT* plane = (const SrcT*)a.src + n * a.ssn;

    // ---- horizontal pass, straight from global memory
    for (int i = tid; i < ROWS * 4; i += kFastThreads) {
        const int s = i & 3, r = i >> 2;
        const int sy = border_index(ty0 - R + r, h, a.border);
        float x[24];
        if (sy < 0) {
#pragma unroll
            for (int k = 0; k < 24; ++k) x[k] = 0.0f;
        } else {
            load_row24<SrcT>(plane + (int64_t)sy * a.ssh, tx0 + 16 * s, w, a.border, x);
        }
        float* mrow = s_mid + r * kPMa + 8 * s;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float acc = __fmul_rn(wx.w[0], x[4 * k + j + 4 - R]);
#pragma unroll
                for (int t = 1; t <= 2 * R; ++t) acc = __fmaf_rn(wx.w[t], x[4 * k + j + 4 - R + t], acc);
                o[j] = acc;
            }
            // quad 4s+k: even quads at word 4*(2s + k/2), odd quads at 48 + 4*(2s + k/2)
            *reinterpret_cast<float4*>(mrow + ((k & 1) ? 48 : 0) + 4 * (k >> 1)) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();

    // ---- vertical pass + epilogue (warps 0..7)
    if (warp < 8) {
        const int q = tid & 15, rb = tid >> 4;
        const int qoff = quad_off(q);
        float4 win[4 + 2 * R];
#pragma unroll
        for (int k = 0; k < 4 + 2 * R; ++k)
            win[k] = *reinterpret_cast<const float4*>(s_mid + (rb * 4 + k) * kPMa + qoff);
        uint8_t* ip = a.idx + n * (int64_t)h * w + (int64_t)(ty0 + rb * 4) * w + tx0 + q * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g[4];
            g[0] = __fmul_rn(wy.w[0], win[j].x); g[1] = __fmul_rn(wy.w[0], win[j].y);
            g[2] = __fmul_rn(wy.w[0], win[j].z); g[3] = __fmul_rn(wy.w[0], win[j].w);
#pragma unroll
            for (int t = 1; t <= 2 * R; ++t) {
                g[0] = __fmaf_rn(wy.w[t], win[j + t].x, g[0]); g[1] = __fmaf_rn(wy.w[t], win[j + t].y, g[1]);
                g[2] = __fmaf_rn(wy.w[t], win[j + t].z, g[2]); g[3] = __fmaf_rn(wy.w[t], win[j + t].w, g[3]);
            }
            *reinterpret_cast<uint32_t*>(ip + (int64_t)j * w) =
                pack_low_bytes(fast_idx_bits<NN>(g[0]), fast_idx_bits<NN>(g[1]), fast_idx_bits<NN>(g[2]),
                               fast_idx_bits<NN>(g[3]));
#pragma unroll
            for (int k = 0; k < 4; ++k) hist_add_nobranch(s_hist, fast_bin<NN>(g[k]));
        }
    }
    __syncthreads();
    if (warp == 8) warp_build_lut(s_hist, a.lp, a.luts + tile * kBins, lane);
}

// ================================================================ chain_b (fast)
// One block per 64x64 output tile (aligned with the CLAHE tile grid, so the haloed
// 72x72 region touches exactly 2x2 interpolation cells):
//   tables : s_cell[a][b][grey] = the four neighbouring LUT entries (tl,tr,bl,br) of
//            cell (a,b) packed in one word; per-row / per-column weight tables;
//   C pass : CLAHE output C for every haloed pixel -> s_in;
//   row / col pass, epilogue: C + (C - blur(C)) -> quantise -> 64-bit stores.

// CLAHE output of one pixel: e = packed (tl,tr,bl,br); bytes become floats by OR-ing them into the
// mantissa of 2^23 (differences of two such floats are exact).
__device__ __forceinline__ float clahe_px(uint32_t e, float wxv, float wyv) {
    const float A = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7650));  // 2^23 + tl
    const float B = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7651));  // 2^23 + tr
    const float C = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7652));  // 2^23 + bl
    const float D = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7653));  // 2^23 + br
    const float t = __fmaf_rn(wxv, __fsub_rn(A, B), __fsub_rn(B, 8388608.0f));
    const float b = __fmaf_rn(wxv, __fsub_rn(C, D), __fsub_rn(D, 8388608.0f));
    return div255(__fmaf_rn(wyv, __fsub_rn(t, b), b));
}

template <typename DstT, int R>
__global__ void __launch_bounds__(kFastThreads)
chain_b_fast_kernel(ChainBArgs a, Taps wx, Taps wy) {
    constexpr int E = kTile + 2 * R;
    constexpr int PIN = TileSmem<R>::pin;
    __shared__ __align__(16) float s_in[E * PIN];
    __shared__ __align__(16) float s_mid[E * kPMid];
    __shared__ __align__(16) uint32_t s_cell[4 * kBins];
    __shared__ __align__(16) float s_wx[E + 8];
    __shared__ float s_wy[E];
    __shared__ int s_sy[E], s_sx[E + 8];

    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int tx = (int)(tile % a.tiles_x), ty = (int)((tile / a.tiles_x) % a.tiles_y);
    const int64_t n = tile / ((int64_t)a.tiles_x * a.tiles_y);
    const int tx0 = tx * kTile, ty0 = ty * kTile;
    const int h = a.g.h, w = a.g.w, gh = a.g.gh, gw = a.g.gw;

    // ---- tables
    if (tid < kBins) {
        const uint8_t* nl = a.luts + n * (int64_t)gh * gw * kBins + tid;
        // cell 0 = source rows (columns) before the tile centre, cell 1 = after it
        const int jt[2] = {ty == 0 ? 0 : ty - 1, ty};
        const int jb[2] = {ty, ty == gh - 1 ? gh - 1 : ty + 1};
        const int il[2] = {tx == 0 ? 0 : tx - 1, tx};
        const int ir[2] = {tx, tx == gw - 1 ? gw - 1 : tx + 1};
#pragma unroll
        for (int ca = 0; ca < 2; ++ca)
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                const uint32_t tl = __ldg(nl + ((int64_t)jt[ca] * gw + il[cb]) * kBins);
                const uint32_t tr = __ldg(nl + ((int64_t)jt[ca] * gw + ir[cb]) * kBins);
                const uint32_t bl = __ldg(nl + ((int64_t)jb[ca] * gw + il[cb]) * kBins);
                const uint32_t br = __ldg(nl + ((int64_t)jb[ca] * gw + ir[cb]) * kBins);
                s_cell[(ca * 2 + cb) * kBins + tid] = tl | (tr << 8) | (bl << 16) | (br << 24);
            }
    } else {
        // warp 8: per-row and per-column source coordinate + interpolation weight
        for (int k = tid - kBins; k < 2 * E; k += 32) {
            const bool is_row = k < E;
            const int kk = is_row ? k : k - E;
            const int len = is_row ? h : w;
            const int src = border_index((is_row ? ty0 : tx0) - R + kk, len, a.border);
            int j0, j1;
            float wgt = 0.0f;
            if (src >= 0) kornia_axis(src, kTile, is_row ? gh : gw, j0, j1, wgt);
            if (is_row) { s_sy[kk] = src; s_wy[kk] = wgt; } else { s_sx[kk] = src; s_wx[kk] = wgt; }
        }
    }
    __syncthreads();

    // ---- CLAHE output of the haloed tile: item = (row r, 8-column chunk u)
    const uint8_t* iplane = a.idx + n * (int64_t)h * w;
    constexpr int CH = (E + 7) / 8;
    for (int i = tid; i < E * CH; i += kFastThreads) {
        const int u = i % CH, r = i / CH;
        const int sy = s_sy[r];
        const int c_first = 8 * u;  // first haloed column of the chunk
        float cval[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cval[k] = 0.0f;
        if (sy >= 0) {
            const float wyv = s_wy[r];
            const int ca = (sy >= ty0 + kTile / 2) ? 2 * kBins : 0;
            const uint8_t* irow = iplane + (int64_t)sy * w;
            if (R == 4) {
                // chunk = image columns gx0 .. gx0+7, gx0 = tx0 - 4 + 8u (4-byte aligned)
                const int gx0 = tx0 - 4 + c_first;
                uint32_t w0, w1;
                bool z0 = false, z1 = false;
                if (gx0 < 0) {  // columns -4..-1 mirror onto 4,3,2,1
                    w1 = __ldg(reinterpret_cast<const uint32_t*>(irow));
                    if (a.border == MIE_BORDER_REFLECT)
                        w0 = __byte_perm(w1, __ldg(reinterpret_cast<const uint32_t*>(irow + 4)), 0x1234);
                    else if (a.border == MIE_BORDER_REPLICATE) w0 = __byte_perm(w1, 0u, 0x0000);
                    else { w0 = 0u; z0 = true; }
                } else if (gx0 + 8 > w) {  // columns w..w+3 mirror onto w-2..w-5
                    w0 = __ldg(reinterpret_cast<const uint32_t*>(irow + gx0));
                    if (a.border == MIE_BORDER_REFLECT)
                        w1 = __byte_perm(w0, __ldg(reinterpret_cast<const uint32_t*>(irow + gx0 - 4)), 0x7012);
                    else if (a.border == MIE_BORDER_REPLICATE) w1 = __byte_perm(w0, 0u, 0x3333);
                    else { w1 = 0u; z1 = true; }
                } else {
                    w0 = __ldg(reinterpret_cast<const uint32_t*>(irow + gx0));
                    w1 = __ldg(reinterpret_cast<const uint32_t*>(irow + gx0 + 4));
                }
                // the cell column flips between haloed columns 35 and 36 (source column tx0+32)
                const uint32_t* cellL = s_cell + ca + (c_first >= 36 ? kBins : 0);
                const uint32_t* cellH = s_cell + ca + (c_first + 4 >= 36 ? kBins : 0);
                const float4 wa = *reinterpret_cast<const float4*>(s_wx + c_first);
                const float4 wb = *reinterpret_cast<const float4*>(s_wx + c_first + 4);
                cval[0] = clahe_px(cellL[w0 & 0xFFu], wa.x, wyv);
                cval[1] = clahe_px(cellL[(w0 >> 8) & 0xFFu], wa.y, wyv);
                cval[2] = clahe_px(cellL[(w0 >> 16) & 0xFFu], wa.z, wyv);
                cval[3] = clahe_px(cellL[w0 >> 24], wa.w, wyv);
                cval[4] = clahe_px(cellH[w1 & 0xFFu], wb.x, wyv);
                cval[5] = clahe_px(cellH[(w1 >> 8) & 0xFFu], wb.y, wyv);
                cval[6] = clahe_px(cellH[(w1 >> 16) & 0xFFu], wb.z, wyv);
                cval[7] = clahe_px(cellH[w1 >> 24], wb.w, wyv);
                if (z0) { cval[0] = 0.f; cval[1] = 0.f; cval[2] = 0.f; cval[3] = 0.f; }
                if (z1) { cval[4] = 0.f; cval[5] = 0.f; cval[6] = 0.f; cval[7] = 0.f; }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = c_first + k;
                    const int sx = c < E ? s_sx[c] : -1;
                    if (sx >= 0) {
                        const int cb = (sx >= tx0 + kTile / 2) ? kBins : 0;
                        cval[k] = clahe_px(s_cell[ca + cb + irow[sx]], s_wx[c], wyv);
                    }
                }
            }
        }
        float* dstp = s_in + r * PIN + c_first;
        if (c_first + 8 <= PIN) {
            *reinterpret_cast<float4*>(dstp) = make_float4(cval[0], cval[1], cval[2], cval[3]);
            *reinterpret_cast<float4*>(dstp + 4) = make_float4(cval[4], cval[5], cval[6], cval[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c_first + k < PIN) dstp[k] = cval[k];
        }
    }
    __syncthreads();

    // ---- horizontal pass out of s_in (lanes on consecutive rows: conflict-free LDS.128 / STS.128)
    {
        constexpr int NV = (8 + 2 * R + 3) / 4;
        for (int i = tid; i < E * 8; i += kFastThreads) {
            const int r = i % E, s = i / E;
            const float4* p = reinterpret_cast<const float4*>(s_in + r * PIN + s * 8);
            float win[NV * 4];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 t = p[v];
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
    __syncthreads();

    // ---- vertical pass + unsharp + quantise (warps 0..7)
    if (tid < 256) {
        const int q = tid & 15, rb = tid >> 4;
        float4 win[4 + 2 * R];
#pragma unroll
        for (int k = 0; k < 4 + 2 * R; ++k)
            win[k] = *reinterpret_cast<const float4*>(s_mid + (rb * 4 + k) * kPMid + q * 4);
        DstT* op = (DstT*)a.dst + n * a.dsn + (int64_t)(ty0 + rb * 4) * a.dsh + tx0 + q * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g[4];
            g[0] = __fmul_rn(wy.w[0], win[j].x); g[1] = __fmul_rn(wy.w[0], win[j].y);
            g[2] = __fmul_rn(wy.w[0], win[j].z); g[3] = __fmul_rn(wy.w[0], win[j].w);
#pragma unroll
            for (int t = 1; t <= 2 * R; ++t) {
                g[0] = __fmaf_rn(wy.w[t], win[j + t].x, g[0]); g[1] = __fmaf_rn(wy.w[t], win[j + t].y, g[1]);
                g[2] = __fmaf_rn(wy.w[t], win[j + t].z, g[2]); g[3] = __fmaf_rn(wy.w[t], win[j + t].w, g[3]);
            }
            const float* cp = s_in + (rb * 4 + j + R) * PIN + q * 4 + R;
            float c[4];
            if (R == 4) {
                const float4 cv = *reinterpret_cast<const float4*>(cp);
                c[0] = cv.x; c[1] = cv.y; c[2] = cv.z; c[3] = cv.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) c[k] = cp[k];
            }
            float y[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = __fadd_rn(c[k], __fsub_rn(c[k], g[k]));
            Fast<DstT>::store4(op + (int64_t)j * a.dsh, y);
        }
    }
}

// ================================================================ host side
static bool default_range(int dtype, float lo, float hi) {
    switch (dtype) {
        case MIE_U8: return lo == 0.0f && hi == 255.0f;
        case MIE_U16: return lo == 0.0f && hi == 65535.0f;
        case MIE_I16: return lo == -32768.0f && hi == 32767.0f;
        default: return true;
    }
}

bool fast_chain_ok(const ClaheGeom& g, int sd, int dd, const void* src, int64_t ssn, int64_t ssh, const void* dst,
                   int64_t dsn, int64_t dsh, int kg, int ku, int border, float lo, float hi) {
    static const int esz[4] = {1, 2, 2, 4};
    if (g.th != kTile || g.tw != kTile || g.hp != g.h || g.wp != g.w) return false;
    if (kg < 3 || kg > 9 || ku < 3 || ku > 9) return false;
    if (border == MIE_BORDER_CIRCULAR) return false;
    if (!default_range(sd, lo, hi) || !default_range(dd, lo, hi)) return false;
    // 16-byte aligned rows for the vector loads / stores
    if (((uintptr_t)src % 16) || ((ssn * esz[sd]) % 16) || ((ssh * esz[sd]) % 16)) return false;
    if (((uintptr_t)dst % 16) || ((dsn * esz[dd]) % 16) || ((dsh * esz[dd]) % 16)) return false;
    return true;
}

template <typename SrcT>
static int launch_a_t(const ChainAArgs& a, const Taps& wx, const Taps& wy, int R, unsigned blocks, cudaStream_t st) {
    switch (R) {
        case 1: chain_a_fast_kernel<SrcT, 1><<<blocks, kFastThreads, 0, st>>>(a, wx, wy); break;
        case 2: chain_a_fast_kernel<SrcT, 2><<<blocks, kFastThreads, 0, st>>>(a, wx, wy); break;
        case 3: chain_a_fast_kernel<SrcT, 3><<<blocks, kFastThreads, 0, st>>>(a, wx, wy); break;
        default: chain_a_fast_kernel<SrcT, 4><<<blocks, kFastThreads, 0, st>>>(a, wx, wy); break;
    }
    return check_launch();
}

int launch_chain_a_fast(const ChainAArgs& a, int sd, const Taps& wx, const Taps& wy, int R, int64_t n,
                        cudaStream_t st) {
    const unsigned blocks = (unsigned)(n * a.g.gh * a.g.gw);
    MIE_DISPATCH_SRC(sd, return launch_a_t<SrcT>(a, wx, wy, R, blocks, st));
    return MIE_OK;
}

template <typename DstT>
static int launch_b_t(const ChainBArgs& b, const Taps& wx, const Taps& wy, int R, unsigned blocks, cudaStream_t st) {
    switch (R) {
        case 1: chain_b_fast_kernel<DstT, 1><<<blocks, kFastThreads, 0, st>>>(b, wx, wy); break;
        case 2: chain_b_fast_kernel<DstT, 2><<<blocks, kFastThreads, 0, st>>>(b, wx, wy); break;
        case 3: chain_b_fast_kernel<DstT, 3><<<blocks, kFastThreads, 0, st>>>(b, wx, wy); break;
        default: chain_b_fast_kernel<DstT, 4><<<blocks, kFastThreads, 0, st>>>(b, wx, wy); break;
    }
    return check_launch();
}

int launch_chain_b_fast(const ChainBArgs& b, int dd, const Taps& wx, const Taps& wy, int R, int64_t n,
                        cudaStream_t st) {
    const unsigned blocks = (unsigned)(n * b.tiles_x * b.tiles_y);
    MIE_DISPATCH_SRC(dd, return launch_b_t<SrcT>(b, wx, wy, R, blocks, st));
    return MIE_OK;
}

}  // namespace mie
