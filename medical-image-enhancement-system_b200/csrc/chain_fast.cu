// chain_fast.cu — tuned kernels of the fused Gaussian -> CLAHE -> unsharp chain for
// 64x64-pixel CLAHE tiles (see chain_fast.cuh for the instruction-level tricks and
// chain.cu for the algorithm and the generic kernels these two must equal bit for bit).
#include "chain_fast.cuh"

namespace mie {

// ================================================================ chain_a (fast)
// One block (9 warps) per CLAHE tile:
//   row pass   : 72 rows x 8 segments; a thread loads 16 pixels of one row with
//                64/128-bit loads, converts them, forms 8 horizontal sums and stores
//                them to s_mid (conflict-free STS.128 order);
//   col pass   : warps 0-7, 4 columns x 4 rows per thread out of s_mid;
//   epilogue   : lookup index -> 32-bit stores into the index plane; histogram bin ->
//                warp-voted adds into the warp's private histogram;
//   LUT        : 256 threads fold the 8 private histograms, warp 8 clips / scans.
template <typename SrcT, int R>
__global__ void __launch_bounds__(kFastThreads)
chain_a_fast_kernel(ChainAArgs a, Taps wx, Taps wy) {
    constexpr int ROWS = kTile + 2 * R;
    // s_mid row layout: segment s (8 floats) is split — floats 0..3 at word 4s, floats 4..7 at word
    // 48+4s — so that both the row pass (8 lanes of a row storing 16 B each) and the column pass
    // (16 lanes of a row loading 16 B each) touch 32 distinct banks per quarter-warp.
    constexpr int PM = 80, HI = 48;
    __shared__ __align__(16) float s_mid[ROWS * PM];
    __shared__ __align__(16) int s_hist[8 * kBins];
    __shared__ __align__(16) int s_tot[kBins];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8 * kBins; i += kFastThreads) s_hist[i] = 0;

    const int64_t tile = blockIdx.x;
    const int tx = (int)(tile % a.g.gw), ty = (int)((tile / a.g.gw) % a.g.gh);
    const int64_t n = tile / ((int64_t)a.g.gw * a.g.gh);
    const int tx0 = tx * kTile, ty0 = ty * kTile;
    const int h = a.g.h, w = a.g.w;
    const SrcT* plane = (const SrcT*)a.src + n * a.ssn;

    // ---- horizontal pass, straight from global memory
    for (int i = tid; i < ROWS * 8; i += kFastThreads) {
        const int s = i & 7, r = i >> 3;
        const int sy = border_index(ty0 - R + r, h, a.border);
        float x[16];
        if (sy < 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = 0.0f;
        } else {
            const SrcT* row = plane + (int64_t)sy * a.ssh;
            const int c0 = tx0 + 8 * s;  // first output column of this segment
            const bool ledge = (c0 == 0), redge = (c0 + 8 == w);
            if (!ledge && !redge) {
                Fast<SrcT>::load16(row + c0 - 4, x);
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int sx = border_index(c0 - 4 + k, w, a.border);
                    x[k] = sx < 0 ? 0.0f : Fast<SrcT>::one(row[sx]);
                }
            }
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float acc = __fmul_rn(wx.w[0], x[j + 4 - R]);
#pragma unroll
            for (int t = 1; t <= 2 * R; ++t) acc = __fmaf_rn(wx.w[t], x[j + 4 - R + t], acc);
            o[j] = acc;
        }
        *reinterpret_cast<float4*>(s_mid + r * PM + 4 * s) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(s_mid + r * PM + HI + 4 * s) = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();

    // ---- vertical pass + epilogue (warps 0..7)
    if (warp < 8) {
        const int q = tid & 15, rb = tid >> 4;
        const int qoff = (q & 1) ? HI + 4 * (q >> 1) : 4 * (q >> 1);
        float4 win[4 + 2 * R];
#pragma unroll
        for (int k = 0; k < 4 + 2 * R; ++k)
            win[k] = *reinterpret_cast<const float4*>(s_mid + (rb * 4 + k) * PM + qoff);
        uint8_t* iplane = a.idx + n * (int64_t)h * w;
        int* my_hist = s_hist + warp * kBins;
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
            const uint32_t pack = fast_idx(g[0]) | (fast_idx(g[1]) << 8) | (fast_idx(g[2]) << 16) |
                                  (fast_idx(g[3]) << 24);
            *reinterpret_cast<uint32_t*>(iplane + (int64_t)(ty0 + rb * 4 + j) * w + tx0 + q * 4) = pack;
#pragma unroll
            for (int k = 0; k < 4; ++k) hist_vote_add(my_hist, fast_bin(g[k]), lane);
        }
    }
    __syncthreads();
    if (tid < kBins) {
        int hv = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) hv += s_hist[i * kBins + tid];
        s_tot[tid] = hv;
    }
    __syncthreads();
    if (warp == 8) warp_build_lut(s_tot, a.lp, a.luts + tile * kBins, lane);
}

// ================================================================ chain_b (fast)
// One block per 64x64 output tile (aligned with the CLAHE tile grid, so the haloed
// 72x72 region touches exactly 2x2 interpolation cells):
//   tables : s_cell[a][b][grey] = the four neighbouring LUT entries (tl,tr,bl,br) of
//            cell (a,b) packed in one word; per-row / per-column weight tables;
//   C pass : CLAHE output C for every haloed pixel -> s_in;
//   row / col pass, epilogue: C + (C - blur(C)) -> quantise -> 64-bit stores.
template <typename DstT, int R>
__global__ void __launch_bounds__(kFastThreads)
chain_b_fast_kernel(ChainBArgs a, Taps wx, Taps wy) {
    constexpr int E = kTile + 2 * R;
    constexpr int PIN = TileSmem<R>::pin;
    __shared__ __align__(16) float s_in[E * PIN];
    __shared__ __align__(16) float s_mid[E * kPMid];
    __shared__ uint32_t s_cell[4 * kBins];
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
        // cell 0 = rows / columns before the tile centre, cell 1 = after it
        const int j00 = ty == 0 ? 0 : ty - 1, j01 = ty;                      // a = 0: (top, bottom) LUT rows
        const int j10 = ty, j11 = ty == gh - 1 ? gh - 1 : ty + 1;            // a = 1
        const int i00 = tx == 0 ? 0 : tx - 1, i01 = tx;
        const int i10 = tx, i11 = tx == gw - 1 ? gw - 1 : tx + 1;
        const int jt[2] = {ty == 0 ? 0 : j00, ty == gh - 1 ? gh - 1 : j10};
        const int jb[2] = {ty == 0 ? 0 : j01, j11};
        const int il[2] = {tx == 0 ? 0 : i00, tx == gw - 1 ? gw - 1 : i10};
        const int ir[2] = {tx == 0 ? 0 : i01, i11};
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
    } else if (tid - kBins < 32) {
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

    // ---- CLAHE output of the haloed tile: item = (row r, 8-column chunk u), chunk u covers
    //      tile columns 8u-R' .. 8u-R'+7 with the chunk grid anchored at column -4
    const uint8_t* iplane = a.idx + n * (int64_t)h * w;
    constexpr int CH = (E + 7) / 8;  // chunks per row (9 for R=4 .. R=1: 66 -> 9)
    for (int i = tid; i < E * CH; i += kFastThreads) {
        const int u = i % CH, r = i / CH;
        const int sy = s_sy[r];
        const int c_first = 8 * u;              // first haloed column of the chunk (0-based in the E-wide row)
        float cval[8];
        if (sy < 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) cval[k] = 0.0f;
        } else {
            const float wyv = s_wy[r];
            const int ca = (sy >= ty0 + kTile / 2) ? 2 : 0;
            const uint8_t* irow = iplane + (int64_t)sy * w;
            const int gx0 = tx0 - R + c_first;  // image column of the chunk's first pixel
            uint32_t id[8];
            const bool interior = (gx0 >= 0) && (gx0 + 8 <= w) && ((gx0 & 3) == 0);
            if (interior) {
                const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(irow + gx0));
                const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t*>(irow + gx0 + 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) { id[k] = (w0 >> (8 * k)) & 0xFFu; id[4 + k] = (w1 >> (8 * k)) & 0xFFu; }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int sx = (c_first + k < E) ? s_sx[c_first + k] : -1;
                    id[k] = sx < 0 ? 0x100u : (uint32_t)irow[sx];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int c = c_first + k;
                float v = 0.0f;
                if (id[k] < 0x100u && c < E) {
                    const int sx = interior ? gx0 + k : s_sx[c];
                    const int cb = (sx >= tx0 + kTile / 2) ? 1 : 0;
                    const uint32_t e = s_cell[(ca + cb) * kBins + id[k]];
                    const float A = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7650));  // 2^23 + tl
                    const float B = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7651));  // 2^23 + tr
                    const float C2 = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7652)); // 2^23 + bl
                    const float D = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7653));  // 2^23 + br
                    const float wxv = s_wx[c];
                    const float t = __fmaf_rn(wxv, __fsub_rn(A, B), __fsub_rn(B, 8388608.0f));
                    const float b = __fmaf_rn(wxv, __fsub_rn(C2, D), __fsub_rn(D, 8388608.0f));
                    v = div255(__fmaf_rn(wyv, __fsub_rn(t, b), b));
                }
                cval[k] = v;
            }
        }
        float* dstp = s_in + r * PIN + c_first;
        if (c_first + 8 <= PIN) {
            float4* q = reinterpret_cast<float4*>(dstp);
            const float4 lo4 = make_float4(cval[0], cval[1], cval[2], cval[3]);
            const float4 hi4 = make_float4(cval[4], cval[5], cval[6], cval[7]);
            if (u & 4) { q[1] = hi4; q[0] = lo4; } else { q[0] = lo4; q[1] = hi4; }
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
        DstT* oplane = (DstT*)a.dst + n * a.dsn;
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
            const int r = rb * 4 + j;
            const float* ctr = s_in + (r + R) * PIN + q * 4 + R;
            float y[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = __fadd_rn(ctr[k], __fsub_rn(ctr[k], g[k]));
            Fast<DstT>::store4(oplane + (int64_t)(ty0 + r) * a.dsh + tx0 + q * 4, y);
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
