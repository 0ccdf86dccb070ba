// march.cuh — building blocks of the "marching" kernels (chain_march.cu, gauss_march.cu): full-width
// bands walked two rows per step, row-pair buffers in shared memory, a register ring for the vertical
// pass, packed fma.rn.f32x2 in both passes, and a bulk-copy (TMA) ring for the source rows.
// See chain_march.cu for the schedule these pieces implement.
#pragma once

#include "chain_fast.cuh"

namespace mie {


constexpr int kMR = 4;                     // 9-tap kernels
constexpr int kMRows = kTile + 2 * kMR;    // 72 source rows per band
constexpr int kMPairs = kMRows / 2;        // 36 row pairs
constexpr int kMPro = kMR;                 // prologue: 4 row pairs (8 rows) that produce no output
constexpr int kMRing = 16;                 // register-ring slots (row s lives in slot s % 16)
constexpr int kMUnroll = kMRing / 2;       // main loop: 8 row pairs per iteration
constexpr int kMOffRows = kMRows + 12;     // rows covered by the source-offset table (incl. over-fetch)
constexpr int kHistPitch = 264;            // 257 slots (256 = value 1.0 / ignored pixels), padded
static_assert((kMPairs - kMPro) % kMUnroll == 0, "main loop must tile the band");

// Source row of band row r (|overshoot| <= 4 + 12 < h, so one reflection suffices).
// -1 = outside the image with a constant border.
template <int BORDER>
__device__ __forceinline__ int march_src_row(int r, int h) {
    if (BORDER == MIE_BORDER_REFLECT) {
        const int m = abs(r);
        return min(m, 2 * (h - 1) - m);
    }
    if (BORDER == MIE_BORDER_REPLICATE) return min(max(r, 0), h - 1);
    return (unsigned)r < (unsigned)h ? r : -1;
}

// ---------------------------------------------------------------- row-pair buffers
// One buffer holds two image rows (s, s+1) of W + 8 haloed columns as "quads": quad q = haloed
// columns 4q .. 4q+3 (q = 0: left halo, q = t + 1: thread t, q = T + 1: right halo).  Region A holds
// the first two columns of every quad, region B the last two, each as (row s, row s+1) pairs:
//     A[4q .. 4q+3] = (x0[c], x1[c], x0[c+1], x1[c+1]),   B[4q .. 4q+3] = (x0[c+2], x1[c+2], x0[c+3], x1[c+3])
// Every access below is a 16-byte access with a 16-byte thread stride: conflict-free.
__device__ __forceinline__ int pairbuf_floats(int T) { return 8 * (T + 2); }

template <int BORDER>
__device__ __forceinline__ void pair_store(float* buf, int T, int tid, bool first_warp, bool last_warp,
                                           const float* x0, const float* x1) {
    float* A = buf;
    float* B = buf + 4 * (T + 2);
    *reinterpret_cast<float4*>(A + 4 * (tid + 1)) = make_float4(x0[0], x1[0], x0[1], x1[1]);
    *reinterpret_cast<float4*>(B + 4 * (tid + 1)) = make_float4(x0[2], x1[2], x0[3], x1[3]);
    if (first_warp) {  // columns -4..-1 mirror onto 4, 3, 2, 1
        float n0 = 0.f, n1 = 0.f;
        if (BORDER == MIE_BORDER_REFLECT) {
            n0 = __shfl_down_sync(0xffffffffu, x0[0], 1);
            n1 = __shfl_down_sync(0xffffffffu, x1[0], 1);
        }
        if (tid == 0) {
            float4 a, b;
            if (BORDER == MIE_BORDER_REFLECT) {
                a = make_float4(n0, n1, x0[3], x1[3]);
                b = make_float4(x0[2], x1[2], x0[1], x1[1]);
            } else if (BORDER == MIE_BORDER_REPLICATE) {
                a = b = make_float4(x0[0], x1[0], x0[0], x1[0]);
            } else {
                a = b = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4*>(A) = a;
            *reinterpret_cast<float4*>(B) = b;
        }
    }
    if (last_warp) {  // columns W..W+3 mirror onto W-2, W-3, W-4, W-5
        float n0 = 0.f, n1 = 0.f;
        if (BORDER == MIE_BORDER_REFLECT) {
            n0 = __shfl_up_sync(0xffffffffu, x0[3], 1);
            n1 = __shfl_up_sync(0xffffffffu, x1[3], 1);
        }
        if (tid == T - 1) {
            float4 a, b;
            if (BORDER == MIE_BORDER_REFLECT) {
                a = make_float4(x0[2], x1[2], x0[1], x1[1]);
                b = make_float4(x0[0], x1[0], n0, n1);
            } else if (BORDER == MIE_BORDER_REPLICATE) {
                a = b = make_float4(x0[3], x1[3], x0[3], x1[3]);
            } else {
                a = b = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4*>(A + 4 * (T + 1)) = a;
            *reinterpret_cast<float4*>(B + 4 * (T + 1)) = b;
        }
    }
}

// The same from packed (row s, row s + 1) pairs xp[k] of the thread's four columns: the buffer layout IS that pairing,
// so the two 16-byte stores need no register shuffling.
template <int BORDER>
__device__ __forceinline__ void pair_store_packed(float* buf, int T, int tid, bool first_warp, bool last_warp,
                                                  const f32x2* xp) {
    float* A = buf;
    float* B = buf + 4 * (T + 2);
    *reinterpret_cast<ulonglong2*>(A + 4 * (tid + 1)) = make_ulonglong2(xp[0], xp[1]);
    *reinterpret_cast<ulonglong2*>(B + 4 * (tid + 1)) = make_ulonglong2(xp[2], xp[3]);
    if (first_warp) {  // columns -4..-1 mirror onto 4, 3, 2, 1
        f32x2 nb = 0ull;
        if (BORDER == MIE_BORDER_REFLECT) nb = __shfl_down_sync(0xffffffffu, xp[0], 1);
        if (tid == 0) {
            ulonglong2 a, b;
            if (BORDER == MIE_BORDER_REFLECT) { a = make_ulonglong2(nb, xp[3]); b = make_ulonglong2(xp[2], xp[1]); }
            else if (BORDER == MIE_BORDER_REPLICATE) a = b = make_ulonglong2(xp[0], xp[0]);
            else a = b = make_ulonglong2(0ull, 0ull);
            *reinterpret_cast<ulonglong2*>(A) = a;
            *reinterpret_cast<ulonglong2*>(B) = b;
        }
    }
    if (last_warp) {  // columns W..W+3 mirror onto W-2, W-3, W-4, W-5
        f32x2 nb = 0ull;
        if (BORDER == MIE_BORDER_REFLECT) nb = __shfl_up_sync(0xffffffffu, xp[3], 1);
        if (tid == T - 1) {
            ulonglong2 a, b;
            if (BORDER == MIE_BORDER_REFLECT) { a = make_ulonglong2(xp[2], xp[1]); b = make_ulonglong2(xp[0], nb); }
            else if (BORDER == MIE_BORDER_REPLICATE) a = b = make_ulonglong2(xp[3], xp[3]);
            else a = b = make_ulonglong2(0ull, 0ull);
            *reinterpret_cast<ulonglong2*>(A + 4 * (T + 1)) = a;
            *reinterpret_cast<ulonglong2*>(B + 4 * (T + 1)) = b;
        }
    }
}

// Horizontal 9-tap pass of both rows of a pair for the thread's four columns, written straight into
// the register ring.  Tap order and rounding as everywhere else (acc = w0 x0; acc = fma(w_t, x_t, acc)).
// The window comes out of shared memory already packed — a 16-byte load is two (row s, row s+1)
// operand pairs — so taps 0..7 run two rows per instruction.  The LAST tap is issued as scalar fmas:
// their results are fresh registers, which lets ptxas place (column c, column c+1) of one row in an
// aligned pair for the vertical pass at no cost (re-pairing the halves of packed results instead costs
// two register copies at every one of their nine uses).
// Row-pass arithmetic on a ready window: win[j] = (row s, row s + 1) pair of haloed column j (12 columns for the
// thread's four outputs).
__device__ __forceinline__ void row_pass_window(const f32x2* win, const Taps& wx, f32x2 (&ring)[kMRing][2], const int slot) {
    f32x2 acc[4];
    const f32x2 w0 = f2_pack(wx.w[0], wx.w[0]);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = f2_mul(w0, win[j]);
#pragma unroll
    for (int t = 1; t < 2 * kMR; ++t) {
        const f32x2 wt = f2_pack(wx.w[t], wx.w[t]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = f2_fma(wt, win[j + t], acc[j]);
    }
    float m0[4], m1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float al, ah, xl, xh;
        f2_unpack(acc[j], al, ah);
        f2_unpack(win[j + 2 * kMR], xl, xh);
        m0[j] = __fmaf_rn(wx.w[2 * kMR], xl, al);
        m1[j] = __fmaf_rn(wx.w[2 * kMR], xh, ah);
    }
    ring[slot % kMRing][0] = f2_pack(m0[0], m0[1]);
    ring[slot % kMRing][1] = f2_pack(m0[2], m0[3]);
    ring[(slot + 1) % kMRing][0] = f2_pack(m1[0], m1[1]);
    ring[(slot + 1) % kMRing][1] = f2_pack(m1[2], m1[3]);
}

// `own` (optional): the thread's own four pairs, still in registers from the store — the middle third of the window
// then costs no shared-memory loads.
__device__ __forceinline__ void pair_row_pass(const float* buf, int T, int tid, const Taps& wx,
                                              f32x2 (&ring)[kMRing][2], const int slot, const f32x2* own = nullptr) {
    const ulonglong2* A = reinterpret_cast<const ulonglong2*>(buf) + tid;
    const ulonglong2* B = reinterpret_cast<const ulonglong2*>(buf + 4 * (T + 2)) + tid;
    f32x2 win[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (q == 1 && own) {
            win[4] = own[0]; win[5] = own[1]; win[6] = own[2]; win[7] = own[3];
            continue;
        }
        const ulonglong2 a = A[q], b = B[q];
        win[4 * q] = a.x; win[4 * q + 1] = a.y; win[4 * q + 2] = b.x; win[4 * q + 3] = b.y;
    }
    row_pass_window(win, wx, ring, slot);
}

// Vertical 9-tap pass out of the register ring; OLDEST = ring slot of the topmost tap (a constant
// after unrolling, so the ring stays in registers).
__device__ __forceinline__ void march_col_pass(const f32x2 (&ring)[kMRing][2], const int OLDEST, const Taps& wy,
                                               float* g) {
    const f32x2 w0 = f2_pack(wy.w[0], wy.w[0]);
    f32x2 a = f2_mul(w0, ring[OLDEST % kMRing][0]);
    f32x2 b = f2_mul(w0, ring[OLDEST % kMRing][1]);
#pragma unroll
    for (int t = 1; t <= 2 * kMR; ++t) {
        const f32x2 wt = f2_pack(wy.w[t], wy.w[t]);
        a = f2_fma(wt, ring[(OLDEST + t) % kMRing][0], a);
        b = f2_fma(wt, ring[(OLDEST + t) % kMRing][1], b);
    }
    f2_unpack(a, g[0], g[1]);
    f2_unpack(b, g[2], g[3]);
}

// ---------------------------------------------------------------- bulk-copy (TMA) row ring
// chain_a streams its source rows through shared memory with cp.async.bulk: one elected thread issues
// four row copies per mbarrier, up to sixteen rows (16 KB for uint16, W = 512) ahead of their use.  A
// register prefetch of the same depth would cost 32 registers per thread; with two 8-byte loads per
// thread in flight the kernel could not cover the DRAM latency (Little's law: 16 KB per SM in flight at
// ~1 us is 2.4 TB/s for the whole chip, barely the 1.6 TB/s this kernel needs).
constexpr int kRawRows = 16;   // ring slots (row s lives in slot s % 16)
constexpr int kRawBatch = 4;   // rows per mbarrier
constexpr int kRawBars = kRawRows / kRawBatch;

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// The same with an L2 evict-first policy: source rows that are read exactly once should not displace the intermediates
// (index plane, cell tables) the next kernel of the chain reads back.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_stream(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(mbar), "r"(parity) : "memory");
}

// Byte offsets (int32) of the band's source rows, mirrored at the image border; -1 marks rows outside
// the image with a constant border.  stride_bytes * h < 2^31 is checked on the host.
template <int BORDER>
__device__ __forceinline__ void fill_row_offsets(int* s_off, int ty0, int h, int stride_bytes, int tid, int T) {
    for (int s = tid; s < kMOffRows; s += T) {
        const int sy = march_src_row<BORDER>(ty0 - kMR + s, h);
        s_off[s] = sy < 0 ? -1 : sy * stride_bytes;
    }
}


}  // namespace mie
