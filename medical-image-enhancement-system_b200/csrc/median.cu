// median.cu — 2-D K x K and 3-D 3x3x3 median filters (pure selection, bit-exact).
//
// mie_median2d replaces kornia.filters.median_blur (zero padding, lower median ==
// true median for odd window sizes) and skimage.filters.median on 2-D input
// ('nearest' border); SURVEY.md §8(a) A5.
// mie_median3d replaces skimage.filters.median on a volume ->
// scipy.ndimage.median_filter(footprint=ones((3,3,3)), mode='nearest'), rank
// 27//2 = 13 (site-packages/scipy/ndimage/_filters.py:1961-1962); §8(a) A6.
//
// Selection is done in registers with "forgetful selection": keep a working set
// of N/2+2 candidates, repeatedly drop its minimum and maximum (neither can be
// the median) and insert the next sample.  All indices are compile-time, so the
// whole network is min/max instructions on registers.
#include <cstdlib>

#include "mie_common.cuh"

#define MIE_HAVE_MEDIAN 1

namespace mie {

template <typename C>
__device__ __forceinline__ void cswap(C& a, C& b) {
    C lo = a < b ? a : b;
    C hi = a < b ? b : a;
    a = lo; b = hi;
}
template <>
__device__ __forceinline__ void cswap<int>(int& a, int& b) {
    int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
template <>
__device__ __forceinline__ void cswap<float>(float& a, float& b) {
    float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
}

// Moves the minimum of a[0..CNT) to a[0] and the maximum to a[CNT-1].
template <int CNT, typename C>
__device__ __forceinline__ void minmax_to_ends(C* a) {
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) cswap(a[i], a[CNT - 1 - i]);
#pragma unroll
    for (int i = 1; i < (CNT + 1) / 2; ++i) cswap(a[0], a[i]);
#pragma unroll
    for (int i = CNT / 2; i < CNT - 1; ++i) cswap(a[i], a[CNT - 1]);
}

template <int CNT, int NEXT, int N, typename C>
struct Forget {
    static __device__ __forceinline__ C run(C* a, const C* v) {
        minmax_to_ends<CNT>(a);
        if constexpr (NEXT < N) {
            // drop max (slot CNT-1) and replace min (slot 0) with the next sample
            a[0] = v[NEXT];
            return Forget<CNT - 1, NEXT + 1, N, C>::run(a, v);
        } else if constexpr (CNT == 3) {
            return a[1];
        } else {
            // no samples left: drop both ends and continue on the interior
            return Forget<CNT - 2, NEXT, N, C>::run(a + 1, v);
        }
    }
};
template <int NEXT, int N, typename C>
struct Forget<1, NEXT, N, C> {
    static __device__ __forceinline__ C run(C* a, const C*) { return a[0]; }
};
template <int NEXT, int N, typename C>
struct Forget<2, NEXT, N, C> {
    static __device__ __forceinline__ C run(C* a, const C*) { return a[0] < a[1] ? a[0] : a[1]; }
};

// Median (rank N/2, 0-based, of the sorted samples) of v[0..N), N odd.
template <int N, typename C>
__device__ __forceinline__ C median_of(const C* v) {
    constexpr int M = N / 2 + 2 < N ? N / 2 + 2 : N;
    C a[M];
#pragma unroll
    for (int i = 0; i < M; ++i) a[i] = v[i];
    return Forget<M, M, N, C>::run(a, v);
}

template <typename T> struct Cmp { using type = int; };
template <> struct Cmp<float> { using type = float; };

// ---------------------------------------------------------------- 2-D
template <typename T, int KY, int KX>
__global__ void __launch_bounds__(256)
median2d_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t ssn, int64_t ssh, int64_t dsn, int64_t dsh,
                int h, int w, int tiles_x, int tiles_y, int border) {
    using C = typename Cmp<T>::type;
    constexpr int TW = 32, TH = 32, RY = KY / 2, RX = KX / 2, EW = TW + 2 * RX, EH = TH + 2 * RY;
    __shared__ C s[EH][EW + 1];
    const int64_t tile = blockIdx.x;
    const int tx0 = (int)(tile % tiles_x) * TW, ty0 = (int)((tile / tiles_x) % tiles_y) * TH;
    const int64_t n = tile / ((int64_t)tiles_x * tiles_y);
    const T* plane = src + n * ssn;
    for (int i = threadIdx.x; i < EH * EW; i += 256) {
        const int r = i / EW, c = i - r * EW;
        const int sy = border_index(ty0 - RY + r, h, border), sx = border_index(tx0 - RX + c, w, border);
        s[r][c] = (sy < 0 || sx < 0) ? (C)0 : (C)plane[(int64_t)sy * ssh + sx];
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
#pragma unroll 1
    for (int ly = ly0; ly < TH; ly += 8) {
        const int y = ty0 + ly, x = tx0 + lx;
        if (y >= h || x >= w) continue;
        C v[KY * KX];
#pragma unroll
        for (int dy = 0; dy < KY; ++dy)
#pragma unroll
            for (int dx = 0; dx < KX; ++dx) v[dy * KX + dx] = s[ly + dy][lx + dx];
        dst[n * dsn + (int64_t)y * dsh + x] = (T)median_of<KY * KX, C>(v);
    }
}

template <typename T, int KY, int KX>
static int launch_median2d(const void* src, void* dst, int64_t n, int h, int w, int64_t ssn, int64_t ssh, int64_t dsn,
                           int64_t dsh, int border, cudaStream_t st) {
    const int tiles_x = ceil_div(w, 32), tiles_y = ceil_div(h, 32);
    const int64_t blocks = n * tiles_x * tiles_y;
    if (blocks > 2147483647LL) return MIE_E_SHAPE;
    median2d_kernel<T, KY, KX><<<(unsigned)blocks, 256, 0, st>>>((const T*)src, (T*)dst, ssn, ssh, dsn, dsh, h, w,
                                                                tiles_x, tiles_y, border);
    return check_launch();
}

template <typename T>
static int dispatch_median2d(int ky, int kx, const void* src, void* dst, int64_t n, int h, int w, int64_t ssn,
                             int64_t ssh, int64_t dsn, int64_t dsh, int border, cudaStream_t st) {
#define MIE_M2D(KY_, KX_) \
    if (ky == KY_ && kx == KX_) return launch_median2d<T, KY_, KX_>(src, dst, n, h, w, ssn, ssh, dsn, dsh, border, st);
    MIE_M2D(3, 3) MIE_M2D(5, 5) MIE_M2D(7, 7) MIE_M2D(1, 3) MIE_M2D(3, 1) MIE_M2D(3, 5) MIE_M2D(5, 3) MIE_M2D(1, 1)
#undef MIE_M2D
    return MIE_E_KERNEL;
}

// ---------------------------------------------------------------- 3-D, 3x3x3
// One block owns a 32x8 (x,y) tile and walks a chunk of z with a rolling ring of
// three haloed planes in shared memory, so each plane is read once per block.
template <typename T>
__global__ void __launch_bounds__(256)
median3d_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t ssd, int64_t ssh, int64_t dsd, int64_t dsh,
                int d, int h, int w, int tiles_x, int tiles_y, int zchunk, const T* __restrict__ halo_lo,
                const T* __restrict__ halo_hi, int border) {
    using C = typename Cmp<T>::type;
    constexpr int TW = 32, TH = 8, EW = TW + 2, EH = TH + 2;
    __shared__ C ring[3][EH][EW + 1];
    const int tx0 = (int)(blockIdx.x % tiles_x) * TW, ty0 = (int)(blockIdx.x / tiles_x) * TH;
    const int z0 = blockIdx.y * zchunk, z1 = min(z0 + zchunk, d);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;

    auto load_plane = [&](int z) {
        // z in [-1, d]: -1 / d come from the neighbouring slab's halo plane or the border rule
        const T* p = nullptr;
        int64_t rs = ssh;
        if (z < 0) {
            if (halo_lo) { p = halo_lo; rs = w; }
            else if (border == MIE_BORDER_REPLICATE) p = src;
        } else if (z >= d) {
            if (halo_hi) { p = halo_hi; rs = w; }
            else if (border == MIE_BORDER_REPLICATE) p = src + (int64_t)(d - 1) * ssd;
        } else {
            p = src + (int64_t)z * ssd;
        }
        C(*slot)[EW + 1] = ring[(z + 3) % 3];
        for (int i = threadIdx.x; i < EH * EW; i += 256) {
            const int r = i / EW, c = i - r * EW;
            const int sy = border_index(ty0 - 1 + r, h, border), sx = border_index(tx0 - 1 + c, w, border);
            slot[r][c] = (!p || sy < 0 || sx < 0) ? (C)0 : (C)p[(int64_t)sy * rs + sx];
        }
    };

    load_plane(z0 - 1);
    load_plane(z0);
    for (int z = z0; z < z1; ++z) {
        load_plane(z + 1);
        __syncthreads();
        const int y = ty0 + ly, x = tx0 + lx;
        if (y < h && x < w) {
            C v[27];
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                C(*slot)[EW + 1] = ring[(z + dz + 2) % 3];  // planes z-1, z, z+1
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) v[dz * 9 + dy * 3 + dx] = slot[ly + dy][lx + dx];
            }
            dst[(int64_t)z * dsd + (int64_t)y * dsh + x] = (T)median_of<27, C>(v);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 3-D, 3x3x3, 16-bit voxels, two per lane
// Packed variant for uint16 / int16 volumes: a lane owns TWO x-adjacent voxels held in one 32-bit word and
// runs the same forgetful selection on min.{u,s}16x2 / max.{u,s}16x2 (VIMNMX.U16 / .S16), which issue at
// the rate of the scalar min/max (profiles/microbench/minmax_rates_b200_r1.log: 2.0 warp instructions per
// clock per SM) — so the ~300 min/max of a 27-sample median serve two voxels.  The three z-planes of the
// window live in registers as 3x3 packed words each; a z-step loads one new plane (nine 32-bit
// shared-memory loads + six PRMT to form the x-1 / x+1 neighbours), the other two are carried over.
struct PackedU16 {
    uint32_t v;
};
struct PackedS16 {
    uint32_t v;
};
__device__ __forceinline__ void cswap(PackedU16& a, PackedU16& b) {
    uint32_t lo, hi;
    asm("min.u16x2 %0, %2, %3;\n\tmax.u16x2 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a.v), "r"(b.v));
    a.v = lo; b.v = hi;
}
__device__ __forceinline__ void cswap(PackedS16& a, PackedS16& b) {
    uint32_t lo, hi;
    asm("min.s16x2 %0, %2, %3;\n\tmax.s16x2 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a.v), "r"(b.v));
    a.v = lo; b.v = hi;
}
__device__ __forceinline__ bool operator<(const PackedU16& a, const PackedU16& b) { return a.v < b.v; }  // unused
__device__ __forceinline__ bool operator<(const PackedS16& a, const PackedS16& b) { return a.v < b.v; }  // unused
template <typename P>
__device__ __forceinline__ P packed_min(P a, P b) { cswap(a, b); return a; }

template <int NEXT, int N>
struct Forget<2, NEXT, N, PackedU16> {
    static __device__ __forceinline__ PackedU16 run(PackedU16* a, const PackedU16*) { return packed_min(a[0], a[1]); }
};
template <int NEXT, int N>
struct Forget<2, NEXT, N, PackedS16> {
    static __device__ __forceinline__ PackedS16 run(PackedS16* a, const PackedS16*) { return packed_min(a[0], a[1]); }
};

template <typename T> struct PackedOf;
template <> struct PackedOf<uint16_t> { using type = PackedU16; };
template <> struct PackedOf<int16_t> { using type = PackedS16; };

template <typename T>
__global__ void __launch_bounds__(256)
median3d_packed_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t ssd, int64_t ssh, int64_t dsd,
                       int64_t dsh, int d, int h, int w, int tiles_x, int tiles_y, int zchunk,
                       const T* __restrict__ halo_lo, const T* __restrict__ halo_hi, int border) {
    using P = typename PackedOf<T>::type;
    constexpr int TW = 64, TH = 8, EW = TW + 4, EH = TH + 2;   // haloed x range tx0-2 .. tx0+65 (aligned words)
    constexpr int PITCH = EW + 2;                              // 70 voxels = 35 words per row
    __shared__ __align__(4) T ring[3][EH][PITCH];
    const int tx0 = (int)(blockIdx.x % tiles_x) * TW, ty0 = (int)(blockIdx.x / tiles_x) * TH;
    const int z0 = blockIdx.y * zchunk, z1 = min(z0 + zchunk, d);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;

    auto load_plane = [&](int z) {
        const T* p = nullptr;
        int64_t rs = ssh;
        if (z < 0) {
            if (halo_lo) { p = halo_lo; rs = w; }
            else if (border == MIE_BORDER_REPLICATE) p = src;
        } else if (z >= d) {
            if (halo_hi) { p = halo_hi; rs = w; }
            else if (border == MIE_BORDER_REPLICATE) p = src + (int64_t)(d - 1) * ssd;
        } else {
            p = src + (int64_t)z * ssd;
        }
        T(*slot)[PITCH] = ring[(z + 3) % 3];
        for (int i = threadIdx.x; i < EH * EW; i += 256) {
            const int r = i / EW, c = i - r * EW;
            // columns beyond the image + 1 only feed masked outputs: clamp them into border_index's domain
            const int sy = border_index(ty0 - 1 + r, h, border), sx = border_index(min(tx0 - 2 + c, w), w, border);
            slot[r][c] = (!p || sy < 0 || sx < 0) ? (T)0 : p[(int64_t)sy * rs + sx];
        }
    };
    // nine packed words of one plane around this lane's voxel pair: rows ly..ly+2, x-1 / x / x+1 neighbours
    auto plane_words = [&](int z, P* out) {
        const T(*slot)[PITCH] = ring[(z + 3) % 3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(&slot[ly + dy][0]) + lx;
            const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];   // voxels (x-2,x-1) (x,x+1) (x+2,x+3)
            out[dy * 3 + 0].v = __byte_perm(w0, w1, 0x5432);        // (x-1, x)
            out[dy * 3 + 1].v = w1;                                 // (x, x+1)
            out[dy * 3 + 2].v = __byte_perm(w1, w2, 0x5432);        // (x+1, x+2)
        }
    };

    load_plane(z0 - 1);
    load_plane(z0);
    __syncthreads();
    P win[3][9];   // win[k] = plane z - 1 + k of the current step (rotated by register renaming below)
    plane_words(z0 - 1, win[0]);
    plane_words(z0, win[1]);
    const int y = ty0 + ly, x = tx0 + 2 * lx;
    for (int zb = z0; zb < z1; zb += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int z = zb + u;
            if (z < z1) {   // uniform
                load_plane(z + 1);
                __syncthreads();
                plane_words(z + 1, win[(u + 2) % 3]);
                P v[27];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    v[k] = win[u % 3][k]; v[9 + k] = win[(u + 1) % 3][k]; v[18 + k] = win[(u + 2) % 3][k];
                }
                const uint32_t m = median_of<27, P>(v).v;
                if (y < h && x < w) {
                    T* o = dst + (int64_t)z * dsd + (int64_t)y * dsh + x;
                    if (x + 1 < w && ((reinterpret_cast<uintptr_t>(o) & 3) == 0)) {
                        *reinterpret_cast<uint32_t*>(o) = m;
                    } else {
                        o[0] = (T)(m & 0xFFFFu);
                        if (x + 1 < w) o[1] = (T)(m >> 16);
                    }
                }
                // no second barrier: the next step writes ring slot (z + 2) % 3, and only slot (z + 1) % 3
                // is still being read (the other two planes of the window are held in registers)
            }
        }
    }
}

template <typename T>
static int launch_median3d(const void* src, void* dst, int d, int h, int w, int64_t ssd, int64_t ssh, int64_t dsd,
                           int64_t dsh, const void* lo, const void* hi, int border, cudaStream_t st) {
    const int zchunk = d >= 64 ? 32 : (d >= 16 ? 8 : d);
    if constexpr (sizeof(T) == 2) {
        static const bool off = [] { const char* e = getenv("MIE_MEDIAN_NO_PACKED"); return e && e[0] == '1'; }();
        if (!off) {
            const int tiles_x = ceil_div(w, 64), tiles_y = ceil_div(h, 8);
            dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)ceil_div(d, zchunk));
            median3d_packed_kernel<T><<<grid, 256, 0, st>>>((const T*)src, (T*)dst, ssd, ssh, dsd, dsh, d, h, w, tiles_x,
                                                           tiles_y, zchunk, (const T*)lo, (const T*)hi, border);
            return check_launch();
        }
    }
    const int tiles_x = ceil_div(w, 32), tiles_y = ceil_div(h, 8);
    dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)ceil_div(d, zchunk));
    median3d_kernel<T><<<grid, 256, 0, st>>>((const T*)src, (T*)dst, ssd, ssh, dsd, dsh, d, h, w, tiles_x, tiles_y,
                                            zchunk, (const T*)lo, (const T*)hi, border);
    return check_launch();
}

}  // namespace mie

using namespace mie;

extern "C" {

int mie_median2d(const void* src, void* dst, int dtype, int64_t n, int h, int w, int64_t src_stride_n,
                 int64_t src_stride_h, int64_t dst_stride_n, int64_t dst_stride_h, int ky, int kx, int border,
                 void* stream) {
    int rc = check_planes(src, dst, n, h, w, src_stride_n, src_stride_h, dst_stride_n, dst_stride_h);
    if (rc) return rc;
    if (border != MIE_BORDER_CONSTANT && border != MIE_BORDER_REPLICATE && border != MIE_BORDER_REFLECT)
        return MIE_E_BORDER;
    if (ky <= 0 || kx <= 0 || !(ky & 1) || !(kx & 1)) return MIE_E_KERNEL;
    if (n == 0) return MIE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    MIE_DISPATCH_SRC(dtype, return dispatch_median2d<SrcT>(ky, kx, src, dst, n, h, w, src_stride_n, src_stride_h,
                                                          dst_stride_n, dst_stride_h, border, st));
    return MIE_OK;
}

int mie_median3d(const void* src, void* dst, int dtype, int d, int h, int w, int64_t src_stride_d,
                 int64_t src_stride_h, int64_t dst_stride_d, int64_t dst_stride_h, const void* halo_lo,
                 const void* halo_hi, int border, void* stream) {
    int rc = check_planes(src, dst, d, h, w, src_stride_d, src_stride_h, dst_stride_d, dst_stride_h);
    if (rc) return rc;
    if (border != MIE_BORDER_CONSTANT && border != MIE_BORDER_REPLICATE) return MIE_E_BORDER;
    if (d == 0) return MIE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    MIE_DISPATCH_SRC(dtype, return launch_median3d<SrcT>(src, dst, d, h, w, src_stride_d, src_stride_h, dst_stride_d,
                                                        dst_stride_h, halo_lo, halo_hi, border, st));
    return MIE_OK;
}

}  // extern "C"
