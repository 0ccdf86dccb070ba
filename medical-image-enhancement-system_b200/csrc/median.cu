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

template <typename T>
static int launch_median3d(const void* src, void* dst, int d, int h, int w, int64_t ssd, int64_t ssh, int64_t dsd,
                           int64_t dsh, const void* lo, const void* hi, int border, cudaStream_t st) {
    const int tiles_x = ceil_div(w, 32), tiles_y = ceil_div(h, 8);
    const int zchunk = d >= 64 ? 32 : (d >= 16 ? 8 : d);
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
