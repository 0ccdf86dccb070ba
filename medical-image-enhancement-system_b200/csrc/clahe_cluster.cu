// clahe_cluster.cu — standalone CLAHE (kornia semantics, 256 bins) in ONE launch: a thread-block cluster per image.
//
// The two-pass path (clahe_fast.cu) reads every pixel twice from HBM — once for the tile histograms, once for the
// interpolation — and writes 2 KB of packed cell tables per interpolation cell in a third launch: 6 B/pixel plus the
// tables for 4 algorithmic bytes.  Here a cluster of gh CTAs owns one image and CTA ty owns tile row ty:
//   1. its th x w slab of pixels arrives in shared memory through bulk copies (TMA, four mbarrier stages);
//   2. the slab's gw tile histograms are counted out of shared memory (ATOMS.POPC.INC), one warp per tile turns
//      them into LUTs (warp_build_lut: the clip / redistribute / cumulate rule of every other CLAHE kernel here);
//   3. one cluster barrier later every CTA copies the LUT rows of its two neighbour tile rows out of their shared
//      memory (distributed shared memory), packs the two rows of interpolation-cell tables its pixels fall into
//      (the packing of chain_pack_cells_kernel, over the dead histograms) and
//   4. interpolates its slab out of shared memory: one 8-byte table lookup per pixel, packed f32x2 blend
//      (clahe_px2) — the arithmetic, bit for bit, of clahe_apply_fast_kernel.
// A pixel crosses HBM once in each direction and nothing else is written but the LUTs (256 B per tile, kept in the
// workspace for callers of the stage API).  Geometry: unpadded, default-range integer pixels, w a power of two
// <= 2048, tile width a multiple of 8, even tile height, gh <= 8, slab + tables <= 200 KB of shared memory (two
// CTAs per SM up to 113 KB: the config-2 slices need 107 KB).  Everything else keeps the two-pass kernels.

#include <cooperative_groups.h>

#include "march.cuh"
#include "window.cuh"

namespace mie {

constexpr int kCcThreads = 512;
constexpr int kCcStages = 4;

struct ClusterClaheArgs {
    const void* src;
    void* dst;
    int64_t ssn, ssh, dsn, dsh;
    ClaheGeom g;
    LutParams lp;
    uint8_t* luts;   // [n][gh][gw][256]
};

template <typename T> struct SlabQuad;   // 4 / 8 pixel codes (v - dtype_min) out of shared memory
template <> struct SlabQuad<uint16_t> {
    static __device__ __forceinline__ void load4(const uint16_t* p, uint32_t* u) {
        const uint2 b = *reinterpret_cast<const uint2*>(p);
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
    }
    static __device__ __forceinline__ void load8(const uint16_t* p, uint32_t* u) {
        const uint4 b = *reinterpret_cast<const uint4*>(p);
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
        u[4] = b.z & 0xFFFFu; u[5] = b.z >> 16; u[6] = b.w & 0xFFFFu; u[7] = b.w >> 16;
    }
};
template <> struct SlabQuad<int16_t> {
    static __device__ __forceinline__ void load4(const int16_t* p, uint32_t* u) {
        uint2 b = *reinterpret_cast<const uint2*>(p);
        b.x ^= 0x80008000u; b.y ^= 0x80008000u;
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
    }
    static __device__ __forceinline__ void load8(const int16_t* p, uint32_t* u) {
        uint4 b = *reinterpret_cast<const uint4*>(p);
        b.x ^= 0x80008000u; b.y ^= 0x80008000u; b.z ^= 0x80008000u; b.w ^= 0x80008000u;
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
        u[4] = b.z & 0xFFFFu; u[5] = b.z >> 16; u[6] = b.w & 0xFFFFu; u[7] = b.w >> 16;
    }
};
template <> struct SlabQuad<uint8_t> {
    static __device__ __forceinline__ void load4(const uint8_t* p, uint32_t* u) {
        const uint32_t b = *reinterpret_cast<const uint32_t*>(p);
        u[0] = b & 0xFFu; u[1] = (b >> 8) & 0xFFu; u[2] = (b >> 16) & 0xFFu; u[3] = b >> 24;
    }
    static __device__ __forceinline__ void load8(const uint8_t* p, uint32_t* u) {
        const uint2 b = *reinterpret_cast<const uint2*>(p);
        u[0] = b.x & 0xFFu; u[1] = (b.x >> 8) & 0xFFu; u[2] = (b.x >> 16) & 0xFFu; u[3] = b.x >> 24;
        u[4] = b.y & 0xFFu; u[5] = (b.y >> 8) & 0xFFu; u[6] = (b.y >> 16) & 0xFFu; u[7] = b.y >> 24;
    }
};

__device__ __forceinline__ uint2 cc_lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// weight of the upper / left tile at local coordinate r of a tile of size T (kornia_axis, as clahe_fast.cu)
__device__ __forceinline__ float cc_axis_weight(int y, int T) {
    const int hh = T >> 1;
    int r = y - hh;
    r = r < 0 ? 0 : r % T;
    return __fdiv_rn((float)(T - 1 - r), (float)(T - 1));
}

// shared-memory carve-up (bytes), shared by the kernel and the host
struct CcLayout {
    size_t slab, tab, lut, wy, total;
};
__host__ __device__ inline CcLayout cc_layout(const ClaheGeom& g, int esz) {
    CcLayout L;
    L.slab = ((size_t)g.th * g.w * esz + 127) & ~(size_t)127;
    const size_t tables = (size_t)2 * (g.gw + 1) * kBins * 8, hists = (size_t)g.gw * kHistPitch * 4;
    L.tab = tables > hists ? tables : hists;          // histograms first, cell tables over them later
    L.lut = (size_t)3 * g.gw * kBins;                 // LUT rows ty - 1, ty, ty + 1
    L.wy = ((size_t)g.th * 4 + 15) & ~(size_t)15;
    L.total = L.slab + L.tab + L.lut + L.wy;
    return L;
}

template <typename SrcT, typename DstT>
__global__ void __launch_bounds__(kCcThreads)
clahe_cluster_kernel(const __grid_constant__ ClusterClaheArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_bar[kCcStages];
    const ClaheGeom g = a.g;
    const CcLayout L = cc_layout(g, (int)sizeof(SrcT));
    SrcT* slab = reinterpret_cast<SrcT*>(s_raw);
    int* s_hist = reinterpret_cast<int*>(s_raw + L.slab);
    uint2* s_tab = reinterpret_cast<uint2*>(s_raw + L.slab);
    uint8_t* s_lut = s_raw + L.slab + L.tab;                       // [3][gw][256]
    float* s_wy = reinterpret_cast<float*>(s_raw + L.slab + L.tab + L.lut);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = (int)cluster.block_rank();
    const int64_t n = blockIdx.x / g.gh;
    const int w = g.w, th = g.th, gw = g.gw;
    const uint32_t row_bytes = (uint32_t)w * (uint32_t)sizeof(SrcT);
    const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(s_bar);
    const int rows_per_stage = (th + kCcStages - 1) / kCcStages;

    for (int i = tid; i < gw * kHistPitch; i += kCcThreads) s_hist[i] = 0;
    for (int i = tid; i < th; i += kCcThreads) s_wy[i] = cc_axis_weight(ty * th + i, th);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kCcStages; ++s) mbar_init(bar32 + 8 * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane < kCcStages) {
            const int r0 = min(lane * rows_per_stage, th), r1 = min(r0 + rows_per_stage, th);
            mbar_expect_tx(bar32 + 8 * lane, (uint32_t)(r1 - r0) * row_bytes);
        }
        __syncwarp();
        const char* plane = reinterpret_cast<const char*>((const SrcT*)a.src + n * a.ssn + (int64_t)ty * th * a.ssh);
        const uint32_t slab32 = (uint32_t)__cvta_generic_to_shared(slab);
        for (int r = lane; r < th; r += 32)
            bulk_g2s(slab32 + (uint32_t)r * row_bytes, plane + (int64_t)r * a.ssh * (int64_t)sizeof(SrcT), row_bytes,
                     bar32 + 8 * (r / rows_per_stage));
    }

    // ---- 1. tile histograms out of the slab: a thread keeps its 8-pixel column group (and so its tile) for the walk
    {
        const int groups = w >> 3;                       // 8-pixel groups per row: a power of two <= 256
        const int gcol = tid % groups, r00 = tid / groups, rstep = kCcThreads / groups;
        int* my_hist = s_hist + ((8 * gcol) / g.tw) * kHistPitch;
        int r = r00;   // the thread's rows r00, r00 + rstep, ... simply continue from one stage into the next
        for (int s = 0; s < kCcStages; ++s) {
            const int r1 = min((s + 1) * rows_per_stage, th);
            if (s * rows_per_stage < th) mbar_wait(bar32 + 8 * s, 0u);   // every thread observes every stage: it reads other rows in step 4
#pragma unroll 2
            for (; r < r1; r += rstep) {
                uint32_t u[8];
                SlabQuad<SrcT>::load8(slab + (size_t)r * w + 8 * gcol, u);
#pragma unroll
                for (int k = 0; k < 8; ++k) hist_add_nobranch(my_hist, (int)Codes<SrcT>::bin(u[k]));
            }
        }
    }
    __syncthreads();

    // ---- 2. LUTs of this tile row (one warp per tile) -> s_lut[1] and the workspace
    uint8_t* my_luts = s_lut + (size_t)gw * kBins;
    for (int t = warp; t < gw; t += kCcThreads / 32) warp_build_lut<false>(s_hist + t * kHistPitch, a.lp, my_luts + t * kBins, lane);
    __syncthreads();
    {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(my_luts);
        uint32_t* g32 = reinterpret_cast<uint32_t*>(a.luts + (n * g.gh + ty) * (int64_t)gw * kBins);
        for (int i = tid; i < gw * kBins / 4; i += kCcThreads) g32[i] = s32[i];
    }

    // ---- 3. neighbour LUT rows through distributed shared memory, then the two rows of cell tables
    cluster.sync();
    {
        const int up = max(ty - 1, 0), dn = min(ty + 1, g.gh - 1);
        const uint32_t* pu = reinterpret_cast<const uint32_t*>(cluster.map_shared_rank(my_luts, up));
        const uint32_t* pd = reinterpret_cast<const uint32_t*>(cluster.map_shared_rank(my_luts, dn));
        uint32_t* d0 = reinterpret_cast<uint32_t*>(s_lut);
        uint32_t* d2 = reinterpret_cast<uint32_t*>(s_lut + (size_t)2 * gw * kBins);
        for (int i = tid; i < gw * kBins / 4; i += kCcThreads) { d0[i] = pu[i]; d2[i] = pd[i]; }
    }
    cluster.barrier_arrive();   // "I have read my neighbours": the wait stands in front of the exit
    __syncthreads();
    {   // table (cr, cx): kCcThreads / 256 tables per step, grey level = tid & 255
        const int gl = tid & (kBins - 1);
        for (int c = tid >> 8; c < 2 * (gw + 1); c += kCcThreads / kBins) {
            const int cr = c > gw ? 1 : 0, cx = c - cr * (gw + 1);
            const uint8_t* top = s_lut + (size_t)cr * gw * kBins + gl;      // cell row ty + cr: LUT rows ty + cr - 1, ty + cr
            const uint8_t* bot = top + (size_t)gw * kBins;
            const int il = max(cx - 1, 0), ir = min(cx, gw - 1);
            const int tl = top[il * kBins], tr = top[ir * kBins], bl = bot[il * kBins], br = bot[ir * kBins];
            s_tab[c * kBins + gl] = make_uint2(cell_word(tl - tr, tr), cell_word(bl - br, br));
        }
    }
    __syncthreads();

    // ---- 4. interpolation out of the slab: 4 consecutive columns per thread, kCcThreads / (w / 4) rows per step
    {
        const int quads = w >> 2;
        const int x0 = 4 * (tid % quads), rstep = kCcThreads / quads;
        const int cx = (x0 + (g.tw >> 1)) / g.tw;
        const uint32_t tb0 = (uint32_t)__cvta_generic_to_shared(s_tab + cx * kBins);
        const uint32_t tb_step = (uint32_t)((gw + 1) * kBins * 8);
        float wxv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wxv[k] = cc_axis_weight(x0 + k, g.tw);
        const f32x2 wxa = f2_pack(wxv[0], wxv[1]), wxb = f2_pack(wxv[2], wxv[3]);
        DstT* dp = (DstT*)a.dst + n * a.dsn + (int64_t)ty * th * a.dsh + x0;
        const WinCvt cv = {};
        const int half = th >> 1;
#pragma unroll 4
        for (int r = tid / quads; r < th; r += rstep) {
            const uint32_t tb = tb0 + (r >= half ? tb_step : 0u);
            const float wyv = s_wy[r];
            const f32x2 wy2 = f2_pack(wyv, wyv);
            uint32_t u[4];
            SlabQuad<SrcT>::load4(slab + (size_t)r * w + x0, u);
            const f32x2 ya = clahe_px2(cc_lds64(tb + Codes<SrcT>::entry_offset(u[0])), cc_lds64(tb + Codes<SrcT>::entry_offset(u[1])), wxa, wy2);
            const f32x2 yb = clahe_px2(cc_lds64(tb + Codes<SrcT>::entry_offset(u[2])), cc_lds64(tb + Codes<SrcT>::entry_offset(u[3])), wxb, wy2);
            float y[4];
            f2_unpack(ya, y[0], y[1]);
            f2_unpack(yb, y[2], y[3]);
            PixIO<DstT, false>::store4(dp + (int64_t)r * a.dsh, y, cv);
        }
    }
    cluster.barrier_wait();
}

static const int kCcEsz[4] = {1, 2, 2, 4};

bool clahe_cluster_ok(const ClaheGeom& g, int sd, int dd, const void* src, const void* dst, int64_t n, int64_t ssn,
                      int64_t ssh, int64_t dsn, int64_t dsh, float lo, float hi) {
    if (kernel_policy(MIE_POLICY_GENERIC_CLAHE) || kernel_policy(MIE_POLICY_CLAHE_FLOAT_RULES) ||
        kernel_policy(MIE_POLICY_CLAHE_TWO_PASS))
        return false;
    if (sd == MIE_F32 || !default_range_c(sd, lo, hi) || !int_rules_ok(sd)) return false;
    if (dd != sd && dd != MIE_F32) return false;
    if (g.hp != g.h || g.wp != g.w || (g.tw & 7) || (g.th & 1) || g.gh > 8 || g.gw > 32) return false;
    if (g.w < 32 || g.w > 2048 || (g.w & (g.w - 1))) return false;   // w / 4 and w / 8 divide the block
    if ((int64_t)g.th * g.tw >= (1 << 24) || n * g.gh > 2147483647LL) return false;
    const int e = kCcEsz[sd], de = kCcEsz[dd];
    if (((uintptr_t)src % 16) || ((ssn * e) % 16) || ((ssh * e) % 16) || (((int64_t)g.w * e) % 16)) return false;
    if (((uintptr_t)dst % 16) || ((dsn * de) % 16) || ((dsh * de) % 16)) return false;
    return cc_layout(g, e).total <= 200 * 1024;
}

template <typename SrcT, typename DstT>
static int launch_cc(const ClusterClaheArgs& a, int64_t n, cudaStream_t st) {
    const size_t smem = cc_layout(a.g, (int)sizeof(SrcT)).total;
    MIE_ENSURE_SMEM((clahe_cluster_kernel<SrcT, DstT>), 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * a.g.gh));
    cfg.blockDim = dim3(kCcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)a.g.gh;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, clahe_cluster_kernel<SrcT, DstT>, a);
    return e == cudaSuccess ? check_launch() : (int)e;
}

int launch_clahe_cluster(const void* src, void* dst, int sd, int dd, int64_t n, int64_t ssn, int64_t ssh, int64_t dsn,
                         int64_t dsh, const ClaheGeom& g, const LutParams& lp, uint8_t* luts, cudaStream_t st) {
    if (n == 0) return MIE_OK;
    ClusterClaheArgs a;
    a.src = src; a.dst = dst; a.ssn = ssn; a.ssh = ssh; a.dsn = dsn; a.dsh = dsh; a.g = g; a.lp = lp; a.luts = luts;
    switch (sd) {
        case MIE_U8: return dd == MIE_F32 ? launch_cc<uint8_t, float>(a, n, st) : launch_cc<uint8_t, uint8_t>(a, n, st);
        case MIE_U16: return dd == MIE_F32 ? launch_cc<uint16_t, float>(a, n, st) : launch_cc<uint16_t, uint16_t>(a, n, st);
        case MIE_I16: return dd == MIE_F32 ? launch_cc<int16_t, float>(a, n, st) : launch_cc<int16_t, int16_t>(a, n, st);
        default: return MIE_E_DTYPE;
    }
}

}  // namespace mie
