// chain_march.cu — "marching" kernels of the fused Gaussian -> CLAHE -> unsharp chain for the
// headline geometry (BASELINE.json config 2: 64x64-pixel CLAHE tiles, 9-tap Gaussians, image width a
// multiple of 128 and <= 1024).  Same arithmetic, bit for bit, as chain.cu / chain_fast.cu and
// oracle/mie_oracle.c; what changes is the schedule:
//
//   * one block owns a full-width band of 64 rows (one row of CLAHE tiles) and walks down its
//     64 + 8 source rows TWO ROWS PER STEP; thread t owns image columns 4t .. 4t+3 for the whole walk;
//   * a row is converted (chain_a) / CLAHE-blended (chain_b) exactly once per band and exchanged
//     through a small ring of row-pair buffers in shared memory in which the two rows of a pair are
//     interleaved per column — so the horizontal pass reads ready-made (row s, row s+1) operand pairs
//     and runs on packed fma.rn.f32x2 (one issue slot per two results);
//   * the vertical pass never touches shared memory: the horizontally filtered rows of the thread's
//     four columns live in a register ring (16 slots, ten of them live; the walk is unrolled by eight
//     row pairs so that the ring is addressed statically) and feed packed fma.rn.f32x2 on column pairs;
//   * nothing is computed for a horizontal halo (the image-border halo is four warp shuffles per row
//     pair), and only 8 of 72 rows are vertical halo: per-pixel work is 1.125x the arithmetic minimum;
//   * global loads (4 pixels / 4 index bytes per thread and row) are issued four rows at a time, four
//     to seven rows ahead of their use; source-row offsets (mirrored at the image border) come from a
//     small shared-memory table instead of per-row index arithmetic.

#include "march.cuh"
#include "window.cuh"

#ifndef MIE_A_MINB
#define MIE_A_MINB 5
#endif
#ifndef MIE_B_AHEAD
#define MIE_B_AHEAD 6
#endif

namespace mie {

// ================================================================ chain_a (marching)
// grid = n * gh blocks of W/4 threads.  Per row pair: load (ahead) -> convert -> pair buffer ->
// horizontal pass -> register ring -> vertical pass of the two rows that became complete -> lookup index
// (32-bit store into the index plane) and histogram bin (ATOMS.POPC.INC into the thread's tile
// histogram).  After the walk each warp turns tile histograms into LUTs.
// LE1: the host has proven that every blurred value lies in [0, 1] (integer pixels and
// gauss_of_ones_le1), which removes the range tests from the index / bin rules.
// WIN: integer value_range window (window.cuh): windowed conversion of the source pixels; the blurred values can
// then lie outside [0, 1], so bins and lookups take the range-checked rules of float pixels.
template <typename SrcT, int BORDER, bool LE1, int MAXT, int MINB, bool WIN = false>
__global__ void __launch_bounds__(MAXT, MINB)
chain_a_march_kernel(ChainAArgs a, Taps wx, Taps wy, WinCvt cv) {
    typedef typename Fast<SrcT>::raw4 raw4;
    constexpr bool NN = !(sizeof(SrcT) == 4) && !WIN;
    static_assert(!(LE1 && WIN), "a window gives no [0, 1] guarantee");
    static_assert(!LE1 || NN, "LE1 needs integer pixels");
    extern __shared__ __align__(16) float smem[];
    const int W = a.g.w, T = blockDim.x, gw = a.g.gw, h = a.g.h;
    const int pbuf = pairbuf_floats(T);
    float* s_buf = smem;                                              // 2 pair buffers
    int* s_hist = reinterpret_cast<int*>(smem + 2 * pbuf);            // gw x kHistPitch
    int* s_off = s_hist + gw * kHistPitch;                            // kMOffRows
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_off + kMOffRows);  // kRawBars mbarriers
    float* s_raw = reinterpret_cast<float*>(s_bar + kRawBars);        // kRawRows source rows (raw pixels)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int ty = (int)(blockIdx.x % a.g.gh);
    const int64_t n = blockIdx.x / a.g.gh;
    const int ty0 = ty * kTile;
    for (int i = tid; i < gw * kHistPitch; i += T) s_hist[i] = 0;
    fill_row_offsets<BORDER>(s_off, ty0, h, (int)a.ssh * (int)sizeof(SrcT), tid, T);
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < kRawBars; ++b)
            mbar_init((uint32_t)__cvta_generic_to_shared(s_bar + b), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int* const my_hist = s_hist + (tid >> 4) * kHistPitch;
    const uint32_t my_hist32 = hist_base32(my_hist);
    uint8_t* ip = a.idx + (n * h + ty0) * (int64_t)W + 4 * tid;
    const bool first_warp = warp == 0, last_warp = warp == nwarps - 1;

    // source rows through the bulk-copy ring: batch b = band rows 4b .. 4b+3 -> ring slots (4b + j) % 16,
    // mbarrier b % 4, phase parity (b / 4) & 1
    const int row_bytes = W * (int)sizeof(SrcT);
    const uint32_t ring32 = (uint32_t)__cvta_generic_to_shared(s_raw);
    const uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(s_bar);
    const char* plane0 = reinterpret_cast<const char*>((const SrcT*)a.src + n * a.ssn);
    // the producer is the first lane of warp 1 when there is one: warps 0 and nwarps-1 already carry the
    // image-border halo work, and every warp waits for the slowest one at the per-step barrier
    const int producer = nwarps > 2 ? 32 : 0;
    const uint64_t pol = l2_evict_first_policy();
    auto issue_batch = [&](const int b) {  // producer thread only
        const int4 o = *reinterpret_cast<const int4*>(s_off + kRawBatch * b);
        const int off[4] = {o.x, o.y, o.z, o.w};
        const uint32_t mb = bar32 + 8 * (b % kRawBars);
        int valid = 0;
#pragma unroll
        for (int j = 0; j < kRawBatch; ++j) valid += (BORDER != MIE_BORDER_CONSTANT || off[j] >= 0) ? 1 : 0;
        mbar_expect_tx(mb, (uint32_t)(valid * row_bytes));
#pragma unroll
        for (int j = 0; j < kRawBatch; ++j)
            if (BORDER != MIE_BORDER_CONSTANT || off[j] >= 0)
                bulk_g2s_stream(ring32 + (uint32_t)(((kRawBatch * b + j) % kRawRows) * row_bytes), plane0 + (unsigned)off[j],
                                (uint32_t)row_bytes, mb, pol);
    };
    if (tid == producer) {
#pragma unroll
        for (int b = 0; b < kRawBars; ++b) issue_batch(b);
    }
    f32x2 ring[kMRing][2];
    const char* my_raw = reinterpret_cast<const char*>(s_raw) + 4 * tid * (int)sizeof(SrcT);

    // rows 2p, 2p+1 (ring slots rslot, rslot + 1); p even = first use of batch p / 2
    // -> packed (row 2p, row 2p + 1) pairs of the thread's four columns: the layout of the pair buffer
    uint32_t exp_magic = 0u;
    if constexpr (sizeof(SrcT) == 2 && !WIN) exp_magic = Fast<SrcT>::exp_magic();
    auto convert = [&](const int p, const int rslot, f32x2* xp) {
        if (p % 2 == 0) mbar_wait(bar32 + 8 * ((p / 2) % kRawBars), (uint32_t)((p / 2 / kRawBars) & 1));
        const raw4 r0 = *reinterpret_cast<const raw4*>(my_raw + (rslot % kRawRows) * row_bytes);
        const raw4 r1 = *reinterpret_cast<const raw4*>(my_raw + ((rslot + 1) % kRawRows) * row_bytes);
        if constexpr (sizeof(SrcT) == 2 && !WIN) {
            Fast<SrcT>::cvt_pair4(r0, r1, xp, exp_magic);   // 16-bit default range: the conversion itself runs packed
            if (BORDER == MIE_BORDER_CONSTANT) {
                const bool z0 = s_off[2 * p] < 0, z1 = s_off[2 * p + 1] < 0;
                if (z0 || z1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float a, b;
                        f2_unpack(xp[k], a, b);
                        xp[k] = f2_pack(z0 ? 0.0f : a, z1 ? 0.0f : b);
                    }
                }
            }
        } else {
            float x0[4], x1[4];
            PixIO<SrcT, WIN>::cvt_raw4(r0, x0, cv);
            PixIO<SrcT, WIN>::cvt_raw4(r1, x1, cv);
            if (BORDER == MIE_BORDER_CONSTANT) {
                if (s_off[2 * p] < 0) x0[0] = x0[1] = x0[2] = x0[3] = 0.0f;
                if (s_off[2 * p + 1] < 0) x1[0] = x1[1] = x1[2] = x1[3] = 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) xp[k] = f2_pack(x0[k], x1[k]);
        }
    };
    // after the barrier of an odd row pair p its batch (p - 1) / 2 is consumed: refill the slots
    auto refill = [&](const int p) {
        const int b = (p - 1) / 2 + kRawBars;
        if (tid == producer && b < kMRows / kRawBatch) issue_batch(b);
    };
    auto emit = [&](const float* g) {
        if (LE1) {
            const f32x2 ga = f2_pack(g[0], g[1]), gb = f2_pack(g[2], g[3]);   // as the column pass produced them
            uint32_t b0, b1, b2, b3;
            idx_bits_le1_x2(ga, b0, b1);
            idx_bits_le1_x2(gb, b2, b3);
            *reinterpret_cast<uint32_t*>(ip) = pack_low_bytes(b0, b1, b2, b3);
            hist_add_le1_x2(my_hist32, ga);
            hist_add_le1_x2(my_hist32, gb);
        } else {
            *reinterpret_cast<uint32_t*>(ip) =
                pack_low_bytes(fast_idx_bits<NN>(g[0]), fast_idx_bits<NN>(g[1]), fast_idx_bits<NN>(g[2]),
                               fast_idx_bits<NN>(g[3]));
#pragma unroll
            for (int k = 0; k < 4; ++k) hist_add_nobranch(my_hist, fast_bin<NN>(g[k]));
        }
        ip += W;
    };

    // ---- prologue: row pairs 0..3 fill ring slots 0..7
#pragma unroll
    for (int p = 0; p < kMPro; ++p) {
        f32x2 xp[4];
        convert(p, 2 * p, xp);
        float* buf = s_buf + (p % 2) * pbuf;
        pair_store_packed<BORDER>(buf, T, tid, first_warp, last_warp, xp);
        __syncthreads();
        if (p % 2 == 1) refill(p);
        pair_row_pass(buf, T, tid, wx, ring, 2 * p, xp);
    }
    // ---- main loop: 4 x 8 row pairs; pair q of an iteration = rows 8 + 16 i + 2q (+1) -> ring slots
    //      8 + 2q (+1); the two rows that become complete are the ones eight rows above
    for (int p0 = kMPro; p0 < kMPairs; p0 += kMUnroll) {
#pragma unroll
        for (int q = 0; q < kMUnroll; ++q) {
            const int p = p0 + q;
            f32x2 xp[4];
            convert(p, 2 * kMPro + 2 * q, xp);
            float* buf = s_buf + (q % 2) * pbuf;
            pair_store_packed<BORDER>(buf, T, tid, first_warp, last_warp, xp);
            __syncthreads();
            if (q % 2 == 1) refill(p);
            pair_row_pass(buf, T, tid, wx, ring, 2 * kMPro + 2 * q, xp);
            float g[4];
            march_col_pass(ring, 2 * q, wy, g);
            emit(g);
            march_col_pass(ring, 2 * q + 1, wy, g);
            emit(g);
        }
    }
    __syncthreads();
    uint8_t* luts = a.luts + (n * a.g.gh + ty) * (int64_t)gw * kBins;
    for (int t = warp; t < gw; t += nwarps)
        warp_build_lut<LE1>(s_hist + t * kHistPitch, a.lp, luts + t * kBins, lane);
}

// ================================================================ chain_b (marching, 9-tap unsharp)
// grid = n * gh blocks of W/4 threads.  The band's 72 source rows lie in two rows of interpolation
// cells; their 2 (gw+1) cell tables (2 KB each, contiguous in the workspace) are staged in shared
// memory once.  Per row pair: index bytes (ahead) -> one 64-bit table lookup per pixel -> CLAHE output
// C -> pair buffer -> horizontal pass -> register ring -> vertical pass -> C + (C - blur(C)) -> quantise
// -> one 64-bit store per thread and row.  The C values of the last four row pairs stay in the
// pair-buffer ring, which is where the epilogue re-reads its centre pixels.
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

template <typename DstT, int BORDER, bool WIN = false>
__global__ void __launch_bounds__(256)
chain_b_march_kernel(ChainBArgs a, const uint2* __restrict__ cells, AxisWeights aw, Taps wx, Taps wy, WinCvt cv) {
    extern __shared__ __align__(16) float smem[];
    const int W = a.g.w, T = blockDim.x, gw = a.g.gw, gh = a.g.gh, h = a.g.h;
    const int pbuf = pairbuf_floats(T);
    uint2* s_tab = reinterpret_cast<uint2*>(smem);            // 2 x (gw+1) x 256 entries
    float* s_buf = smem + 2 * (gw + 1) * kBins * 2;           // 4 pair buffers
    int* s_off = reinterpret_cast<int*>(s_buf + 4 * pbuf);    // kMOffRows
    unsigned long long* s_tbar = reinterpret_cast<unsigned long long*>(s_off + kMOffRows);   // 1 mbarrier (8-byte aligned)

    const int tid = threadIdx.x, warp = tid >> 5, nwarps = T >> 5;
    const int ty = (int)(blockIdx.x % gh);
    const int64_t n = blockIdx.x / gh;
    const int ty0 = ty * kTile;
    const bool first_warp = warp == 0, last_warp = warp == nwarps - 1;

    fill_row_offsets<BORDER>(s_off, ty0, h, W, tid, T);
    // The band's two rows of cell tables are contiguous in the workspace (36.9 KB for gw = 8): ONE bulk copy (TMA) by one
    // thread brings them in while the block computes its row offsets and issues its first index loads.  The
    // LDG -> STS loop it replaces cost every thread 18 load / store round trips before the walk could start (8 % of
    // the kernel's stall samples, profiles/r2_ncu_full_chain_kernels_v7.txt).
    const uint32_t tbar32 = (uint32_t)__cvta_generic_to_shared(s_tbar);
    if (tid == 0) {
        mbar_init(tbar32, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = (uint32_t)(2 * (gw + 1) * kBins * 8);
        mbar_expect_tx(tbar32, bytes);
        // bulk copies are limited to sizes the hardware counter can hold; chunks of 16 KB keep every size legal
        const char* src = reinterpret_cast<const char*>(cells + (n * (gh + 1) + ty) * (int64_t)(gw + 1) * kBins);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_tab);
        for (uint32_t o = 0; o < bytes; o += 16384u)
            bulk_g2s(dst + o, src + o, min(16384u, bytes - o), tbar32);
    }
    __syncthreads();

    const uint8_t* iplane = a.idx + n * (int64_t)h * W + 4 * tid;
    // raw[p % kRawSlots][j] = index word of row 2p + j, loaded kAhead row pairs ahead of its use.
    // ptxas tracks ALL of this kernel's global loads on ONE scoreboard (SASS control words: every LDG of the loop
    // writes barrier 5), and a scoreboard is a counter: the first read of ANY loaded word waits until EVERY
    // outstanding load has landed, including a group issued a few instructions earlier — which exposed the full
    // memory latency four times per loop iteration however far ahead the loads were issued (ncu: long_scoreboard was
    // the top stall, 1.02 warps per issue).  So the order is forced: first DRAIN (read one word of the group that is
    // due: everything outstanding is at least two row pairs old by then), and only then issue the next group, whose
    // addresses depend on the drained word through `zero` — a value that is always 0 (threadIdx.y of a 1-D block) but
    // that the compiler cannot know — so the loads cannot be hoisted above the drain.
    constexpr int kAhead = MIE_B_AHEAD;   // row pairs between an index load and its use (even)
    constexpr int kRawSlots = kAhead <= 2 ? 4 : 8;
    uint32_t raw[kRawSlots][2];
    const uint32_t zero = threadIdx.y;
    auto fetch_pairs = [&](const int p, const int slot, const uint32_t dep) {
        const int4 o = *reinterpret_cast<const int4*>(s_off + 2 * p);
        const int off[4] = {o.x, o.y, o.z, o.w};
        const uint8_t* base = iplane + dep;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v = 0u;
            if (BORDER != MIE_BORDER_CONSTANT || off[j] >= 0)
                v = __ldg(reinterpret_cast<const uint32_t*>(base + (unsigned)off[j]));
            raw[(slot + j / 2) % kRawSlots][j & 1] = v;
        }
    };
    // drain: a real read of the due word; the result is always 0
    auto drain = [&](const int slot) -> uint32_t { return raw[slot][0] & zero; };
#pragma unroll
    for (int p = 0; p < kAhead; p += 2) fetch_pairs(p, p % kRawSlots, 0u);

    // cells of this thread's columns: column cell (4t + 32) / 64; row cell 0 (rows above the tile
    // centre line ty0 + 32, i.e. band rows < 36) or 1.  Shared-window byte addresses: lookup = LEA + LDS.64.
    const uint32_t tb0 = (uint32_t)__cvta_generic_to_shared(s_tab + ((tid + 8) >> 4) * kBins);
    const uint32_t tb_step = (uint32_t)((gw + 1) * kBins * 8);
    float wxv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wxv[k] = aw.w[4 + ((4 * tid + k) & (kTile - 1))];
    DstT* op = (DstT*)a.dst + n * a.dsn + (int64_t)ty0 * a.dsh + 4 * tid;
    const int dsh = (int)a.dsh;
    f32x2 ring[kMRing][2];
    mbar_wait(tbar32, 0u);   // cell tables have landed

    // both rows of a pair lie in the same cell row (the switch is at band row 36); the two rows of column k are blended
    // as one packed pair — which is also the layout of the pair buffer
    f32x2 wx2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wx2[k] = f2_pack(wxv[k], wxv[k]);
    auto clahe_pair = [&](const int p, const int pslot, f32x2* xp) {
        const uint32_t tb = tb0 + (2 * p >= kTile / 2 + kMR ? tb_step : 0u);
        const uint32_t i0 = raw[pslot][0], i1 = raw[pslot][1];
        const f32x2 wy2 = f2_pack(aw.w[2 * p], aw.w[2 * p + 1]);
        xp[0] = clahe_px2(lds64(tb + (__byte_perm(i0, 0u, 0x4440) << 3)), lds64(tb + (__byte_perm(i1, 0u, 0x4440) << 3)), wx2[0], wy2);
        xp[1] = clahe_px2(lds64(tb + (__byte_perm(i0, 0u, 0x4441) << 3)), lds64(tb + (__byte_perm(i1, 0u, 0x4441) << 3)), wx2[1], wy2);
        xp[2] = clahe_px2(lds64(tb + (__byte_perm(i0, 0u, 0x4442) << 3)), lds64(tb + (__byte_perm(i1, 0u, 0x4442) << 3)), wx2[2], wy2);
        xp[3] = clahe_px2(lds64(tb + (__byte_perm(i0, 0u, 0x4443) << 3)), lds64(tb + (__byte_perm(i1, 0u, 0x4443) << 3)), wx2[3], wy2);
        if (BORDER == MIE_BORDER_CONSTANT) {
            const bool z0 = s_off[2 * p] < 0, z1 = s_off[2 * p + 1] < 0;
            if (z0 || z1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a, b;
                    f2_unpack(xp[k], a, b);
                    xp[k] = f2_pack(z0 ? 0.0f : a, z1 ? 0.0f : b);
                }
            }
        }
    };

#pragma unroll
    for (int p = 0; p < kMPro; ++p) {
        if (p % 2 == 0) fetch_pairs(p + kAhead, (p + kAhead) % kRawSlots, drain(p % kRawSlots));
        f32x2 xp[4];
        clahe_pair(p, p % kRawSlots, xp);
        float* buf = s_buf + (p % 4) * pbuf;
        pair_store_packed<BORDER>(buf, T, tid, first_warp, last_warp, xp);
        __syncthreads();
        pair_row_pass(buf, T, tid, wx, ring, 2 * p, xp);
    }
    for (int p0 = kMPro; p0 < kMPairs; p0 += kMUnroll) {
#pragma unroll
        for (int q = 0; q < kMUnroll; ++q) {
            const int p = p0 + q;
            if (q % 2 == 0) fetch_pairs(p + kAhead, (kMPro + q + kAhead) % kRawSlots, drain((kMPro + q) % kRawSlots));
            f32x2 xp[4];
            clahe_pair(p, (kMPro + q) % kRawSlots, xp);
            float* buf = s_buf + (q % 4) * pbuf;
            pair_store_packed<BORDER>(buf, T, tid, first_warp, last_warp, xp);
            __syncthreads();
            pair_row_pass(buf, T, tid, wx, ring, 2 * kMPro + 2 * q, xp);
            // centre pixels of the two completed rows: rows (2p - 4, 2p - 3) = pair p - 2
            const float* cbuf = s_buf + ((q + 2) % 4) * pbuf;
            const float4 ca = *reinterpret_cast<const float4*>(cbuf + 4 * (tid + 1));
            const float4 cb = *reinterpret_cast<const float4*>(cbuf + 4 * (T + 2) + 4 * (tid + 1));
            const float c0[4] = {ca.x, ca.z, cb.x, cb.z}, c1[4] = {ca.y, ca.w, cb.y, cb.w};
            float g[4], y[4];
            march_col_pass(ring, 2 * q, wy, g);
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = __fadd_rn(c0[k], __fsub_rn(c0[k], g[k]));
            PixIO<DstT, WIN>::store4(op, y, cv);
            op += dsh;
            march_col_pass(ring, 2 * q + 1, wy, g);
#pragma unroll
            for (int k = 0; k < 4; ++k) y[k] = __fadd_rn(c1[k], __fsub_rn(c1[k], g[k]));
            PixIO<DstT, WIN>::store4(op, y, cv);
            op += dsh;
        }
    }
}

// ================================================================ host side
bool march_chain_ok(const ClaheGeom& g, int kg, int ku) {
    if (kg != 9 || ku != 9) return false;
    if (g.th != kTile || g.tw != kTile || g.hp != g.h || g.wp != g.w) return false;
    if (g.w % 128 != 0 || g.w > 1024) return false;
    return true;
}

static size_t march_a_smem(const ClaheGeom& g, int elem_bytes) {
    return (size_t)(2 * 8 * (g.w / 4 + 2) + g.gw * kHistPitch + kMOffRows) * 4 + kRawBars * 8 +
           (size_t)kRawRows * g.w * elem_bytes;
}
static size_t march_b_smem(const ClaheGeom& g) {
    return (size_t)(2 * (g.gw + 1) * kBins * 2 + 4 * 8 * (g.w / 4 + 2) + kMOffRows) * 4 + 8;
}

// Blur of an all-ones image in the kernels' operation order.  Every fma is monotone in its data
// operands (the weights are >= 0), so for pixels in [0, 1] no blurred value exceeds this one.
static bool gauss_of_ones_le1(const Taps& wx, const Taps& wy) {
    for (int t = 0; t < 2 * kMR + 1; ++t)
        if (!(wx.w[t] >= 0.0f) || !(wy.w[t] >= 0.0f)) return false;
    float r = wx.w[0] * 1.0f;
    for (int t = 1; t < 2 * kMR + 1; ++t) r = fmaf(wx.w[t], 1.0f, r);
    float c = wy.w[0] * r;
    for (int t = 1; t < 2 * kMR + 1; ++t) c = fmaf(wy.w[t], r, c);
    return r <= 1.0f && c <= 1.0f;
}

template <typename SrcT, int BORDER>
static int launch_a_march_win(const ChainAArgs& a, const Taps& wx, const Taps& wy, unsigned blocks, cudaStream_t st,
                              const WinCvt& cv) {
    const size_t smem = march_a_smem(a.g, (int)sizeof(SrcT));
    if (a.g.w <= 512) {
        MIE_ENSURE_SMEM((chain_a_march_kernel<SrcT, BORDER, false, 128, 5, true>), 100 * 1024);
        chain_a_march_kernel<SrcT, BORDER, false, 128, 5, true><<<blocks, a.g.w / 4, smem, st>>>(a, wx, wy, cv);
    } else {
        MIE_ENSURE_SMEM((chain_a_march_kernel<SrcT, BORDER, false, 256, 2, true>), 100 * 1024);
        chain_a_march_kernel<SrcT, BORDER, false, 256, 2, true><<<blocks, a.g.w / 4, smem, st>>>(a, wx, wy, cv);
    }
    return check_launch();
}
template <typename SrcT, int BORDER, bool LE1>
static int launch_a_march_tbl(const ChainAArgs& a, const Taps& wx, const Taps& wy, unsigned blocks, cudaStream_t st) {
    const WinCvt cv = {};
    // W <= 512: 128 threads per block, registers capped so that five blocks fit on an SM (4 and 6 measured slower)
    const size_t smem = march_a_smem(a.g, (int)sizeof(SrcT));
    if (a.g.w <= 512) {
        MIE_ENSURE_SMEM((chain_a_march_kernel<SrcT, BORDER, LE1, 128, MIE_A_MINB>), 100 * 1024);
        chain_a_march_kernel<SrcT, BORDER, LE1, 128, MIE_A_MINB><<<blocks, a.g.w / 4, smem, st>>>(a, wx, wy, cv);
    } else {
        MIE_ENSURE_SMEM((chain_a_march_kernel<SrcT, BORDER, LE1, 256, 2>), 100 * 1024);
        chain_a_march_kernel<SrcT, BORDER, LE1, 256, 2><<<blocks, a.g.w / 4, smem, st>>>(a, wx, wy, cv);
    }
    return check_launch();
}
template <typename SrcT, int BORDER>
static int launch_a_march_tb(const ChainAArgs& a, const Taps& wx, const Taps& wy, unsigned blocks, cudaStream_t st,
                             const WinCvt* win) {
    if (win) return launch_a_march_win<SrcT, BORDER>(a, wx, wy, blocks, st, *win);
    if constexpr (sizeof(SrcT) != 4) {
        if (gauss_of_ones_le1(wx, wy)) return launch_a_march_tbl<SrcT, BORDER, true>(a, wx, wy, blocks, st);
    }
    return launch_a_march_tbl<SrcT, BORDER, false>(a, wx, wy, blocks, st);
}
template <typename SrcT>
static int launch_a_march_t(const ChainAArgs& a, const Taps& wx, const Taps& wy, unsigned blocks, cudaStream_t st,
                            const WinCvt* win) {
    switch (a.border) {
        case MIE_BORDER_REFLECT: return launch_a_march_tb<SrcT, MIE_BORDER_REFLECT>(a, wx, wy, blocks, st, win);
        case MIE_BORDER_REPLICATE: return launch_a_march_tb<SrcT, MIE_BORDER_REPLICATE>(a, wx, wy, blocks, st, win);
        default: return launch_a_march_tb<SrcT, MIE_BORDER_CONSTANT>(a, wx, wy, blocks, st, win);
    }
}

int launch_chain_a_march(const ChainAArgs& a, int sd, const Taps& wx, const Taps& wy, int64_t n, cudaStream_t st,
                         const WinCvt* win) {
    const unsigned blocks = (unsigned)(n * a.g.gh);
    MIE_DISPATCH_SRC(sd, return launch_a_march_t<SrcT>(a, wx, wy, blocks, st, win));
    return MIE_OK;
}

template <typename DstT, int BORDER>
static int launch_b_march_tb(const ChainBArgs& b, const uint2* cells, const AxisWeights& aw, const Taps& wx,
                             const Taps& wy, unsigned blocks, cudaStream_t st, const WinCvt* win) {
    if (win) {
        MIE_ENSURE_SMEM((chain_b_march_kernel<DstT, BORDER, true>), 128 * 1024);
        chain_b_march_kernel<DstT, BORDER, true><<<blocks, b.g.w / 4, march_b_smem(b.g), st>>>(b, cells, aw, wx, wy, *win);
    } else {
        const WinCvt cv = {};
        MIE_ENSURE_SMEM((chain_b_march_kernel<DstT, BORDER, false>), 128 * 1024);
        chain_b_march_kernel<DstT, BORDER, false><<<blocks, b.g.w / 4, march_b_smem(b.g), st>>>(b, cells, aw, wx, wy, cv);
    }
    return check_launch();
}
template <typename DstT>
static int launch_b_march_t(const ChainBArgs& b, const uint2* cells, const AxisWeights& aw, const Taps& wx,
                            const Taps& wy, unsigned blocks, cudaStream_t st, const WinCvt* win) {
    switch (b.border) {
        case MIE_BORDER_REFLECT: return launch_b_march_tb<DstT, MIE_BORDER_REFLECT>(b, cells, aw, wx, wy, blocks, st, win);
        case MIE_BORDER_REPLICATE: return launch_b_march_tb<DstT, MIE_BORDER_REPLICATE>(b, cells, aw, wx, wy, blocks, st, win);
        default: return launch_b_march_tb<DstT, MIE_BORDER_CONSTANT>(b, cells, aw, wx, wy, blocks, st, win);
    }
}

int launch_chain_b_march(const ChainBArgs& b, int dd, void* cells_raw, const Taps& wx, const Taps& wy, int64_t n,
                         cudaStream_t st, const WinCvt* win) {
    int rc = launch_pack_cells(b.luts, cells_raw, n, b.g.gh, b.g.gw, st);
    if (rc) return rc;
    AxisWeights aw;
    fill_axis_weights(aw);
    const unsigned blocks = (unsigned)(n * b.g.gh);
    MIE_DISPATCH_SRC(dd, return launch_b_march_t<SrcT>(b, (const uint2*)cells_raw, aw, wx, wy, blocks, st, win));
    return MIE_OK;
}

}  // namespace mie
