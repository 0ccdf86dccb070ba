// window.cuh — integer value_range windows for the tuned kernels: the divide-free windowed pixel conversion
// (device) and its exhaustive host-side verification.  See the comment on WinCvt.
#pragma once

#include <cmath>
#include <map>
#include <mutex>
#include <tuple>

#include "chain_fast.cuh"

namespace mie {

// ---------------------------------------------------------------- integer windows (value_range)
// x01 = (float(v) - lo) / (hi - lo) for an integer window [lo, hi] inside the dtype's range (e.g. the HU
// window -1024 .. 3071 of int16 CT data) without I2F and without an IEEE divide:
//   a  = uint_as_float(0x4B000000 | (v + bias)) - (2^23 + bias + lo)          exact: float(v) - lo
//   q0 = a * r;  x01 = fma(fma(-rg, q0, a), r, q0)        with r = RN(1 / rg) (Markstein's correction)
// The host checks this against the IEEE quotient for EVERY code of the dtype (65 536 at most) the first
// time a window is used and caches the verdict; a window that fails keeps the generic kernels.  Pixels
// outside the window land outside [0, 1]: not counted by the histogram (torch.histc rule), clamped by the
// lookup — the float-input rules, so windowed kernels take the float code path after the conversion.
// Back: rint(clamp(y, 0, 1) * rg) + lo as an integer add on the mantissa of (c * rg + 2^23).
struct WinCvt {
    float in_magic, r, neg_rg, rg;
    int lo_out;
};

template <typename T, bool WIN>
struct PixIO {   // WIN == false: the default-range conversions of chain_fast.cuh
    static __device__ __forceinline__ void load8(const T* p, float* x, const WinCvt&) { Fast<T>::load8(p, x); }
    static __device__ __forceinline__ void load4(const T* p, float* x, const WinCvt&) { Fast<T>::load4(p, x); }
    static __device__ __forceinline__ void store4(T* p, const float* y, const WinCvt&) { Fast<T>::store4(p, y); }
    static __device__ __forceinline__ void cvt_raw4(typename Fast<T>::raw4 r, float* x, const WinCvt&) { Fast<T>::cvt_raw4(r, x); }
};
__device__ __forceinline__ float win_one(uint32_t mant_bits, const WinCvt& c) {
    const float a = __fsub_rn(__uint_as_float(mant_bits), c.in_magic);
    const float q0 = __fmul_rn(a, c.r);
    return __fmaf_rn(__fmaf_rn(c.neg_rg, q0, a), c.r, q0);
}
__device__ __forceinline__ uint32_t win_quant(float y, const WinCvt& c) {   // low bits = rint(clamp(y) * rg) + lo
    return __float_as_uint(__fadd_rn(__fmul_rn(__saturatef(y), c.rg), 8388608.0f)) + (uint32_t)c.lo_out;
}
template <>
struct PixIO<uint16_t, true> {
    static __device__ __forceinline__ void cvt2(uint32_t w, float* x, const WinCvt& c) {
        x[0] = win_one(__byte_perm(w, 0x4B000000u, 0x7610), c);
        x[1] = win_one(__byte_perm(w, 0x4B000000u, 0x7632), c);
    }
    static __device__ __forceinline__ void load8(const uint16_t* p, float* x, const WinCvt& c) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p));
        cvt2(b.x, x, c); cvt2(b.y, x + 2, c); cvt2(b.z, x + 4, c); cvt2(b.w, x + 6, c);
    }
    static __device__ __forceinline__ void load4(const uint16_t* p, float* x, const WinCvt& c) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        cvt2(b.x, x, c); cvt2(b.y, x + 2, c);
    }
    static __device__ __forceinline__ void cvt_raw4(uint2 r, float* x, const WinCvt& c) { cvt2(r.x, x, c); cvt2(r.y, x + 2, c); }
    static __device__ __forceinline__ void store4(uint16_t* p, const float* y, const WinCvt& c) {
        uint2 o;
        o.x = __byte_perm(win_quant(y[0], c), win_quant(y[1], c), 0x5410);
        o.y = __byte_perm(win_quant(y[2], c), win_quant(y[3], c), 0x5410);
        *reinterpret_cast<uint2*>(p) = o;
    }
};
template <>
struct PixIO<int16_t, true> {   // v + 32768 by flipping the sign bit; the bias is part of in_magic
    static __device__ __forceinline__ void load8(const int16_t* p, float* x, const WinCvt& c) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p));
        PixIO<uint16_t, true>::cvt2(b.x ^ 0x80008000u, x, c); PixIO<uint16_t, true>::cvt2(b.y ^ 0x80008000u, x + 2, c);
        PixIO<uint16_t, true>::cvt2(b.z ^ 0x80008000u, x + 4, c); PixIO<uint16_t, true>::cvt2(b.w ^ 0x80008000u, x + 6, c);
    }
    static __device__ __forceinline__ void load4(const int16_t* p, float* x, const WinCvt& c) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        PixIO<uint16_t, true>::cvt2(b.x ^ 0x80008000u, x, c); PixIO<uint16_t, true>::cvt2(b.y ^ 0x80008000u, x + 2, c);
    }
    static __device__ __forceinline__ void cvt_raw4(uint2 r, float* x, const WinCvt& c) {
        PixIO<uint16_t, true>::cvt2(r.x ^ 0x80008000u, x, c); PixIO<uint16_t, true>::cvt2(r.y ^ 0x80008000u, x + 2, c);
    }
    static __device__ __forceinline__ void store4(int16_t* p, const float* y, const WinCvt& c) {
        PixIO<uint16_t, true>::store4(reinterpret_cast<uint16_t*>(p), y, c);   // two's complement low halves
    }
};
template <>
struct PixIO<uint8_t, true> {
    static __device__ __forceinline__ void cvt4(uint32_t w, float* x, const WinCvt& c) {
        x[0] = win_one(__byte_perm(w, 0x4B000000u, 0x7650), c);
        x[1] = win_one(__byte_perm(w, 0x4B000000u, 0x7651), c);
        x[2] = win_one(__byte_perm(w, 0x4B000000u, 0x7652), c);
        x[3] = win_one(__byte_perm(w, 0x4B000000u, 0x7653), c);
    }
    static __device__ __forceinline__ void load8(const uint8_t* p, float* x, const WinCvt& c) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        cvt4(b.x, x, c); cvt4(b.y, x + 4, c);
    }
    static __device__ __forceinline__ void load4(const uint8_t* p, float* x, const WinCvt& c) {
        cvt4(__ldg(reinterpret_cast<const uint32_t*>(p)), x, c);
    }
    static __device__ __forceinline__ void cvt_raw4(uint32_t r, float* x, const WinCvt& c) { cvt4(r, x, c); }
    static __device__ __forceinline__ void store4(uint8_t* p, const float* y, const WinCvt& c) {
        *reinterpret_cast<uint32_t*>(p) = pack_low_bytes(win_quant(y[0], c), win_quant(y[1], c), win_quant(y[2], c),
                                                         win_quant(y[3], c));
    }
};
template <>
struct PixIO<float, true> : PixIO<float, false> {};

// load_row24 (chain_fast.cuh) with the windowed conversion: 24 converted pixels x[0..23] = image columns
// c0-4 .. c0+19 of one row, the columns outside the image filled from the registers already loaded.
template <typename SrcT, bool WIN>
__device__ __forceinline__ void load_row24_win(const SrcT* row, int c0, int w, int border, float* x, const WinCvt& cv) {
    PixIO<SrcT, WIN>::load8(row + c0, x + 4, cv);
    PixIO<SrcT, WIN>::load8(row + c0 + 8, x + 12, cv);
    if (c0 != 0) {
        PixIO<SrcT, WIN>::load4(row + c0 - 4, x, cv);
    } else if (border == MIE_BORDER_REFLECT) {
        x[0] = x[8]; x[1] = x[7]; x[2] = x[6]; x[3] = x[5];
    } else {
        const float e = border == MIE_BORDER_REPLICATE ? x[4] : 0.0f;
        x[0] = e; x[1] = e; x[2] = e; x[3] = e;
    }
    if (c0 + 16 != w) {
        PixIO<SrcT, WIN>::load4(row + c0 + 16, x + 20, cv);
    } else if (border == MIE_BORDER_REFLECT) {
        x[20] = x[18]; x[21] = x[17]; x[22] = x[16]; x[23] = x[15];
    } else {
        const float e = border == MIE_BORDER_REPLICATE ? x[19] : 0.0f;
        x[20] = e; x[21] = e; x[22] = e; x[23] = e;
    }
}

// ---------------------------------------------------------------- integer rules for the default range
// For integer pixels with the dtype's default range, the kornia histogram bin floor(x01 * 256) and lookup
// index trunc(x01 * 255) (and equalize's floor(RN(RN(x01 * 255) / 255) * 256)) are pure integer functions
// of the code u = v - dtype_min:
//     16-bit:  bin = u >> 8,   index = u / 257 = (u * 65281) >> 24        8-bit:  bin = index = u
// (the float roundings never cross an integer boundary).  int_rules_ok() checks this against the float
// formulas for every code on the host before the first launch, so the standalone CLAHE / equalize kernels
// need no float conversion of their INPUT at all.
template <typename T> struct Codes;
template <> struct Codes<uint16_t> {
    static __device__ __forceinline__ void load8(const uint16_t* p, uint32_t* u) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p));
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
        u[4] = b.z & 0xFFFFu; u[5] = b.z >> 16; u[6] = b.w & 0xFFFFu; u[7] = b.w >> 16;
    }
    static __device__ __forceinline__ void load4(const uint16_t* p, uint32_t* u) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
    }
    static __device__ __forceinline__ uint32_t bin(uint32_t u) { return u >> 8; }
    static __device__ __forceinline__ uint32_t index(uint32_t u) { return (u * 65281u) >> 24; }
    // byte offset of the 8-byte cell-table entry: the index is the top byte of the product (IMAD, PRMT, shift-add)
    static __device__ __forceinline__ uint32_t entry_offset(uint32_t u) { return __byte_perm(u * 65281u, 0u, 0x4443) << 3; }
};
template <> struct Codes<int16_t> {
    static __device__ __forceinline__ void load8(const int16_t* p, uint32_t* u) {
        uint4 b = __ldg(reinterpret_cast<const uint4*>(p));
        b.x ^= 0x80008000u; b.y ^= 0x80008000u; b.z ^= 0x80008000u; b.w ^= 0x80008000u;   // v + 32768
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
        u[4] = b.z & 0xFFFFu; u[5] = b.z >> 16; u[6] = b.w & 0xFFFFu; u[7] = b.w >> 16;
    }
    static __device__ __forceinline__ void load4(const int16_t* p, uint32_t* u) {
        uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        b.x ^= 0x80008000u; b.y ^= 0x80008000u;
        u[0] = b.x & 0xFFFFu; u[1] = b.x >> 16; u[2] = b.y & 0xFFFFu; u[3] = b.y >> 16;
    }
    static __device__ __forceinline__ uint32_t bin(uint32_t u) { return u >> 8; }
    static __device__ __forceinline__ uint32_t index(uint32_t u) { return (u * 65281u) >> 24; }
    static __device__ __forceinline__ uint32_t entry_offset(uint32_t u) { return __byte_perm(u * 65281u, 0u, 0x4443) << 3; }
};
template <> struct Codes<uint8_t> {
    static __device__ __forceinline__ void load8(const uint8_t* p, uint32_t* u) {
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p));
        u[0] = b.x & 0xFFu; u[1] = (b.x >> 8) & 0xFFu; u[2] = (b.x >> 16) & 0xFFu; u[3] = b.x >> 24;
        u[4] = b.y & 0xFFu; u[5] = (b.y >> 8) & 0xFFu; u[6] = (b.y >> 16) & 0xFFu; u[7] = b.y >> 24;
    }
    static __device__ __forceinline__ void load4(const uint8_t* p, uint32_t* u) {
        const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(p));
        u[0] = b & 0xFFu; u[1] = (b >> 8) & 0xFFu; u[2] = (b >> 16) & 0xFFu; u[3] = b >> 24;
    }
    static __device__ __forceinline__ uint32_t bin(uint32_t u) { return u; }
    static __device__ __forceinline__ uint32_t index(uint32_t u) { return u; }
    static __device__ __forceinline__ uint32_t entry_offset(uint32_t u) { return u << 3; }
};
template <> struct Codes<float> {   // never used (IDX kernels are integer-only); keeps the dispatch macros compiling
    static __device__ __forceinline__ void load8(const float*, uint32_t* u) { for (int k = 0; k < 8; ++k) u[k] = 0; }
    static __device__ __forceinline__ void load4(const float*, uint32_t* u) { for (int k = 0; k < 4; ++k) u[k] = 0; }
    static __device__ __forceinline__ uint32_t bin(uint32_t) { return 0; }
    static __device__ __forceinline__ uint32_t index(uint32_t) { return 0; }
    static __device__ __forceinline__ uint32_t entry_offset(uint32_t) { return 0; }
};

// ---------------------------------------------------------------- host side
inline bool default_range_c(int dtype, float lo, float hi) {
    switch (dtype) {
        case MIE_U8: return lo == 0.0f && hi == 255.0f;
        case MIE_U16: return lo == 0.0f && hi == 65535.0f;
        case MIE_I16: return lo == -32768.0f && hi == 32767.0f;
        default: return true;
    }
}

// 0 = the dtype's default range (or float pixels), 1 = integer window usable by the windowed kernels
// (*cv filled in), -1 = neither.  The exhaustive check runs once per (dtype, lo, hi).
inline int range_mode(int dtype, float lo, float hi, WinCvt* cv) {
    if (default_range_c(dtype, lo, hi)) return 0;
    static const int kLo[3] = {0, 0, -32768}, kHi[3] = {255, 65535, 32767}, kBias[3] = {0, 0, 32768};
    if (dtype < MIE_U8 || dtype > MIE_I16) return -1;
    if (lo != std::floor(lo) || hi != std::floor(hi) || !(hi > lo)) return -1;
    if (lo < (float)kLo[dtype] || hi > (float)kHi[dtype]) return -1;   // outputs stay inside the dtype: no final clamp
    const float rg = hi - lo;
    WinCvt c;
    c.in_magic = 8388608.0f + (float)kBias[dtype] + lo;
    c.r = 1.0f / rg; c.neg_rg = -rg; c.rg = rg; c.lo_out = (int)lo;
    static std::mutex mu;
    static std::map<std::tuple<int, float, float>, bool> verdicts;
    bool ok;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto key = std::make_tuple(dtype, lo, hi);
        auto it = verdicts.find(key);
        if (it == verdicts.end()) {
            ok = true;
            for (int v = kLo[dtype]; v <= kHi[dtype] && ok; ++v) {
                volatile float exact = ((float)v - lo) / rg;                       // Px<T>::to01
                volatile float a = (8388608.0f + (float)(v + kBias[dtype])) - c.in_magic;
                volatile float q0 = a * c.r;
                volatile float rem = std::fmaf(c.neg_rg, q0, a);
                volatile float q = std::fmaf(rem, c.r, q0);
                float e = exact, f = q;
                ok = (e == f) && (std::signbit(e) == std::signbit(f) || e != 0.0f);
            }
            verdicts[key] = ok;
        } else {
            ok = it->second;
        }
    }
    if (!ok) return -1;
    *cv = c;
    return 1;
}

// Exhaustive host check of the integer rules above against the float formulas of the kernels / the oracle
// (IEEE float ops on the host are the ops the kernels reproduce).  Evaluated once.
inline bool int_rules_ok(int dtype) {
    static const bool ok16 = [] {
        for (int u = 0; u < 65536; ++u) {
            volatile float x = (float)u / 65535.0f;
            volatile float b = x * 256.0f, i = x * 255.0f;
            volatile float g = (float)i / 255.0f;
            volatile float eb = g * 256.0f;
            int bin = (int)std::floor((float)b); if (bin > 255) bin = 255;
            int ebin = (int)std::floor((float)eb); if (ebin > 255) ebin = 255;
            const int idx = (int)std::floor((float)i);
            if (bin != (u >> 8) || ebin != (u >> 8) || idx != (int)(((unsigned)u * 65281u) >> 24)) return false;
        }
        return true;
    }();
    static const bool ok8 = [] {
        for (int u = 0; u < 256; ++u) {
            volatile float x = (float)u / 255.0f;
            volatile float b = x * 256.0f, i = x * 255.0f;
            volatile float g = (float)i / 255.0f;
            volatile float eb = g * 256.0f;
            int bin = (int)std::floor((float)b); if (bin > 255) bin = 255;
            int ebin = (int)std::floor((float)eb); if (ebin > 255) ebin = 255;
            if (bin != u || ebin != u || (int)std::floor((float)i) != u) return false;
        }
        return true;
    }();
    return dtype == MIE_U8 ? ok8 : (dtype == MIE_U16 || dtype == MIE_I16) ? ok16 : false;
}

}  // namespace mie
