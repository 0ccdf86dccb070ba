/*
 * mie_oracle.c — CPU ORACLE (test infrastructure, NOT a product path).
 *
 * Plain-C restatement of the enhancement hot path with a fixed fp32 operation
 * order, used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs to check and to time against the CUDA
 * kernels.  Nothing in the shipped package imports or links this file.
 *
 * PARITY UNPINNED: the reference repository contains no implementation, tests or
 * golden vectors for this path (reference README.md and configs/__init__.py are
 * 0 bytes).  The algorithms live in third-party packages that are not on disk:
 *   kornia 0.8.2        (reference pyproject.toml:8,  uv.lock:219-230)
 *   scikit-image 0.26.0 (reference pyproject.toml:12, uv.lock:619-650)
 *   scipy 1.17.0        (uv.lock:653-681; scipy 1.18.1 is installed and readable)
 * Each function below names the upstream function it restates; the kornia /
 * skimage bodies are restated from their published algorithms (SURVEY.md §8(a),
 * Appendix A/B).  Pinning available in this image: cv2 4.13 (OpenCV CLAHE, bit
 * exact), scipy.ndimage (median, Gaussian), torchvision (global equalize) and a
 * torch-CPU kornia-style twin (oracle/kornia_twin.py) — see tests/test_oracle_*.py.
 *
 * Floating point: every multiply-add that the CUDA kernels fuse is an explicit
 * fmaf() here; compile with -ffp-contract=off so nothing else is fused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(MIE_ORACLE_NO_CLONES)
#define MIE_CLONES __attribute__((target_clones("arch=x86-64-v3", "default")))
#else
#define MIE_CLONES
#endif

enum { B_CONSTANT = 0, B_REFLECT = 1, B_REPLICATE = 2, B_CIRCULAR = 3, B_SYMMETRIC = 4 };
enum { DT_U8 = 0, DT_U16 = 1, DT_I16 = 2, DT_F32 = 3 };

/* Source index for coordinate i on an axis of length n; -1 = constant (zero). */
static inline int border_index(int i, int n, int mode) {
    if (i >= 0 && i < n) return i;
    if (mode == B_REFLECT) {
        if (n == 1) return 0;
        int p = 2 * (n - 1);
        int m = i % p;
        if (m < 0) m += p;
        return m < n ? m : p - m;
    }
    if (mode == B_REPLICATE) return i < 0 ? 0 : n - 1;
    if (mode == B_CIRCULAR) {
        int m = i % n;
        return m < 0 ? m + n : m;
    }
    if (mode == B_SYMMETRIC) { /* scipy.ndimage 'reflect': d c b a | a b c d | d c b a */
        int p = 2 * n;
        int m = i % p;
        if (m < 0) m += p;
        return m < n ? m : p - 1 - m;
    }
    return -1;
}

/* ------------------------------------------------------------------ pixel <-> [0,1]
 * x01 = (float(v) - lo) / (hi - lo);  back: rint(clamp(y,0,1)*(hi-lo)) + lo.
 * This is the normalisation a kornia user applies by hand to integer slices
 * (SURVEY.md §3 S1, §8(b) "Extensions").                                          */
MIE_CLONES
void orc_to01(const void* src, int dtype, int64_t count, float lo, float hi, float* out) {
    const float rg = hi - lo;
    int64_t i;
    switch (dtype) {
        case DT_U8: for (i = 0; i < count; ++i) out[i] = ((float)((const uint8_t*)src)[i] - lo) / rg; break;
        case DT_U16: for (i = 0; i < count; ++i) out[i] = ((float)((const uint16_t*)src)[i] - lo) / rg; break;
        case DT_I16: for (i = 0; i < count; ++i) out[i] = ((float)((const int16_t*)src)[i] - lo) / rg; break;
        default: memcpy(out, src, (size_t)count * 4); break;
    }
}

static inline float quant(float y, float lo, float rg, float tlo, float thi) {
    float c = fminf(fmaxf(y, 0.0f), 1.0f);
    float q = rintf(c * rg) + lo;
    return fminf(fmaxf(q, tlo), thi);
}

MIE_CLONES
void orc_from01(const float* in, int dtype, int64_t count, float lo, float hi, void* dst) {
    const float rg = hi - lo;
    int64_t i;
    switch (dtype) {
        case DT_U8: for (i = 0; i < count; ++i) ((uint8_t*)dst)[i] = (uint8_t)lrintf(quant(in[i], lo, rg, 0.f, 255.f)); break;
        case DT_U16: for (i = 0; i < count; ++i) ((uint16_t*)dst)[i] = (uint16_t)lrintf(quant(in[i], lo, rg, 0.f, 65535.f)); break;
        case DT_I16: for (i = 0; i < count; ++i) ((int16_t*)dst)[i] = (int16_t)lrintf(quant(in[i], lo, rg, -32768.f, 32767.f)); break;
        default: memcpy(dst, in, (size_t)count * 4); break;
    }
}

/* ------------------------------------------------------------------ Gaussian / unsharp
 * kornia.filters.gaussian_blur2d(separable=True) -> filter2d_separable: pad with
 * border_type, correlate rows with wx, then columns with wy (SURVEY.md §8(a) A3,
 * Appendix B2).  kornia.filters.unsharp_mask: x + (x - blur) (A4).
 * Accumulation: acc = w[0]*x[0]; acc = fmaf(w[t], x[t], acc), t = 1..K-1.          */
MIE_CLONES
static void sep_plane(const float* in, float* out, float* tmp, int h, int w, const float* wx, int kx,
                      const float* wy, int ky, int border, int unsharp, float amount, int clip) {
    const int rx = kx / 2, ry = ky / 2;
    for (int y = 0; y < h; ++y) {
        const float* row = in + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = 0.f;
            for (int t = 0; t < kx; ++t) {
                int sx = border_index(x - rx + t, w, border);
                float v = sx < 0 ? 0.0f : row[sx];
                acc = t == 0 ? wx[0] * v : fmaf(wx[t], v, acc);
            }
            tmp[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float acc = 0.f;
            for (int t = 0; t < ky; ++t) {
                int sy = border_index(y - ry + t, h, border);
                float v = sy < 0 ? 0.0f : tmp[(size_t)sy * w + x];
                acc = t == 0 ? wy[0] * v : fmaf(wy[t], v, acc);
            }
            if (unsharp) {
                float c = in[(size_t)y * w + x];
                acc = fmaf(amount, c - acc, c); /* amount == 1: c + (c - acc), kornia's unsharp_mask */
                if (clip) acc = fminf(fmaxf(acc, 0.0f), 1.0f);
            }
            out[(size_t)y * w + x] = acc;
        }
    }
}

/* skimage.filters.unsharp_mask(image, radius, amount): x + amount (x - blur), optionally clipped to [0,1]. */
int orc_gaussian2d_ex(const float* in, float* out, int64_t n, int h, int w, const float* wx, int kx, const float* wy,
                      int ky, int border, int unsharp, float amount, int clip);

int orc_gaussian2d(const float* in, float* out, int64_t n, int h, int w, const float* wx, int kx, const float* wy,
                   int ky, int border, int unsharp) {
    return orc_gaussian2d_ex(in, out, n, h, w, wx, kx, wy, ky, border, unsharp, 1.0f, 0);
}

int orc_gaussian2d_ex(const float* in, float* out, int64_t n, int h, int w, const float* wx, int kx, const float* wy,
                      int ky, int border, int unsharp, float amount, int clip) {
    int err = 0;
#pragma omp parallel
    {
        float* tmp = (float*)malloc((size_t)h * w * sizeof(float));
        if (!tmp) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t i = 0; i < n; ++i)
                sep_plane(in + (size_t)i * h * w, out + (size_t)i * h * w, tmp, h, w, wx, kx, wy, ky, border, unsharp,
                          amount, clip);
            free(tmp);
        }
    }
    return err;
}

/* ------------------------------------------------------------------ CLAHE, kornia semantics
 * kornia.enhance.equalize_clahe (enhance/equalization.py: _compute_tiles,
 * _compute_luts, _compute_interpolation_tiles, _compute_equalized_tiles);
 * SURVEY.md §8(a) A1, Appendix B1.                                                 */
typedef struct {
    int h, w, gh, gw, th, tw, hp, wp;
} geom_t;

/* returns 0 ok, -5 bad grid, -6 padding exceeds the image (kornia ValueError /
 * torch reflect-pad RuntimeError) */
static int kornia_geom(int h, int w, int gh, int gw, geom_t* g) {
    if (gh <= 0 || gw <= 0) return -5;
    g->h = h; g->w = w; g->gh = gh; g->gw = gw;
    g->th = (h + gh - 1) / gh; g->tw = (w + gw - 1) / gw;
    if (g->th & 1) g->th += 1;
    if (g->tw & 1) g->tw += 1;
    g->hp = g->th * gh; g->wp = g->tw * gw;
    if (g->hp - h >= h || g->wp - w >= w) return -6;
    return 0;
}

/* torch.histc(tile, bins=256, min=0, max=1): out-of-range and NaN ignored. */
static inline int kornia_bin(float v) {
    if (!(v >= 0.0f && v <= 1.0f)) return -1;
    int b = (int)(v * 256.0f);
    return b > 255 ? 255 : b;
}

int orc_clahe_hist_kornia(const float* in, int64_t n, int h, int w, int gh, int gw, uint32_t* hist) {
    geom_t g;
    int rc = kornia_geom(h, w, gh, gw, &g);
    if (rc) return rc;
    memset(hist, 0, (size_t)n * gh * gw * 256 * sizeof(uint32_t));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const float* img = in + (size_t)i * h * w;
        for (int y = 0; y < g.hp; ++y) {
            int sy = border_index(y, h, B_REFLECT);
            int ty = y / g.th;
            for (int x = 0; x < g.wp; ++x) {
                int sx = border_index(x, w, B_REFLECT);
                int b = kornia_bin(img[(size_t)sy * w + sx]);
                if (b >= 0) hist[(((size_t)i * gh + ty) * gw + x / g.tw) * 256 + b] += 1;
            }
        }
    }
    return 0;
}

/* _compute_luts: clamp to max_val, redistribute the excess evenly, residual to
 * the first bins, cumsum * fp32(255/pixels), clamp, floor. */
int orc_clahe_luts_from_hist_kornia(const uint32_t* hist, int64_t tiles, int th, int tw, double clip_limit,
                                    uint8_t* luts) {
    const int pixels = th * tw;
    int max_val = 0;
    if (clip_limit > 0.0) {
        double q = floor(clip_limit * (double)pixels / 256.0); /* python: clip * pixels // 256 */
        if (q < 1.0) q = 1.0;
        max_val = q > 2147483647.0 ? 2147483647 : (int)q;
    }
    const float lut_scale = (float)(255.0 / (double)pixels);
    for (int64_t t = 0; t < tiles; ++t) {
        int hv[256];
        for (int b = 0; b < 256; ++b) hv[b] = (int)hist[t * 256 + b];
        if (max_val > 0) {
            int sum = 0;
            for (int b = 0; b < 256; ++b) {
                if (hv[b] > max_val) hv[b] = max_val;
                sum += hv[b];
            }
            int clipped = pixels - sum;
            int resid = clipped % 256;
            int redist = (clipped - resid) / 256;
            for (int b = 0; b < 256; ++b) hv[b] += redist + (b < resid ? 1 : 0);
        }
        int cum = 0;
        for (int b = 0; b < 256; ++b) {
            cum += hv[b];
            float f = (float)cum * lut_scale;
            f = floorf(fminf(fmaxf(f, 0.0f), 255.0f));
            luts[t * 256 + b] = (uint8_t)(int)f;
        }
    }
    return 0;
}

static inline void kornia_axis(int y, int T, int G, int* j0, int* j1, float* wgt) {
    int hh = T / 2;
    if (y < hh) {
        *j0 = *j1 = 0; *wgt = 0.0f;
    } else if (y >= T * G - hh) {
        *j0 = *j1 = G - 1; *wgt = 0.0f;
    } else {
        int rel = y - hh;
        *j0 = rel / T;
        int r = rel - *j0 * T;
        *j1 = *j0 + 1;
        *wgt = (float)(T - 1 - r) / (float)(T - 1);
    }
}

static inline int kornia_idx(float v) {
    float f = v * 255.0f;
    f = fminf(fmaxf(f, 0.0f), 255.0f);
    return (int)f; /* truncation, like (x*255).long() */
}

/* _compute_equalized_tiles: gather the <=4 neighbouring LUTs at (x*255).long()
 * and blend: t = tr + wx*(tl-tr); b = br + wx*(bl-br); out = b + wy*(t-b); /255. */
MIE_CLONES
int orc_clahe_apply_kornia(const float* in, float* out, int64_t n, int h, int w, int gh, int gw,
                           const uint8_t* luts) {
    geom_t g;
    int rc = kornia_geom(h, w, gh, gw, &g);
    if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const float* img = in + (size_t)i * h * w;
        float* o = out + (size_t)i * h * w;
        const uint8_t* nl = luts + (size_t)i * gh * gw * 256;
        for (int y = 0; y < h; ++y) {
            int j0, j1; float wy;
            kornia_axis(y, g.th, gh, &j0, &j1, &wy);
            for (int x = 0; x < w; ++x) {
                int i0, i1; float wx;
                kornia_axis(x, g.tw, gw, &i0, &i1, &wx);
                int idx = kornia_idx(img[(size_t)y * w + x]);
                float tl = (float)nl[((size_t)j0 * gw + i0) * 256 + idx], tr = (float)nl[((size_t)j0 * gw + i1) * 256 + idx];
                float bl = (float)nl[((size_t)j1 * gw + i0) * 256 + idx], br = (float)nl[((size_t)j1 * gw + i1) * 256 + idx];
                float t = fmaf(wx, tl - tr, tr);
                float b = fmaf(wx, bl - br, br);
                float r = fmaf(wy, t - b, b);
                o[(size_t)y * w + x] = r / 255.0f;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ CLAHE, OpenCV semantics (uint8)
 * cv::CLAHE::apply — SURVEY.md §8(a) A1', Appendix A (bit-exact against cv2 4.13). */
static void opencv_geom(int h, int w, int gh, int gw, geom_t* g) {
    g->h = h; g->w = w; g->gh = gh; g->gw = gw;
    /* cv::CLAHE pads BOTH axes by tiles - dim % tiles as soon as EITHER is not divisible, so an axis that
     * divides evenly then still grows by a full `tiles` pixels (tile size dim/tiles + 1); checked against
     * cv2 4.13 on 512x500 and 300x512 (tests/golden/cv2_clahe.json). */
    if (h % gh == 0 && w % gw == 0) { g->th = h / gh; g->tw = w / gw; }
    else { g->th = (h + gh - h % gh) / gh; g->tw = (w + gw - w % gw) / gw; }
    g->hp = g->th * gh; g->wp = g->tw * gw;
}

int orc_clahe_hist_opencv_u8(const uint8_t* in, int64_t n, int h, int w, int gh, int gw, uint32_t* hist) {
    if (gh <= 0 || gw <= 0) return -5;
    geom_t g;
    opencv_geom(h, w, gh, gw, &g);
    memset(hist, 0, (size_t)n * gh * gw * 256 * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* img = in + (size_t)i * h * w;
        for (int y = 0; y < g.hp; ++y) {
            int sy = border_index(y, h, B_REFLECT);
            for (int x = 0; x < g.wp; ++x) {
                int sx = border_index(x, w, B_REFLECT);
                hist[(((size_t)i * gh + y / g.th) * gw + x / g.tw) * 256 + img[(size_t)sy * w + sx]] += 1;
            }
        }
    }
    return 0;
}

int orc_clahe_luts_from_hist_opencv(const uint32_t* hist, int64_t tiles, int th, int tw, double clip_limit,
                                    uint8_t* luts) {
    const int area = th * tw;
    int clip = 0;
    if (clip_limit > 0.0) {
        double q = clip_limit * (double)area / 256.0;
        clip = q > 2147483647.0 ? 2147483647 : (int)q;
        if (clip < 1) clip = 1;
    }
    const float lut_scale = 255.0f / (float)area;
    for (int64_t t = 0; t < tiles; ++t) {
        int hv[256];
        for (int b = 0; b < 256; ++b) hv[b] = (int)hist[t * 256 + b];
        if (clip > 0) {
            int clipped = 0;
            for (int b = 0; b < 256; ++b)
                if (hv[b] > clip) { clipped += hv[b] - clip; hv[b] = clip; }
            int rb = clipped / 256, res = clipped - rb * 256;
            for (int b = 0; b < 256; ++b) hv[b] += rb;
            if (res != 0) {
                int step = 256 / res;
                if (step < 1) step = 1;
                for (int b = 0; b < 256 && res > 0; b += step, --res) hv[b] += 1;
            }
        }
        int cum = 0;
        for (int b = 0; b < 256; ++b) {
            cum += hv[b];
            float f = rintf((float)cum * lut_scale);
            f = fminf(fmaxf(f, 0.0f), 255.0f);
            luts[t * 256 + b] = (uint8_t)(int)f;
        }
    }
    return 0;
}

int orc_clahe_apply_opencv_u8(const uint8_t* in, uint8_t* out, int64_t n, int h, int w, int gh, int gw,
                              const uint8_t* luts) {
    if (gh <= 0 || gw <= 0) return -5;
    geom_t g;
    opencv_geom(h, w, gh, gw, &g);
    const float inv_th = 1.0f / (float)g.th, inv_tw = 1.0f / (float)g.tw;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t* img = in + (size_t)i * h * w;
        uint8_t* o = out + (size_t)i * h * w;
        const uint8_t* nl = luts + (size_t)i * gh * gw * 256;
        for (int y = 0; y < h; ++y) {
            float tyf = (float)y * inv_th - 0.5f;
            int ty1 = (int)floorf(tyf);
            float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
            int ty2 = ty1 + 1;
            if (ty1 < 0) ty1 = 0;
            if (ty2 > gh - 1) ty2 = gh - 1;
            for (int x = 0; x < w; ++x) {
                float txf = (float)x * inv_tw - 0.5f;
                int tx1 = (int)floorf(txf);
                float xa = txf - (float)tx1, xa1 = 1.0f - xa;
                int tx2 = tx1 + 1;
                if (tx1 < 0) tx1 = 0;
                if (tx2 > gw - 1) tx2 = gw - 1;
                int v = img[(size_t)y * w + x];
                float l11 = nl[((size_t)ty1 * gw + tx1) * 256 + v], l12 = nl[((size_t)ty1 * gw + tx2) * 256 + v];
                float l21 = nl[((size_t)ty2 * gw + tx1) * 256 + v], l22 = nl[((size_t)ty2 * gw + tx2) * 256 + v];
                float top = l11 * xa1 + l12 * xa;
                float bot = l21 * xa1 + l22 * xa;
                float res = rintf(top * ya1 + bot * ya);
                res = fminf(fmaxf(res, 0.0f), 255.0f);
                o[(size_t)y * w + x] = (uint8_t)(int)res;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ CLAHE, OpenCV semantics (uint16, 65 536 bins)
 * cv::CLAHE::apply on CV_16UC1 — SURVEY.md §8(a) A1', Appendix A (bit-exact against cv2 4.13,
 * tests/test_oracle.py).  Same algorithm as the uint8 mode with bins = 65536:
 * lutScale = float(65535) / float(area); clip = max(int(clipLimit * area / 65536), 1); the residual of
 * the clipped mass goes to bins 0, step, 2 step, ... (step = max(65536 / residual, 1)).
 * luts_out (optional) receives n * gh * gw * 65536 uint16 entries. */
int orc_clahe_opencv_u16(const uint16_t* in, uint16_t* out, int64_t n, int h, int w, int gh, int gw,
                         double clip_limit, uint16_t* luts_out) {
    if (gh <= 0 || gw <= 0) return -5;
    geom_t g;
    opencv_geom(h, w, gh, gw, &g);
    const int bins = 65536;
    const int area = g.th * g.tw;
    int clip = 0;
    if (clip_limit > 0.0) {
        double q = clip_limit * (double)area / (double)bins;
        clip = q > 2147483647.0 ? 2147483647 : (int)q;
        if (clip < 1) clip = 1;
    }
    const float lut_scale = (float)(bins - 1) / (float)area;
    const float inv_th = 1.0f / (float)g.th, inv_tw = 1.0f / (float)g.tw;
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const uint16_t* img = in + (size_t)i * h * w;
        uint16_t* o = out ? out + (size_t)i * h * w : NULL;
        uint16_t* luts = luts_out ? luts_out + (size_t)i * gh * gw * bins
                                  : (uint16_t*)malloc((size_t)gh * gw * bins * sizeof(uint16_t));
        int* hv = (int*)malloc((size_t)bins * sizeof(int));
        if (!luts || !hv) { rc = -20; free(hv); if (!luts_out) free(luts); continue; }
        for (int ty = 0; ty < gh; ++ty)
            for (int tx = 0; tx < gw; ++tx) {
                memset(hv, 0, (size_t)bins * sizeof(int));
                for (int y = ty * g.th; y < (ty + 1) * g.th; ++y) {
                    int sy = border_index(y, h, B_REFLECT);
                    for (int x = tx * g.tw; x < (tx + 1) * g.tw; ++x)
                        hv[img[(size_t)sy * w + border_index(x, w, B_REFLECT)]] += 1;
                }
                if (clip > 0) {
                    int clipped = 0;
                    for (int b = 0; b < bins; ++b)
                        if (hv[b] > clip) { clipped += hv[b] - clip; hv[b] = clip; }
                    int rb = clipped / bins, res = clipped - rb * bins;
                    for (int b = 0; b < bins; ++b) hv[b] += rb;
                    if (res != 0) {
                        int step = bins / res;
                        if (step < 1) step = 1;
                        for (int b = 0; b < bins && res > 0; b += step, --res) hv[b] += 1;
                    }
                }
                uint16_t* lut = luts + ((size_t)ty * gw + tx) * bins;
                int cum = 0;
                for (int b = 0; b < bins; ++b) {
                    cum += hv[b];
                    float f = rintf((float)cum * lut_scale);
                    f = fminf(fmaxf(f, 0.0f), 65535.0f);
                    lut[b] = (uint16_t)(int)f;
                }
            }
        if (o) {
            for (int y = 0; y < h; ++y) {
                float tyf = (float)y * inv_th - 0.5f;
                int ty1 = (int)floorf(tyf);
                float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
                int ty2 = ty1 + 1;
                if (ty1 < 0) ty1 = 0;
                if (ty2 > gh - 1) ty2 = gh - 1;
                for (int x = 0; x < w; ++x) {
                    float txf = (float)x * inv_tw - 0.5f;
                    int tx1 = (int)floorf(txf);
                    float xa = txf - (float)tx1, xa1 = 1.0f - xa;
                    int tx2 = tx1 + 1;
                    if (tx1 < 0) tx1 = 0;
                    if (tx2 > gw - 1) tx2 = gw - 1;
                    int v = img[(size_t)y * w + x];
                    float l11 = luts[((size_t)ty1 * gw + tx1) * bins + v], l12 = luts[((size_t)ty1 * gw + tx2) * bins + v];
                    float l21 = luts[((size_t)ty2 * gw + tx1) * bins + v], l22 = luts[((size_t)ty2 * gw + tx2) * bins + v];
                    float top = l11 * xa1 + l12 * xa;
                    float bot = l21 * xa1 + l22 * xa;
                    float res = rintf(top * ya1 + bot * ya);
                    res = fminf(fmaxf(res, 0.0f), 65535.0f);
                    o[(size_t)y * w + x] = (uint16_t)(int)res;
                }
            }
        }
        free(hv);
        if (!luts_out) free(luts);
    }
    return rc;
}

/* ------------------------------------------------------------------ non-local means (skimage fast mode)
 * skimage.restoration.denoise_nl_means(image, patch_size, patch_distance, h, fast_mode=True, sigma)
 * — scikit-image 0.26.0 (reference pyproject.toml:12); the Cython source is not on disk, the loop below
 * restates _fast_nl_means_denoising_2d [RECALLED] in float64, as skimage computes for
 * integer / float64 images:  pad 'reflect' by o + d + 1;  for every shift t in [-d, d]^2 the patch
 * distance is the integral-image box difference over rows / columns p-o+1 .. p+o of
 * (I(q) - I(q + t))^2 - 2 sigma^2, clamped at 0 and divided by h^2 s^2;  weights exp(-dist) (dist <= 5),
 * the zero shift counted twice (upstream accumulates it into both endpoints, which coincide).
 * in: float64 planes in [0,1] (already normalised); out: float64.  tests/test_oracle.py checks this
 * closed form against a literal transcription of the upstream accumulation loops on small images. */
int orc_nlm_fast(const double* in, double* out, int64_t n, int h, int w, int patch_size, int patch_distance,
                 double hpar, double sigma) {
    int s = patch_size + (patch_size % 2 == 0 ? 1 : 0);
    const int o = s / 2, d = patch_distance;
    if (o < 1 || d < 0 || !(hpar > 0.0) || o + d + 1 >= h || o + d + 1 >= w) return -7;
    const double var2 = 2.0 * sigma * sigma;
    const double h2s2 = hpar * hpar * (double)s * (double)s;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int64_t i = 0; i < n; ++i) {
        for (int y = 0; y < h; ++y) {
            const double* img = in + (size_t)i * h * w;
            for (int x = 0; x < w; ++x) {
                double num = 0.0, den = 0.0;
                for (int ty = -d; ty <= d; ++ty)
                    for (int tx = -d; tx <= d; ++tx) {
                        double D = 0.0;
                        for (int qy = -o + 1; qy <= o; ++qy) {
                            const int ay = border_index(y + qy, h, B_REFLECT), by = border_index(y + qy + ty, h, B_REFLECT);
                            for (int qx = -o + 1; qx <= o; ++qx) {
                                const int ax = border_index(x + qx, w, B_REFLECT), bx = border_index(x + qx + tx, w, B_REFLECT);
                                const double df = img[(size_t)ay * w + ax] - img[(size_t)by * w + bx];
                                D += df * df - var2;
                            }
                        }
                        const double dist = (D > 0.0 ? D : 0.0) / h2s2;
                        if (dist > 5.0) continue;
                        const double wgt = ((ty == 0 && tx == 0) ? 2.0 : 1.0) * exp(-dist);
                        num += wgt * img[(size_t)border_index(y + ty, h, B_REFLECT) * w + border_index(x + tx, w, B_REFLECT)];
                        den += wgt;
                    }
                out[(size_t)i * h * w + (size_t)y * w + x] = num / den;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ non-local means (skimage slow mode)
 * skimage.restoration.denoise_nl_means(image, patch_size, patch_distance, h, fast_mode=False, sigma)
 * — scikit-image 0.26.0 (reference pyproject.toml:12); restates _nl_means_denoising_2d and its helper
 * patch_distance_2d [RECALLED] in float64:
 *   s = patch_size (+1 if even), o = s / 2;  image reflect-padded by o only;
 *   patch weights  w[i][j] = exp(-(di^2 + dj^2) / (2 A^2)),  A = (s - 1) / 4,  then  w *= 1 / (sum(w) h^2);
 *   the search window of pixel (row, col) holds the centres i in [row - min(d, row), row + min(d + 1, H - row)),
 *   j likewise: it is CLIPPED at the image, not padded (unlike fast mode);
 *   distance = sum over patch rows, then columns, of w * ((p1 - p2)^2 - 2 sigma^2); before every patch ROW the
 *   running distance is tested against the cut-off 5 and the weight is 0 if it is exceeded;
 *   weight = exp(-max(0, distance));  out = sum(weight * centre pixel) / sum(weight).
 * `wgt` (s*s doubles, row-major) is filled by orc_nlm_patch_weights so that the CUDA host code and this function
 * can be fed the very same numbers.                                                                          */
int orc_nlm_patch_weights(int patch_size, double hpar, double* wgt) {
    int s = patch_size + (patch_size % 2 == 0 ? 1 : 0);
    const int o = s / 2;
    const double A = ((double)s - 1.0) / 4.0;
    double sum = 0.0;
    for (int i = 0; i < s; ++i)
        for (int j = 0; j < s; ++j) {
            const double di = (double)(i - o), dj = (double)(j - o);
            wgt[i * s + j] = exp(-(di * di + dj * dj) / (2.0 * A * A));
            sum += wgt[i * s + j];
        }
    const double scale = 1.0 / (sum * hpar * hpar);
    for (int i = 0; i < s * s; ++i) wgt[i] *= scale;
    return s;
}

int orc_nlm_slow(const double* in, double* out, int64_t n, int h, int w, int patch_size, int patch_distance,
                 double hpar, double sigma) {
    int s = patch_size + (patch_size % 2 == 0 ? 1 : 0);
    const int o = s / 2, d = patch_distance;
    if (o < 1 || s > 33 || d < 0 || !(hpar > 0.0) || o >= h || o >= w) return -7;
    double wgt[33 * 33];
    orc_nlm_patch_weights(patch_size, hpar, wgt);
    const double var2 = 2.0 * sigma * sigma;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int64_t i = 0; i < n; ++i) {
        for (int row = 0; row < h; ++row) {
            const double* img = in + (size_t)i * h * w;
            const int i0 = row - (d < row ? d : row), i1 = row + (d + 1 < h - row ? d + 1 : h - row);
            for (int col = 0; col < w; ++col) {
                const int j0 = col - (d < col ? d : col), j1 = col + (d + 1 < w - col ? d + 1 : w - col);
                double num = 0.0, den = 0.0;
                for (int ci = i0; ci < i1; ++ci)
                    for (int cj = j0; cj < j1; ++cj) {
                        double dist = 0.0, weight = -1.0;
                        for (int pi = 0; pi < s; ++pi) {
                            if (dist > 5.0) { weight = 0.0; break; }
                            const int ay = border_index(row - o + pi, h, B_REFLECT), by = border_index(ci - o + pi, h, B_REFLECT);
                            for (int pj = 0; pj < s; ++pj) {
                                const int ax = border_index(col - o + pj, w, B_REFLECT), bx = border_index(cj - o + pj, w, B_REFLECT);
                                const double df = img[(size_t)ay * w + ax] - img[(size_t)by * w + bx];
                                dist += wgt[pi * s + pj] * (df * df - var2);
                            }
                        }
                        if (weight < 0.0) weight = exp(-(dist > 0.0 ? dist : 0.0));
                        den += weight;
                        num += weight * img[(size_t)ci * w + cj];
                    }
                out[(size_t)i * h * w + (size_t)row * w + col] = num / den;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ median (pure selection)
 * 2-D: kornia.filters.median_blur (zero padding; torch.median = lower median, the
 * true median for odd windows) / skimage.filters.median 2-D ('nearest');
 * SURVEY.md §8(a) A5.  3-D: skimage.filters.median -> scipy.ndimage.median_filter
 * 3x3x3, rank 27//2 (site-packages/scipy/ndimage/_filters.py:1928-2028); A6.
 * Samples are carried as double so one routine serves every dtype exactly.       */
static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

int orc_median2d(const double* in, double* out, int64_t n, int h, int w, int ky, int kx, int border) {
    const int ry = ky / 2, rx = kx / 2, cnt = ky * kx;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const double* img = in + (size_t)i * h * w;
        double* o = out + (size_t)i * h * w;
        double win[33 * 33];
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                int k = 0;
                for (int dy = -ry; dy <= ry; ++dy) {
                    int sy = border_index(y + dy, h, border);
                    for (int dx = -rx; dx <= rx; ++dx) {
                        int sx = border_index(x + dx, w, border);
                        win[k++] = (sy < 0 || sx < 0) ? 0.0 : img[(size_t)sy * w + sx];
                    }
                }
                qsort(win, cnt, sizeof(double), cmp_double);
                o[(size_t)y * w + x] = win[cnt / 2];
            }
    }
    return 0;
}

/* halo_lo / halo_hi: optional h*w planes standing in for z = -1 and z = d. */
int orc_median3d(const double* in, double* out, int d, int h, int w, const double* halo_lo, const double* halo_hi,
                 int border) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int z = 0; z < d; ++z) {
        double win[27];
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                int k = 0;
                for (int dz = -1; dz <= 1; ++dz) {
                    int zz = z + dz;
                    const double* plane;
                    if (zz < 0) plane = halo_lo ? halo_lo : (border == B_REPLICATE ? in : NULL);
                    else if (zz >= d) plane = halo_hi ? halo_hi : (border == B_REPLICATE ? in + (size_t)(d - 1) * h * w : NULL);
                    else plane = in + (size_t)zz * h * w;
                    for (int dy = -1; dy <= 1; ++dy) {
                        int sy = border_index(y + dy, h, border);
                        for (int dx = -1; dx <= 1; ++dx) {
                            int sx = border_index(x + dx, w, border);
                            win[k++] = (!plane || sy < 0 || sx < 0) ? 0.0 : plane[(size_t)sy * w + sx];
                        }
                    }
                }
                qsort(win, 27, sizeof(double), cmp_double);
                out[((size_t)z * h + y) * w + x] = win[13];
            }
    }
    return 0;
}

/* ------------------------------------------------------------------ bilateral
 * kornia.filters.bilateral_blur, single channel:
 * w = space[dy,dx] * exp(-0.5/sigma_color^2 * (v - c)^2); out = sum(w v)/sum(w).
 * The colour weight is evaluated as 2^t, t = c2 * ((v - c) * (v - c)), c2 = fp32(-0.5 log2(e) / sigma_color^2),
 * by mie_exp2n (below) — a fixed sequence of fp32 operations (max rel err 1.7e-7) that the CUDA kernel
 * reproduces bit for bit.  mie_exp (Cody-Waite on ln 2, degree 6) is the general-purpose variant kept for
 * reference; the bilateral uses the base-2 form because it needs 3 fewer FMA-pipe operations per tap.   */
static inline float mie_exp2n(float t) { /* 2^t for t <= 0 (NaN and t < -125 give 2^-125) */
    t = fmaxf(t, -125.0f);
    float u = t + 12582912.0f;          /* 1.5 * 2^23: the low mantissa bits of u hold n = rint(t) */
    float n = u - 12582912.0f;
    float f = t - n;                    /* exact, |f| <= 0.5 */
    float p = 0.0013264685403555632f;   /* degree-5 fit of 2^f on [-0.5, 0.5], constant term exactly 1 */
    p = fmaf(p, f, 0.009671504609286785f);
    p = fmaf(p, f, 0.05550733953714371f);
    p = fmaf(p, f, 0.24022242426872253f);
    p = fmaf(p, f, 0.6931470036506653f);
    p = fmaf(p, f, 1.0f);
    union { uint32_t i; float f; } a, b, r;
    a.f = p; b.f = u;
    r.i = a.i + (b.i << 23);            /* p * 2^n: p in [0.70, 1.42], n >= -125, never subnormal */
    return r.f;
}

void orc_exp2n(const float* in, float* out, int64_t count) {
    for (int64_t i = 0; i < count; ++i) out[i] = mie_exp2n(in[i]);
}

static inline float mie_exp(float a) {
    a = fminf(fmaxf(a, -87.0f), 88.0f);
    float n = rintf(a * 1.44269504088896341f);
    float r = fmaf(n, -0.693145751953125f, a);      /* ln2 high part (exact in fp32) */
    r = fmaf(n, -1.42860682030941723e-6f, r);       /* ln2 low part */
    float p = 1.3888889225e-3f;                     /* 1/720 */
    p = fmaf(p, r, 8.3333337680e-3f);               /* 1/120 */
    p = fmaf(p, r, 4.1666667908e-2f);               /* 1/24 */
    p = fmaf(p, r, 1.6666667163e-1f);               /* 1/6 */
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    union { int32_t i; float f; } s;
    s.i = ((int32_t)n + 127) << 23;
    return p * s.f;
}

void orc_exp(const float* in, float* out, int64_t count) {
    for (int64_t i = 0; i < count; ++i) out[i] = mie_exp(in[i]);
}

MIE_CLONES
int orc_bilateral(const float* in, float* out, int64_t n, int h, int w, const float* wspace, int ky, int kx,
                  float sigma_color, int border) {
    const int ry = ky / 2, rx = kx / 2;
    const float coef = (float)(-0.5 * 1.4426950408889634 / ((double)sigma_color * (double)sigma_color));
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const float* img = in + (size_t)i * h * w;
        float* o = out + (size_t)i * h * w;
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const float c = img[(size_t)y * w + x];
                float num = 0.0f, den = 0.0f;
                for (int dy = 0; dy < ky; ++dy) {
                    int sy = border_index(y - ry + dy, h, border);
                    for (int dx = 0; dx < kx; ++dx) {
                        int sx = border_index(x - rx + dx, w, border);
                        float v = (sy < 0 || sx < 0) ? 0.0f : img[(size_t)sy * w + sx];
                        float dv = v - c;
                        float wgt = wspace[dy * kx + dx] * mie_exp2n(coef * (dv * dv));
                        num = fmaf(wgt, v, num);
                        den = den + wgt;
                    }
                }
                o[(size_t)y * w + x] = num / den;
            }
    }
    return 0;
}

/* ------------------------------------------------------------------ global equalisation
 * kornia.enhance.equalize -> _scale_channel, the
 * torchvision rule (site-packages/torchvision/transforms/_functional_tensor.py:863-881):
 * v = x*255; hist = histc(v, 256, 0, 255); step = (sum(nz) - nz[-1]) // 255;
 * lut = [0, ((cumsum + step//2) // step)[:-1]] clamped; out = lut[trunc(v)] / 255,
 * or v / 255 unchanged when step == 0.                                            */
int orc_equalize(const float* in, float* out, int64_t n, int h, int w) {
    const size_t px = (size_t)h * w;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        const float* img = in + i * px;
        float* o = out + i * px;
        int64_t hist[256];
        memset(hist, 0, sizeof(hist));
        for (size_t p = 0; p < px; ++p) {
            float v = img[p] * 255.0f;
            if (!(v >= 0.0f && v <= 255.0f)) continue; /* histc ignores out-of-range and NaN */
            int b = (int)((v / 255.0f) * 256.0f);
            if (b > 255) b = 255;
            hist[b] += 1;
        }
        int64_t total = 0, last = 0;
        for (int b = 0; b < 256; ++b) {
            total += hist[b];
            if (hist[b]) last = hist[b];
        }
        int64_t step = (total - last) / 255;
        float lut[256];
        if (step > 0) {
            int64_t cum = 0;
            lut[0] = 0.0f;
            for (int b = 0; b < 255; ++b) {
                cum += hist[b];
                int64_t q = (cum + step / 2) / step;
                lut[b + 1] = (float)(q > 255 ? 255 : q);
            }
        }
        for (size_t p = 0; p < px; ++p) {
            float v = img[p] * 255.0f;
            float r;
            if (step > 0) {
                float c = fminf(fmaxf(v, 0.0f), 255.0f);
                r = lut[(int)c];
            } else {
                r = v;
            }
            o[p] = r / 255.0f;
        }
    }
    return 0;
}

/* ====================================================================== scikit-image exposure / restoration
 * RECALLED restatements (scikit-image 0.26.0, reference pyproject.toml:12 — the package is
 * not on disk: parity unpinned).  Per-pixel formulation, structurally independent of the array-level numpy twin in
 * oracle/skimage_twin.py; the two must agree bit for bit (tests/test_skimage_exposure.py).  The CUDA kernels
 * (csrc/sk_exposure.cu, csrc/sk_bilateral.cu) follow THIS formulation.  All float64 arithmetic is plain IEEE in the
 * order written (-ffp-contract=off). */

/* img_as_float of one code (skimage.util.dtype): unsigned v / max; signed (2 v + 1) / (max - min). */
static inline double sk_as_float(int dtype, int v) {
    if (dtype == DT_U8) return (double)v / 255.0;
    if (dtype == DT_U16) return (double)v / 65535.0;
    return ((double)v * 2.0 + 1.0) / 65535.0; /* int16: [-1, 1] */
}
static inline int sk_code(const void* img, int dtype, size_t p) {
    if (dtype == DT_U8) return ((const uint8_t*)img)[p];
    if (dtype == DT_U16) return ((const uint16_t*)img)[p];
    return ((const int16_t*)img)[p];
}

/* numpy.pad(mode='reflect') index (no edge repeat; repeated reflection for pads longer than the axis). */
static inline int sk_reflect(int i, int n) {
    if (n == 1) return 0;
    int p = 2 * (n - 1);
    int m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - m;
}

/* Stage 1 of equalize_adapthist: img_as_float -> rescale_intensity(out_range=(0, 2^14 - 1)) -> np.round -> uint16.
 * One image (h x w) of an integer dtype or float32 (float32 images are processed in float32, like upstream). */
int orc_sk_adapthist_grey(const void* img, int dtype, int h, int w, uint16_t* grey) {
    const size_t px = (size_t)h * w;
    if (dtype == DT_F32) {
        const float* f = (const float*)img;
        float mn = f[0], mx = f[0];
        for (size_t p = 1; p < px; ++p) { if (f[p] < mn) mn = f[p]; if (f[p] > mx) mx = f[p]; }
        for (size_t p = 0; p < px; ++p) {
            float g;
            if (mn != mx) { float t = (f[p] - mn) / (mx - mn); g = t * 16383.0f + 0.0f; }
            else g = fminf(fmaxf(f[p], 0.0f), 16383.0f);
            grey[p] = (uint16_t)(int64_t)rintf(g);
        }
        return 0;
    }
    int vmin = sk_code(img, dtype, 0), vmax = vmin;
    for (size_t p = 1; p < px; ++p) { int v = sk_code(img, dtype, p); if (v < vmin) vmin = v; if (v > vmax) vmax = v; }
    const double fmin_ = sk_as_float(dtype, vmin), fmax_ = sk_as_float(dtype, vmax);
    for (size_t p = 0; p < px; ++p) {
        double f = sk_as_float(dtype, sk_code(img, dtype, p)), g;
        if (vmin != vmax) { double t = (f - fmin_) / (fmax_ - fmin_); g = t * 16383.0 + 0.0; }
        else g = fmin(fmax(f, 0.0), 16383.0);
        grey[p] = (uint16_t)(int64_t)rint(g); /* np.round: half to even; astype(uint16) */
    }
    return 0;
}

/* skimage.exposure._adapthist.clip_histogram on one histogram (in place). */
static void sk_clip_histogram(int64_t* hist, int nbins, int64_t clim) {
    int64_t n_excess = 0;
    for (int b = 0; b < nbins; ++b)
        if (hist[b] > clim) { n_excess += hist[b] - clim; hist[b] = clim; }
    const int64_t bin_incr = n_excess / nbins;
    const int64_t upper = clim - bin_incr;
    /* the mid mask is taken AFTER the low bins were raised (upstream evaluates it on the updated histogram), so a low
     * bin that lands in [upper, clim) is raised again, to the limit, and the excess shrinks accordingly */
    for (int b = 0; b < nbins; ++b)
        if (hist[b] < upper) { n_excess -= bin_incr; hist[b] += bin_incr; }
    for (int b = 0; b < nbins; ++b)
        if (hist[b] >= upper && hist[b] < clim) { n_excess += hist[b] - clim; hist[b] = clim; }
    while (n_excess > 0) {
        const int64_t prev = n_excess;
        for (int index = 0; index < nbins; ++index) {
            int64_t under = 0;
            for (int b = 0; b < nbins; ++b) under += hist[b] < clim;
            int64_t step = under / n_excess;
            if (step < 1) step = 1;
            int64_t given = 0;
            for (int64_t b = index; b < nbins; b += step)
                if (hist[b] < clim) { hist[b] += 1; ++given; }
            n_excess -= given;
            if (n_excess <= 0) break;
        }
        if (prev == n_excess) break;
    }
}

/* _clahe on the 2^14-level image: per-pixel coordinates instead of upstream's block reshapes.
 * Contextual regions (histogram blocks) tile the ORIGINAL image from its origin with kernel (kr, kc) — ceil(h/kr) x
 * ceil(w/kc) of them, the far ones completed by numpy 'reflect' padding; a pixel at (y, x) lies in interpolation block
 * ((y + kr/2) / kr, (x + kc/2) / kc) at position ((y + kr/2) % kr, (x + kc/2) % kc), blends the mappings of the regions
 * clamp(block - 1) and clamp(block) per axis with coefficients pos / k, accumulating float32(mapped * coef_x * coef_y) in
 * the order (row,col) = (0,0), (0,1), (1,0), (1,1), and truncates to uint16. */
int orc_sk_clahe(const uint16_t* grey, int h, int w, int kr, int kc, double clip_limit, int nbins, uint16_t* out,
                 int64_t* maps_out /* optional: nbr * nbc * nbins */) {
    if (kr <= 0 || kc <= 0 || nbins <= 0 || nbins > 16384) return -3;
    const int nbr = (h + kr - 1) / kr, nbc = (w + kc - 1) / kc;
    const int bin_size = 1 + 16384 / nbins;
    const int64_t kernel_elements = (int64_t)kr * kc;
    int64_t clim;
    if (clip_limit > 0.0) {
        double c = clip_limit * (double)kernel_elements;
        if (c < 1.0) c = 1.0;
        clim = (int64_t)c;
    } else {
        clim = 65535; /* np.iinfo(uint16).max: "no clipping" */
    }
    int64_t* maps = (int64_t*)malloc((size_t)nbr * nbc * nbins * sizeof(int64_t));
    if (!maps) return -1;
    const double scale = 16383.0 / (double)kernel_elements;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nbr * nbc; ++b) {
        const int bi = b / nbc, bj = b % nbc;
        int64_t* hist = maps + (size_t)b * nbins;
        memset(hist, 0, (size_t)nbins * sizeof(int64_t));
        for (int r = 0; r < kr; ++r) {
            const int sy = sk_reflect(bi * kr + r, h);
            for (int c = 0; c < kc; ++c) {
                const int sx = sk_reflect(bj * kc + c, w);
                hist[grey[(size_t)sy * w + sx] / bin_size] += 1;
            }
        }
        sk_clip_histogram(hist, nbins, clim);
        int64_t cum = 0;
        for (int k = 0; k < nbins; ++k) { /* map_histogram */
            cum += hist[k];
            double v = (double)cum * scale;
            v += 0.0;
            if (v > 16383.0) v = 16383.0;
            hist[k] = (int64_t)v;
        }
    }
    if (maps_out) memcpy(maps_out, maps, (size_t)nbr * nbc * nbins * sizeof(int64_t));
    const int r0 = kr / 2, c0 = kc / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const int py = y + r0, by = py / kr, ry = py % kr;
        const double cy = (double)ry / (double)kr;
        int hb_r[2] = {by - 1, by};
        for (int e = 0; e < 2; ++e) { if (hb_r[e] < 0) hb_r[e] = 0; if (hb_r[e] > nbr - 1) hb_r[e] = nbr - 1; }
        const double wy[2] = {1.0 - cy, cy};
        for (int x = 0; x < w; ++x) {
            const int pxx = x + c0, bx = pxx / kc, rx = pxx % kc;
            const double cx = (double)rx / (double)kc;
            int hb_c[2] = {bx - 1, bx};
            for (int e = 0; e < 2; ++e) { if (hb_c[e] < 0) hb_c[e] = 0; if (hb_c[e] > nbc - 1) hb_c[e] = nbc - 1; }
            const double wx[2] = {1.0 - cx, cx};
            const int bin = grey[(size_t)y * w + x] / bin_size;
            float acc = 0.0f;
            for (int er = 0; er < 2; ++er)
                for (int ec = 0; ec < 2; ++ec) {
                    const int64_t m = maps[((size_t)hb_r[er] * nbc + hb_c[ec]) * nbins + bin];
                    const double coef = wx[ec] * wy[er];
                    acc = acc + (float)((double)m * coef);
                }
            out[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    free(maps);
    return 0;
}

/* Last stage of equalize_adapthist: rescale_intensity(out_range=(0, 1)) of the uint16 CLAHE result as float64
 * (as_f32 != 0: float32 input images keep float32 arithmetic, like upstream). */
int orc_sk_rescale01(const uint16_t* c, int h, int w, int as_f32, void* out) {
    const size_t px = (size_t)h * w;
    uint16_t mn = c[0], mx = c[0];
    for (size_t p = 1; p < px; ++p) { if (c[p] < mn) mn = c[p]; if (c[p] > mx) mx = c[p]; }
    for (size_t p = 0; p < px; ++p) {
        if (as_f32) {
            float v = (float)c[p], r;
            if (mn != mx) { r = (v - (float)mn) / ((float)mx - (float)mn); r = r * 1.0f + 0.0f; }
            else r = fminf(fmaxf(v, 0.0f), 1.0f);
            ((float*)out)[p] = r;
        } else {
            double v = (double)c[p], r;
            if (mn != mx) { r = (v - (double)mn) / ((double)mx - (double)mn); r = r * 1.0 + 0.0; }
            else r = fmin(fmax(v, 0.0), 1.0);
            ((double*)out)[p] = r;
        }
    }
    return 0;
}

/* skimage.exposure.equalize_hist on one INTEGER image: one bin per integer value between min and max, cdf = cumsum /
 * total (float64), out = np.interp(image, centers, cdf) = cdf[v - min] because every pixel sits on a bin centre. */
int orc_sk_equalize_hist(const void* img, int dtype, int h, int w, double* out) {
    if (dtype == DT_F32) return -2;
    const size_t px = (size_t)h * w;
    int vmin = sk_code(img, dtype, 0), vmax = vmin;
    for (size_t p = 1; p < px; ++p) { int v = sk_code(img, dtype, p); if (v < vmin) vmin = v; if (v > vmax) vmax = v; }
    const int nb = vmax - vmin + 1;
    int64_t* hist = (int64_t*)calloc((size_t)nb, sizeof(int64_t));
    if (!hist) return -1;
    for (size_t p = 0; p < px; ++p) hist[sk_code(img, dtype, p) - vmin] += 1;
    for (int b = 1; b < nb; ++b) hist[b] += hist[b - 1];
    const double total = (double)hist[nb - 1];
    for (size_t p = 0; p < px; ++p) out[p] = (double)hist[sk_code(img, dtype, p) - vmin] / total;
    free(hist);
    return 0;
}

/* skimage.restoration.denoise_bilateral on one 2-D single-channel image of an integer dtype (float64 arithmetic) —
 * the loop of the upstream Cython kernel.  color_lut (bins entries) and range_lut (win*win entries, row-major) are
 * computed by the caller (numpy exp, as upstream does in Python).  mode: 0 constant(cval), 1 edge, 2 symmetric,
 * 3 reflect, 4 wrap.  Negative images are shifted by their minimum first and shifted back at the end. */
static inline int sk_border(int i, int n, int mode) {
    if (i >= 0 && i < n) return i;
    if (mode == 1) return i < 0 ? 0 : n - 1;
    if (mode == 2) { int p = 2 * n; int m = i % p; if (m < 0) m += p; return m < n ? m : p - 1 - m; }
    if (mode == 3) return sk_reflect(i, n);
    if (mode == 4) { int m = i % n; return m < 0 ? m + n : m; }
    return -1;
}
int orc_sk_bilateral(const void* img, int dtype, int h, int w, int win, int bins, int mode, double cval,
                     const double* color_lut, const double* range_lut, double* out) {
    if (dtype == DT_F32) return -2;
    const size_t px = (size_t)h * w;
    int vmin = sk_code(img, dtype, 0), vmax = vmin;
    for (size_t p = 1; p < px; ++p) { int v = sk_code(img, dtype, p); if (v < vmin) vmin = v; if (v > vmax) vmax = v; }
    const double min_value = sk_as_float(dtype, vmin);
    double max_value = sk_as_float(dtype, vmax);
    if (vmin == vmax) { for (size_t p = 0; p < px; ++p) out[p] = min_value; return 0; }
    const int shift = min_value < 0.0;
    if (shift) max_value = max_value - min_value;
    const double dist_scale = (double)bins / max_value;
    const int ext = (win - 1) / 2;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < h; ++r)
        for (int c = 0; c < w; ++c) {
            double centre = sk_as_float(dtype, sk_code(img, dtype, (size_t)r * w + c));
            if (shift) centre = centre - min_value;
            double total_v = 0.0, total_w = 0.0;
            for (int kr = 0; kr < win; ++kr) {
                const int rr = sk_border(r + kr - ext, h, mode);
                for (int kc = 0; kc < win; ++kc) {
                    const int cc = sk_border(c + kc - ext, w, mode);
                    double v;
                    if (rr < 0 || cc < 0) v = cval; /* np.pad(constant) happens AFTER the shift: the pad value is cval itself */
                    else { v = sk_as_float(dtype, sk_code(img, dtype, (size_t)rr * w + cc)); if (shift) v = v - min_value; }
                    const double t = centre - v;
                    const double dist = sqrt(t * t);
                    int64_t b = (int64_t)(dist * dist_scale);
                    if (b > bins - 1) b = bins - 1;
                    const double weight = range_lut[kr * win + kc] * color_lut[b];
                    total_v = total_v + v * weight;
                    total_w = total_w + weight;
                }
            }
            double o = total_v / total_w;
            if (shift) o = o + min_value;
            out[(size_t)r * w + c] = o;
        }
    return 0;
}
