"""CPU tests that PIN the oracle (no GPU): the reference repository has no golden vectors
("parity unpinned"), so every restatement is checked against the independent
binaries this image does have — cv2 4.13, scipy.ndimage, torchvision — and against the
torch-CPU kornia twin (oracle/kornia_twin.py), plus the committed golden hashes."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
import kornia_twin as K

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def smooth_u8(h, w, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 128 + 60 * np.sin(xx / 37.0) + 50 * np.cos(yy / 23.0) + rng.normal(0, 12, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------- A1' OpenCV CLAHE
@pytest.mark.parametrize("h,w,grid,clip", [
    (512, 512, (8, 8), 2.0), (512, 512, (8, 8), 40.0), (300, 500, (8, 8), 2.0), (1024, 1024, (16, 16), 2.0),
    (512, 512, (8, 8), 0.0), (512, 512, (4, 4), 300.0), (37, 53, (3, 5), 1.5), (64, 64, (8, 8), 0.3)])
def test_opencv_clahe_oracle_is_bit_exact_against_cv2(h, w, grid, clip):
    cv2 = pytest.importorskip("cv2")
    img = smooth_u8(h, w)
    ref = cv2.createCLAHE(clip, (grid[1], grid[0])).apply(img)
    assert np.array_equal(O.opencv_clahe(img, clip, grid), ref)


def smooth_u16(h, w, seed=0, bits=16):
    """Smooth-plus-noise image using `bits` of a uint16 container (12 = CT-like occupancy)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.5 + 0.23 * np.sin(xx / 37.0) + 0.2 * np.cos(yy / 23.0) + rng.normal(0, 0.05, (h, w))
    return (np.clip(img, 0, 1) * (2 ** bits - 1)).astype(np.uint16)


@pytest.mark.parametrize("h,w,grid,clip,bits", [
    (512, 512, (8, 8), 2.0, 16), (512, 512, (8, 8), 40.0, 16), (300, 500, (8, 8), 2.0, 16), (1024, 1024, (16, 16), 2.0, 12),
    (512, 512, (8, 8), 0.0, 16), (512, 512, (4, 4), 300.0, 12), (37, 53, (3, 5), 1.5, 16), (64, 64, (8, 8), 0.3, 16),
    (512, 512, (2, 2), 2.0, 8), (256, 256, (1, 1), 2.0, 4), (512, 512, (8, 8), 40000.0, 12)])
def test_opencv_clahe16_oracle_is_bit_exact_against_cv2(h, w, grid, clip, bits):
    """The 65 536-bin mode of cv::CLAHE (uint16 in, uint16 out): includes tiles of >= 65 536 pixels
    (2x2 and 1x1 grids), few-level images (heavy clipping) and a clip limit that never binds."""
    cv2 = pytest.importorskip("cv2")
    img = smooth_u16(h, w, seed=3, bits=bits)
    ref = cv2.createCLAHE(clip, (grid[1], grid[0])).apply(img)
    assert np.array_equal(O.opencv_clahe(img, clip, grid), ref)


def test_opencv_clahe16_constant_and_extremes_against_cv2():
    cv2 = pytest.importorskip("cv2")
    for img in (np.full((128, 128), 777, np.uint16), np.zeros((128, 128), np.uint16),
                np.full((128, 128), 65535, np.uint16),
                np.random.default_rng(5).integers(0, 65536, (128, 192), dtype=np.uint16)):
        for clip in (2.0, 0.0):
            ref = cv2.createCLAHE(clip, (4, 4)).apply(img)
            assert np.array_equal(O.opencv_clahe(img, clip, (4, 4)), ref)


def _cv2_golden_cases():
    with open(os.path.join(GOLDEN, "cv2_clahe.json")) as f:
        return json.load(f)["cases"]


def _sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(_cv2_golden_cases()))
def test_oracle_reproduces_committed_cv2_vectors(name):
    """tests/golden/cv2_clahe.json holds sha256 of cv2.createCLAHE outputs (made by make_cv2_golden.py):
    the oracle must reproduce them without cv2 being importable (the GPU box runs the same check)."""
    sys_path_golden = os.path.join(GOLDEN)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_cv2_golden_inputs", os.path.join(sys_path_golden, "cv2_inputs.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    c = _cv2_golden_cases()[name]
    img = mod.image(c["h"], c["w"], np.dtype(c["dtype"]), c["seed"], c["bits"])
    assert _sha(img) == c["input_sha256"]
    assert _sha(O.opencv_clahe(img, c["clip"], tuple(c["grid"]))) == c["output_sha256"]


def test_oracle_reproduces_cv2_fixture():
    z = np.load(os.path.join(GOLDEN, "cv2_clahe16_128.npz"))
    assert np.array_equal(O.opencv_clahe(z["input"], float(z["clip"]), tuple(int(v) for v in z["grid"])), z["output"])


# ---------------------------------------------------------------------------- A1 kornia CLAHE vs twin
@pytest.mark.parametrize("h,w,grid,clip", [
    (512, 512, (8, 8), 2.0), (64, 64, (8, 8), 40.0), (100, 130, (4, 6), 2.0), (20, 20, (8, 8), 1.0),
    (57, 91, (2, 3), 0.0), (37, 41, (1, 1), 3.0), (96, 64, (3, 2), 0.7)])
def test_kornia_clahe_oracle_matches_tensor_level_twin(h, w, grid, clip):
    rng = np.random.default_rng(1)
    x = rng.random((2, h, w)).astype(np.float32)
    x[0, : h // 3] = 0.0   # constant "air"
    x[1, -1, -1] = 1.0     # the x == 1 edge of torch.histc
    luts = O.clahe_luts(x, clip, grid)
    twin_luts = K.compute_luts(torch.from_numpy(x)[:, None], clip, grid).numpy()[:, :, :, 0]
    assert np.array_equal(luts, twin_luts.astype(np.uint8))
    out = O.clahe_apply(x, luts, grid)
    twin = K.equalize_clahe(torch.from_numpy(x)[:, None], clip, grid).numpy()[:, 0]
    assert np.abs(out - twin).max() <= 1e-6
    hist = O.clahe_hist(x, grid)
    th, tw = O.kornia_tile_size(h, w, grid)
    assert (hist.sum(-1) == th * tw).all()
    assert (np.diff(luts.astype(int), axis=-1) >= 0).all()


def test_kornia_clahe_errors_like_kornia():
    x = np.zeros((3, 3), np.float32)
    with pytest.raises(ValueError):
        O.clahe_luts(x, 2.0, (8, 8))
    with pytest.raises(ValueError):
        K.compute_luts(torch.zeros(1, 1, 3, 3), 2.0, (8, 8))
    with pytest.raises(ValueError):
        O.clahe_hist(np.zeros((8, 8), np.float32), (0, 8))


def test_histc_rule_and_lookup_quirk():
    """torch.histc(bins=256,min=0,max=1) == min(floor(x*256),255); the lookup index trunc(x*255)
    differs from it for about half of the 16-bit codes — keep the quirk."""
    v = (np.arange(65536, dtype=np.float32) / np.float32(65535.0)).astype(np.float32)
    bins = np.minimum(np.floor(v * np.float32(256.0)), 255).astype(int)
    hist = torch.histc(torch.from_numpy(v), bins=256, min=0, max=1).numpy()
    assert np.array_equal(np.bincount(bins, minlength=256), hist.astype(int))
    h = O.clahe_hist(v.reshape(256, 256), (1, 1)).reshape(256)
    assert np.array_equal(h, hist.astype(np.uint32))
    idx = (v * np.float32(255.0)).astype(int)
    assert 30000 < int((idx != bins).sum()) < 35000


# ---------------------------------------------------------------------------- A3 / A4 Gaussian, unsharp
@pytest.mark.parametrize("k,s", [(9, 1.0), (7, 1.0), (5, 1.0), (9, 2.0), (3, 0.8)])
def test_gaussian_oracle_against_cv2_scipy_and_twin(k, s):
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    x = np.random.default_rng(0).random((3, 200, 333)).astype(np.float32)
    got = O.gaussian_blur2d(x, k, s)
    a = np.stack([cv2.GaussianBlur(p, (k, k), s, borderType=cv2.BORDER_REFLECT_101) for p in x])
    b = np.stack([ndi.gaussian_filter(p, s, mode="mirror", radius=k // 2) for p in x])
    t = K.gaussian_blur2d(torch.from_numpy(x)[:, None], k, s).numpy()[:, 0]
    assert np.abs(got - a).max() <= 5e-7
    assert np.abs(got - b).max() <= 5e-7
    assert np.abs(got - t).max() <= 5e-7
    u = O.unsharp_mask(x, k, s)
    ut = K.unsharp_mask(torch.from_numpy(x)[:, None], k, s).numpy()[:, 0]
    assert np.abs(u - ut).max() <= 1e-6


@pytest.mark.parametrize("border", ["reflect", "replicate", "constant", "circular"])
def test_gaussian_borders_match_torch_pad(border):
    x = np.random.default_rng(2).random((1, 40, 33)).astype(np.float32)
    got = O.gaussian_blur2d(x, (5, 9), (1.0, 1.7), border)
    t = K.gaussian_blur2d(torch.from_numpy(x)[:, None], (5, 9), (1.0, 1.7), border).numpy()[:, 0]
    assert np.abs(got - t).max() <= 5e-7


def test_gaussian_weights_and_constant_image():
    import cv2

    w = O.gaussian_kernel1d(7, 1.0)
    assert np.abs(w - cv2.getGaussianKernel(7, 1.0, cv2.CV_32F).ravel()).max() <= 1e-7
    assert np.abs(w - K.get_gaussian_kernel1d(7, 1.0).numpy()).max() <= 1e-7
    c = np.full((1, 32, 32), 0.37, np.float32)
    assert np.abs(O.gaussian_blur2d(c, 9, 1.0) - 0.37).max() <= 1e-7


def test_pixel_mapping_round_trip_and_division():
    v = np.arange(65536, dtype=np.uint16)
    x = O.to01(v)
    assert np.array_equal(x, v.astype(np.float32) / np.float32(65535.0))
    assert np.array_equal(O.from01(x, np.uint16), v)
    s = np.arange(-32768, 32768).astype(np.int16)
    assert np.array_equal(O.from01(O.to01(s), np.int16), s)
    b = np.arange(256, dtype=np.uint8)
    assert np.array_equal(O.from01(O.to01(b), np.uint8), b)
    # configurable value range (HU window) and clamping
    hu = np.array([-2000, -1024, 0, 3071, 5000], np.int16)
    y = O.to01(hu, (-1024.0, 3071.0))
    assert y[1] == 0.0 and y[3] == 1.0
    assert np.array_equal(O.from01(y, np.int16, (-1024.0, 3071.0)), np.array([-1024, -1024, 0, 3071, 3071], np.int16))


# ---------------------------------------------------------------------------- A5 / A6 median
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_median2d_oracle_against_scipy_cv2_and_twin(dtype):
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(5)
    if dtype == np.float32:
        x = rng.random((50, 70)).astype(np.float32)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max + 1, (50, 70)).astype(dtype)
    for k in (3, 5):
        assert np.array_equal(O.median_blur(x, k), ndi.median_filter(x, size=k, mode="constant", cval=0))
        assert np.array_equal(O.median_blur(x, k, "replicate"), ndi.median_filter(x, size=k, mode="nearest"))
        if dtype != np.float32 or True:
            assert np.array_equal(O.median_blur(x, k, "replicate"), cv2.medianBlur(x, k))
        t = K.median_blur(torch.from_numpy(x.astype(np.float32))[None, None], k).numpy()[0, 0]
        assert np.array_equal(O.median_blur(x, k).astype(np.float32), t)
    assert np.array_equal(O.median_blur(x, (3, 5)), ndi.median_filter(x, size=(3, 5), mode="constant", cval=0))


def test_median3d_oracle_against_scipy_and_slab_halo_consistency():
    ndi = pytest.importorskip("scipy.ndimage")
    vol = np.random.default_rng(6).integers(-1024, 3072, (12, 20, 24)).astype(np.int16)
    ref = ndi.median_filter(vol, size=3, mode="nearest")
    assert np.array_equal(O.median3d(vol), ref)
    assert np.array_equal(O.median3d(vol, mode="constant"), ndi.median_filter(vol, size=3, mode="constant", cval=0))
    # two z-slabs with one halo plane each reproduce the unsharded result
    a = O.median3d(vol[:5], halo_hi=vol[5])
    b = O.median3d(vol[5:], halo_lo=vol[4])
    assert np.array_equal(np.concatenate([a, b]), ref)


# ---------------------------------------------------------------------------- A2 global equalisation
def test_equalize_oracle_against_torchvision_uint8_and_twin():
    tvf = pytest.importorskip("torchvision.transforms.v2.functional")
    rng = np.random.default_rng(7)
    imgs = [smooth_u8(64, 80, 1), rng.integers(0, 256, (40, 40)).astype(np.uint8), np.full((16, 16), 9, np.uint8),
            (rng.integers(0, 4, (32, 32)) * 60).astype(np.uint8)]
    for img in imgs:
        ref = tvf.equalize(torch.from_numpy(img)[None]).numpy()[0]
        got = O.from01(O.equalize(O.to01(img)), np.uint8)
        assert np.array_equal(got, ref)
    x = rng.random((2, 1, 48, 48)).astype(np.float32)
    t = K.equalize(torch.from_numpy(x)).numpy()
    assert np.abs(O.equalize(x) - t).max() <= 1e-6


# ---------------------------------------------------------------------------- A7 bilateral
def test_mie_exp2n_accuracy():
    """The base-2 exponential of the bilateral colour weight: max rel err 1.7e-7 on [-125, 0], exact at 0,
    clamped below -125, NaN -> 2^-125 (never a NaN weight)."""
    import oracle as O

    t = np.linspace(-125.0, 0.0, 500001).astype(np.float32)
    got = O.mie_exp2n(t).astype(np.float64)
    assert np.abs(got / np.exp2(t.astype(np.float64)) - 1.0).max() < 2.5e-7
    assert O.mie_exp2n(np.zeros(1, np.float32))[0] == 1.0
    lo = O.mie_exp2n(np.array([-125.0, -126.0, -1e30, -np.inf, np.nan], np.float32))
    assert np.all(lo == np.float32(2.0 ** -125))


def test_mie_exp_accuracy():
    a = -np.abs(np.random.default_rng(8).normal(0, 8, 200000)).astype(np.float32)
    a = np.concatenate([a, np.array([0.0, -1e-8, -87.0, -100.0, -0.5], np.float32)])
    got = O.mie_exp(a).astype(np.float64)
    ref = np.exp(np.maximum(a.astype(np.float64), -87.0))
    assert (np.abs(got - ref) / ref).max() <= 3e-7


def test_bilateral_oracle_against_twin():
    x = np.random.default_rng(9).random((2, 40, 44)).astype(np.float32)
    for border in ("reflect", "replicate", "constant"):
        got = O.bilateral_blur(x, (5, 7), 0.1, (1.5, 1.2), border)
        t = K.bilateral_blur(torch.from_numpy(x)[:, None], (5, 7), 0.1, (1.5, 1.2), border).numpy()[:, 0]
        assert (np.abs(got - t) / np.maximum(np.abs(t), 1e-6)).max() <= 1e-5   # north-star tolerance


# ---------------------------------------------------------------------------- golden vectors
def test_golden_hashes():
    """Seeded inputs -> sha256 of oracle outputs (tests/golden/make_golden.py wrote them).  A change here
    means the oracle's arithmetic drifted; the GPU parity tests check the same hashes."""
    import make_golden

    with open(os.path.join(GOLDEN, "golden.json")) as f:
        want = json.load(f)
    got = make_golden.compute()
    assert got == want


# ---------------------------------------------------------------------------- A8 non-local means
@pytest.mark.parametrize("ps,pd,h,sigma", [(3, 2, 0.1, 0.0), (5, 3, 0.08, 0.05), (7, 4, 0.15, 0.02), (4, 2, 0.1, 0.0)])
def test_nlm_closed_form_equals_literal_upstream_loops(ps, pd, h, sigma):
    """orc_nlm_fast (per-pixel closed form) vs a literal transcription of skimage's fast-mode loops
    (integral image per shift, symmetric accumulation, alpha = 0.5 on the t_col == 0 column)."""
    rng = np.random.default_rng(0)
    img = np.clip(0.5 + 0.2 * np.sin(np.arange(28)[:, None] / 3.0) + 0.1 * rng.normal(size=(28, 30)), 0, 1)
    a = O.denoise_nl_means(img, ps, pd, h, sigma)
    b = O.nlm_fast_literal(img, ps, pd, h, sigma)
    assert np.abs(a - b).max() < 1e-12


@pytest.mark.parametrize("ps,pd,h,sigma", [(3, 4, 0.08, 0.05), (5, 3, 0.1, 0.0), (7, 6, 0.05, 0.0), (4, 2, 0.15, 0.02)])
def test_nlm_slow_mode_equals_literal_upstream_loops(ps, pd, h, sigma):
    """orc_nlm_slow vs a literal transcription of skimage's _nl_means_denoising_2d / patch_distance_2d (np.pad by the
    patch radius, meshgrid Gaussian weights, clipped search window, cut-off before every patch row)."""
    rng = np.random.default_rng(3)
    img = np.clip(0.5 + 0.2 * rng.standard_normal((14, 17)), 0, 1)
    img[:6, :8] += 0.3
    a = O.denoise_nl_means_slow(img, ps, pd, h, sigma)
    b = O.nlm_slow_literal(img, ps, pd, h, sigma)
    assert np.abs(a - b).max() < 1e-14
    # the row-wise cut-off matters: h small enough that many patches are cut
    if h <= 0.05:
        assert (np.abs(a - img) < 1e-3).mean() > 0.2


def test_nlm_slow_mode_properties():
    const = np.full((20, 24), 0.3)
    assert np.allclose(O.denoise_nl_means_slow(const, 5, 4, 0.1), 0.3, atol=1e-15)
    rng = np.random.default_rng(1)
    noisy = np.clip(0.5 + 0.05 * rng.normal(size=(32, 32)), 0, 1)
    den = O.denoise_nl_means_slow(noisy, 5, 4, 0.1, sigma=0.05)
    assert den.std() < 0.6 * noisy.std()
    assert den.min() >= noisy.min() - 1e-12 and den.max() <= noisy.max() + 1e-12
    import ctypes
    w = np.zeros(49)
    assert O.lib().orc_nlm_patch_weights(7, 0.1, w.ctypes.data_as(ctypes.c_void_p)) == 7
    assert abs(w.sum() - 100.0) < 1e-9 and w[24] == w.max() and np.allclose(w.reshape(7, 7), w.reshape(7, 7).T)
    with pytest.raises(ValueError):
        O.denoise_nl_means_slow(np.zeros((3, 8)), 7, 2, 0.1)


def test_nlm_properties():
    rng = np.random.default_rng(1)
    const = np.full((40, 40), 0.3)
    assert np.allclose(O.denoise_nl_means(const, 7, 5, 0.1), 0.3, atol=1e-15)      # constants are fixed points
    noisy = np.clip(0.5 + 0.05 * rng.normal(size=(48, 48)), 0, 1)
    den = O.denoise_nl_means(noisy, 7, 5, 0.1, sigma=0.05)
    assert den.std() < 0.5 * noisy.std()                                            # it denoises
    assert den.min() >= noisy.min() - 1e-12 and den.max() <= noisy.max() + 1e-12    # convex combination
    with pytest.raises(ValueError):
        O.denoise_nl_means(np.zeros((8, 8)), 7, 11, 0.1)
