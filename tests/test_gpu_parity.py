"""GPU parity tests (run with -m gpu on a B200): CUDA path through the C ABI vs the CPU oracle.

Bars: integer artefacts (histograms, LUTs, quantised outputs, OpenCV-mode
pixels) bit-exact; float outputs: the kernels follow the oracle's fp32 operation order, so
they are asserted bit-exact too, with the north star's tolerance (rel 1e-5) as the documented
fallback bar `RTOL` used only where an op cannot be made bit-reproducible.
"""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north-star tolerance for floating-point stages (pre-quantisation)

NP2T = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16, np.dtype(np.int16): torch.int16,
        np.dtype(np.float32): torch.float32}


def gpu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def cpu(t):
    return t.cpu().numpy()


def images(kind, shape, dtype, seed=0):
    from mie_b200 import synthetic

    if dtype == np.float32:
        if kind == "K":
            return np.full(shape, 0.25, np.float32)
        if kind == "U":
            return np.random.default_rng(seed).random(shape, dtype=np.float32)
        return (synthetic.phantom(shape, np.uint16, seed).astype(np.float32) / np.float32(4095.0)).astype(np.float32)
    return synthetic.make(kind, shape, dtype, seed)


# ---------------------------------------------------------------------------- CLAHE (kornia semantics)
CLAHE_CASES = [
    # (shape, grid, clip)
    ((2, 1, 512, 512), (8, 8), 2.0),
    ((1, 1, 512, 512), (8, 8), 40.0),
    ((1, 2, 100, 130), (4, 6), 2.0),   # padding + odd->even tile bump
    ((1, 1, 20, 20), (8, 8), 1.0),     # pad (12) larger than a tile (4)
    ((1, 1, 57, 91), (2, 3), 0.0),     # clipping disabled
    ((1, 1, 37, 41), (1, 1), 3.0),     # single tile
    ((1, 1, 256, 1024), (2, 16), 0.7),
]


@pytest.mark.parametrize("kind", ["U", "P", "K"])
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8, np.int16])
def test_clahe_hist_and_luts_bit_exact(dev, kind, dtype):
    import mie_b200 as M
    import oracle as O

    for shape, grid, clip in CLAHE_CASES:
        x = images(kind, shape, dtype, seed=11)
        x01 = O.to01(x)
        h_ref = O.clahe_hist(x01, grid)
        l_ref = O.clahe_luts(x01, clip, grid)
        xt = gpu(x, dev)
        h = cpu(M.clahe_histograms(xt, grid))
        l = cpu(M.clahe_luts(xt, clip, grid))
        assert np.array_equal(h.astype(np.uint32), h_ref), (shape, grid)
        assert np.array_equal(l, l_ref), (shape, grid, clip)
        # properties: mass conserved, LUT monotone
        th, tw = O.kornia_tile_size(shape[-2], shape[-1], grid)
        assert (h.sum(-1) == th * tw).all()
        assert (np.diff(l.astype(np.int32), axis=-1) >= 0).all()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_clahe_many_tiles_warp_per_tile_kernel(dev, dtype):
    """Jobs with >= 592 tiles take the warp-per-tile LUT kernel (smaller ones the block-per-tile latency
    variant): 32-pixel tiles (4 chunks per row), 64-pixel tiles, and 24-pixel tiles (3 chunks: the generic
    index loop)."""
    import mie_b200 as M
    import oracle as O

    for shape, grid in [((12, 1, 256, 256), (8, 8)), ((10, 1, 512, 512), (8, 8)), ((20, 1, 192, 192), (8, 8)),
                        ((3, 1, 1024, 512), (16, 16))]:
        x = images("U", shape, dtype, seed=21)
        x01 = O.to01(x)
        h_ref = O.clahe_hist(x01, grid)
        l_ref = O.clahe_luts(x01, 2.0, grid)
        xt = gpu(x, dev)
        assert np.array_equal(cpu(M.clahe_histograms(xt, grid)).reshape(h_ref.shape), h_ref), shape
        assert np.array_equal(cpu(M.clahe_luts(xt, 2.0, grid)).reshape(l_ref.shape), l_ref), shape
        ref = O.equalize_clahe(x01, 2.0, grid)
        assert np.array_equal(cpu(M.equalize_clahe(xt, 2.0, grid, out_dtype=torch.float32)), ref), shape


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8, np.int16])
def test_clahe_apply_teacher_forced(dev, dtype):
    import mie_b200 as M
    import oracle as O

    for shape, grid, clip in CLAHE_CASES:
        x = images("P", shape, dtype, seed=5)
        x01 = O.to01(x)
        luts = O.clahe_luts(x01, clip, grid)
        ref = O.clahe_apply(x01, luts, grid)
        xt, lt = gpu(x, dev), gpu(luts, dev)
        got_f = cpu(M.clahe_apply(xt, lt, grid, out_dtype=torch.float32))
        assert np.array_equal(got_f, ref), f"max abs diff {np.abs(got_f - ref).max()}"
        if dtype != np.float32:
            got_q = cpu(M.clahe_apply(xt, lt, grid))
            assert np.array_equal(got_q, O.from01(ref, dtype))


def test_clahe_end_to_end_and_value_range(dev):
    import mie_b200 as M
    import oracle as O

    x = images("P", (3, 1, 512, 512), np.int16, seed=2)  # HU-like [-1024, 3071]
    vr = (-1024.0, 3071.0)
    ref = O.from01(O.equalize_clahe(O.to01(x, vr), 2.0, (8, 8)), np.int16, vr)
    got = cpu(M.equalize_clahe(gpu(x, dev), 2.0, (8, 8), value_range=vr))
    assert np.array_equal(got, ref)
    # shapes (H,W) and (C,H,W) keep their rank
    assert M.equalize_clahe(gpu(x[0, 0], dev), 2.0, (8, 8)).shape == (512, 512)
    assert M.equalize_clahe(gpu(x[:, 0], dev), 2.0, (8, 8)).shape == (3, 512, 512)
    # kornia raises ValueError when the grid cannot tile the image
    with pytest.raises(ValueError):
        M.equalize_clahe(gpu(np.zeros((3, 3), np.float32), dev), 2.0, (8, 8))


@pytest.mark.parametrize("case", [
    (np.int16, (-1024.0, 3071.0)),     # HU window of CT data
    (np.uint16, (0.0, 4095.0)),        # 12-bit data in a 16-bit container
    (np.uint16, (1000.0, 60000.0)),
    (np.uint8, (10.0, 200.0)),
    (np.int16, (-32768.0, 0.0)),
])
def test_clahe_integer_windows_take_the_tuned_kernels_bit_exactly(dev, case):
    """value_range windows with integer bounds run on the tuned kernels through the divide-free windowed
    conversion (host-verified against the IEEE quotient for every code of the dtype); pixels outside the
    window are ignored by the histograms and clamped by the lookup, exactly as in the generic kernels."""
    import mie_b200 as M
    import oracle as O

    dtype, vr = case
    for shape, grid in [((2, 1, 512, 512), (8, 8)), ((10, 1, 512, 512), (8, 8)), ((12, 1, 256, 256), (8, 8)),
                        ((1, 1, 256, 1024), (2, 16))]:
        x = images("U", shape, dtype, seed=31)          # full dtype range: many pixels outside the window
        x[..., : shape[-2] // 3, :] = images("P", shape, dtype, seed=32)[..., : shape[-2] // 3, :]
        x01 = O.to01(x, vr)
        xt = gpu(x, dev)
        h_ref = O.clahe_hist(x01, grid)
        assert np.array_equal(cpu(M.clahe_histograms(xt, grid, value_range=vr)).reshape(h_ref.shape), h_ref), shape
        l_ref = O.clahe_luts(x01, 2.0, grid)
        assert np.array_equal(cpu(M.clahe_luts(xt, 2.0, grid, value_range=vr)).reshape(l_ref.shape), l_ref), shape
        ref = O.equalize_clahe(x01, 2.0, grid)
        got = cpu(M.equalize_clahe(xt, 2.0, grid, value_range=vr, out_dtype=torch.float32))
        assert np.array_equal(got, ref), shape
        assert np.array_equal(cpu(M.equalize_clahe(xt, 2.0, grid, value_range=vr)), O.from01(ref, dtype, vr)), shape


def test_float_planes_outside_unit_range_and_nan(dev):
    """kornia's native input is float: pixels outside [0, 1] and NaN are ignored by the histograms and clamped by
    the lookup (NaN -> index 0) on the tuned CLAHE kernels and in the fused chain, exactly as in the oracle."""
    import mie_b200 as M
    import oracle as O

    rng = np.random.default_rng(77)
    for shape, grid in [((2, 1, 512, 512), (8, 8)), ((12, 1, 256, 256), (8, 8))]:
        x = (rng.random(shape, dtype=np.float32) * 1.5 - 0.25).astype(np.float32)
        x[0, 0, 10, 30:90] = np.nan
        x[-1, 0, 200:203, :] = np.inf
        xt = gpu(x, dev)
        h_ref = O.clahe_hist(x, grid)
        assert np.array_equal(cpu(M.clahe_histograms(xt, grid)).reshape(h_ref.shape), h_ref)
        l_ref = O.clahe_luts(x, 2.0, grid)
        assert np.array_equal(cpu(M.clahe_luts(xt, 2.0, grid)).reshape(l_ref.shape), l_ref)
        assert np.array_equal(cpu(M.equalize_clahe(xt, 2.0, grid)), O.equalize_clahe(x, 2.0, grid), equal_nan=True)
        # fused chain, every schedule: NaN / inf spread through the Gaussians; bins and lookups follow the same rules
        ref = O.chain_gauss_clahe_unsharp(x, 9, 1.0, 2.0, grid, 9, 1.0, "reflect")
        cfg = M.ChainConfig(grid_size=grid)
        for stages in (3, 3 | 4, 3 | 8):
            assert np.array_equal(cpu(M.enhance_chain(xt, cfg, stages=stages)), ref, equal_nan=True), (shape, stages)


def test_clahe_constant_image_stays_constant(dev):
    import mie_b200 as M

    y = cpu(M.equalize_clahe(gpu(np.full((1, 1, 512, 512), 0.5, np.float32), dev), 2.0, (8, 8)))
    assert (y == y.flat[0]).all()


def test_clahe_opencv_semantics_matches_cv2(dev):
    cv2 = pytest.importorskip("cv2")
    import mie_b200 as M

    rng = np.random.default_rng(0)
    for (h, w, g, clip) in [(512, 512, (8, 8), 2.0), (512, 512, (8, 8), 40.0), (300, 500, (8, 8), 2.0),
                            (1024, 1024, (16, 16), 2.0), (512, 512, (8, 8), 0.0), (512, 512, (4, 4), 300.0),
                            (37, 53, (3, 5), 1.5)]:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.clip(128 + 60 * np.sin(xx / 37.0) + 50 * np.cos(yy / 23.0) + rng.normal(0, 12, (h, w)), 0, 255)
        img = img.astype(np.uint8)
        ref = cv2.createCLAHE(clip, (g[1], g[0])).apply(img)
        got = cpu(M.equalize_clahe(gpu(img, dev), float(clip), g, semantics="opencv"))
        assert np.array_equal(got, ref), (h, w, g, clip, int((got != ref).sum()))


# ---------------------------------------------------------------------------- CLAHE, OpenCV semantics, 65 536 bins
def _cv2_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_clahe.json")) as f:
        return json.load(f)["cases"]


def _golden_sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(_cv2_cases()))
def test_clahe_opencv_reproduces_committed_cv2_vectors(dev, name):
    """CUDA path vs the vectors cv2.createCLAHE produced (tests/golden/make_cv2_golden.py): bit-exact,
    for uint8 (256 bins) and uint16 (65 536 bins); and equal to the oracle incl. the uint16 LUTs."""
    import os
    import sys
    import mie_b200 as M
    import oracle as O

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from cv2_inputs import image

    c = _cv2_cases()[name]
    img = image(c["h"], c["w"], np.dtype(c["dtype"]), c["seed"], c["bits"])
    grid = tuple(c["grid"])
    got = cpu(M.equalize_clahe(gpu(img, dev), float(c["clip"]), grid, semantics="opencv"))
    assert _golden_sha(got) == c["output_sha256"]
    if c["dtype"] == "uint16":
        ref, luts = O.opencv_clahe16(img, c["clip"], grid, return_luts=True)
        assert np.array_equal(got, ref)
        assert np.array_equal(cpu(M.clahe16_luts(gpu(img, dev), float(c["clip"]), grid)), luts[0])


def test_clahe16_batches_groups_and_edge_images(dev):
    """Batches larger than the LUT workspace (processed in groups), constant / extreme images, tiles of
    more than 65 535 equal pixels (a 16-bit counter would wrap), strided batches."""
    import mie_b200 as M
    from mie_b200 import enhance as E
    import oracle as O

    rng = np.random.default_rng(3)
    x = rng.integers(0, 4096, (5, 1, 128, 192), dtype=np.uint16)
    x[1] = 777
    x[2] = 0
    x[3] = 65535
    ref = O.opencv_clahe(x, 2.0, (4, 4))
    old = E.CLAHE16_WORKSPACE_BYTES
    try:
        for budget in (old, 2 * 16 * 131072 + 5, 1):       # all at once / groups of 2 / one image at a time
            E.CLAHE16_WORKSPACE_BYTES = budget
            assert np.array_equal(cpu(M.equalize_clahe(gpu(x, dev), 2.0, (4, 4), semantics="opencv")), ref)
    finally:
        E.CLAHE16_WORKSPACE_BYTES = old
    big = np.full((1, 1, 512, 512), 1234, np.uint16)         # one 262 144-pixel tile, single grey level
    for clip in (2.0, 0.0):
        assert np.array_equal(cpu(M.equalize_clahe(gpu(big, dev), clip, (1, 1), semantics="opencv")),
                              O.opencv_clahe(big, clip, (1, 1)))
    with pytest.raises((ValueError, TypeError, RuntimeError)):
        M.equalize_clahe(gpu(x.astype(np.int16), dev), 2.0, (4, 4), semantics="opencv")


# ---------------------------------------------------------------------------- Gaussian / unsharp
GAUSS_CASES = [
    # (shape, kernel_size, sigma, border)
    ((2, 1, 512, 512), 9, 1.0, "reflect"),
    ((1, 1, 100, 130), 7, 1.0, "reflect"),
    ((1, 2, 65, 63), 5, 1.2, "replicate"),
    ((1, 1, 64, 64), 3, 0.8, "constant"),
    ((1, 1, 33, 200), 9, 2.0, "circular"),
    ((1, 1, 70, 70), 11, 1.5, "reflect"),        # generic path (K > 9)
    ((1, 1, 50, 60), (3, 9), (0.7, 1.5), "reflect"),  # generic path (non-square)
    ((1, 1, 9, 9), 9, 1.0, "reflect"),           # halo == dim-1 (largest legal reflect)
    ((1, 1, 40, 40), 33, 5.0, "replicate"),      # maximum taps
    # marching kernel (9 taps, W % 128 == 0, H % 64 == 0): one-warp bands, 256-thread bands, all borders
    ((2, 1, 128, 128), 9, 1.0, "reflect"),
    ((2, 1, 128, 256), 9, 1.0, "replicate"),
    ((1, 1, 64, 128), 9, 1.5, "constant"),
    ((1, 1, 192, 1024), 9, 1.0, "reflect"),
    ((1, 3, 64, 384), 9, 0.7, "reflect"),
]


@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8, np.int16])
@pytest.mark.parametrize("op", ["gaussian_blur2d", "unsharp_mask"])
def test_gaussian_and_unsharp_bit_exact(dev, dtype, op):
    import mie_b200 as M
    import oracle as O

    for shape, k, s, border in GAUSS_CASES:
        x = images("U", shape, dtype, seed=7)
        ref = getattr(O, op)(O.to01(x), k, s, border)
        xt = gpu(x, dev)
        got = cpu(getattr(M, op)(xt, k, s, border, out_dtype=torch.float32))
        assert np.array_equal(got, ref), (shape, k, border, float(np.abs(got - ref).max()))
        if dtype != np.float32:
            gq = cpu(getattr(M, op)(xt, k, s, border))
            assert np.array_equal(gq, O.from01(ref, dtype)), (shape, k, border)


@pytest.mark.parametrize("case", [(np.int16, (-1024.0, 3071.0)), (np.uint16, (0.0, 4095.0)), (np.uint8, (10.0, 200.0))])
@pytest.mark.parametrize("op", ["gaussian_blur2d", "unsharp_mask"])
def test_gaussian_and_unsharp_integer_windows_on_marching_kernels(dev, case, op):
    """value_range windows with integer bounds on the marching kernels (windowed conversion in, windowed
    quantisation out): bit-exact against the oracle, float and integer output, every border the kernels take."""
    import mie_b200 as M
    import oracle as O

    dtype, vr = case
    for shape in [(2, 1, 128, 256), (1, 1, 64, 512), (3, 1, 192, 128)]:
        x = images("U", shape, dtype, seed=51)          # full dtype range: pixels on both sides of the window
        x01 = O.to01(x, vr)
        xt = gpu(x, dev)
        for border in ("reflect", "replicate", "constant"):
            ref = getattr(O, op)(x01, 9, 1.0, border)
            got = cpu(getattr(M, op)(xt, 9, 1.0, border, value_range=vr, out_dtype=torch.float32))
            assert np.array_equal(got, ref), (shape, border)
            assert np.array_equal(cpu(getattr(M, op)(xt, 9, 1.0, border, value_range=vr)), O.from01(ref, dtype, vr)), (shape, border)


def test_gaussian_against_independent_binaries(dev):
    """cv2.GaussianBlur(BORDER_REFLECT_101) and scipy.ndimage.gaussian_filter(mode='mirror') use the same
    weights and border; they agree with the kernel to fp32 rounding."""
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    import mie_b200 as M

    x = np.random.default_rng(3).random((200, 333), dtype=np.float32)
    got = cpu(M.gaussian_blur2d(gpu(x, dev), 9, 1.0))
    a = cv2.GaussianBlur(x, (9, 9), 1.0, borderType=cv2.BORDER_REFLECT_101)
    b = ndi.gaussian_filter(x, 1.0, mode="mirror", radius=4)
    assert np.abs(got - a).max() <= 5e-7
    assert np.abs(got - b).max() <= 5e-7


def test_pixel_mapping_exhaustive_uint16(dev):
    """All 65 536 uint16 codes: x01 = v/65535 matches IEEE division; quantise(normalise(v)) == v."""
    import mie_b200 as M

    v = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    xt = gpu(v, dev)
    f = cpu(M.gaussian_blur2d(xt, 1, 1.0, out_dtype=torch.float32))   # one tap of weight 1 = identity
    assert np.array_equal(f, v.astype(np.float32) / np.float32(65535.0))
    assert np.array_equal(cpu(M.gaussian_blur2d(xt, 1, 1.0)), v)
    s = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16).reshape(256, 256)
    assert np.array_equal(cpu(M.gaussian_blur2d(gpu(s, dev), 1, 1.0)), s)


# ---------------------------------------------------------------------------- fused chain
@pytest.mark.parametrize("kind", ["P", "U", "K"])
def test_chain_fused_bit_exact_and_equals_composition(dev, kind):
    import mie_b200 as M
    import oracle as O

    x = images(kind, (6, 1, 512, 512), np.uint16, seed=21)
    xt = gpu(x, dev)
    cfg = M.ChainConfig()
    got = cpu(M.enhance_chain(xt, cfg))
    ref, st = O.chain_gauss_clahe_unsharp(x, return_stages=True)
    assert np.array_equal(got, ref), int((got != ref).sum())
    # stage by stage on the GPU (fp32 intermediates) == fused
    g = M.gaussian_blur2d(xt, 9, 1.0, out_dtype=torch.float32)
    assert np.array_equal(cpu(g), st["gauss"])
    luts = M.clahe_luts(g, 2.0, (8, 8))
    assert np.array_equal(cpu(luts), st["luts"])        # "LUT entries differing" must be 0
    c = M.equalize_clahe(g, 2.0, (8, 8))
    assert np.array_equal(cpu(c), st["clahe"])
    u = M.unsharp_mask(c, 9, 1.0)
    assert np.array_equal(cpu(u), st["unsharp"])
    assert np.array_equal(O.from01(cpu(u), np.uint16), got)
    # float output of the fused path
    gf = cpu(M.enhance_chain(xt, cfg, out_dtype=torch.float32))
    assert np.array_equal(gf, st["unsharp"])


@pytest.mark.parametrize("case", [
    dict(shape=(2, 1, 500, 300), grid=(8, 8)),                      # CLAHE padding -> unfused path inside the library
    dict(shape=(1, 1, 256, 256), grid=(4, 4), dk=5, sk=7),          # 64-px tiles, other radii (fused)
    dict(shape=(1, 1, 128, 192), grid=(4, 4), dk=3, sk=3),          # 32x48 tiles (fused, partial stencil tiles)
    dict(shape=(1, 1, 512, 512), grid=(2, 2)),                      # 256-px tiles -> unfused
    dict(shape=(1, 1, 96, 96), grid=(2, 2), dk=11),                 # 11 taps -> unfused
    dict(shape=(1, 1, 64, 64), grid=(8, 8), border="replicate"),    # 8-px tiles: many LUTs per block
    # circular halos wrap to the opposite edge: the first / last blocks need LUTs of BOTH ends of the grid
    dict(shape=(2, 1, 512, 512), grid=(8, 8), border="circular"),
    dict(shape=(1, 1, 128, 192), grid=(4, 4), dk=5, sk=7, border="circular"),
    dict(shape=(1, 1, 64, 64), grid=(8, 8), border="circular"),     # 8-px tiles, 64 LUTs staged per block
    dict(shape=(1, 1, 256, 256), grid=(4, 4), border="constant"),
])
@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8, np.float32])
def test_chain_geometries(dev, case, dtype):
    import mie_b200 as M
    import oracle as O

    x = images("P", case["shape"], dtype, seed=4)
    dk, sk = case.get("dk", 9), case.get("sk", 9)
    border = case.get("border", "reflect")
    cfg = M.ChainConfig(denoise_kernel_size=dk, sharpen_kernel_size=sk, grid_size=case["grid"], border_type=border)
    got = cpu(M.enhance_chain(gpu(x, dev), cfg))
    ref = O.chain_gauss_clahe_unsharp(x, dk, 1.0, 2.0, case["grid"], sk, 1.0, border)
    assert np.array_equal(got, ref), int((got != ref).sum())


@pytest.mark.parametrize("case", [
    dict(shape=(2, 1, 128, 128), grid=(2, 2)),                         # one warp per band, both image edges in it
    dict(shape=(1, 1, 256, 1024), grid=(4, 16)),                       # 256 threads per band
    dict(shape=(3, 1, 64, 256), grid=(1, 4)),                          # single band: top and bottom mirrored
    dict(shape=(2, 1, 192, 384), grid=(3, 6)),
    dict(shape=(2, 1, 128, 256), grid=(2, 4), border="replicate"),
    dict(shape=(2, 1, 128, 256), grid=(2, 4), border="constant"),
    dict(shape=(1, 1, 320, 512), grid=(5, 8), clip=0.0),
])
@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8, np.float32])
def test_chain_marching_kernels(dev, case, dtype):
    """Geometries served by the marching kernels (64-px tiles, 9 taps, W % 128 == 0, W <= 1024):
    bit-exact against the oracle for every dtype and border mode, integer and float output."""
    import mie_b200 as M
    import oracle as O

    assert M._lib().mie_chain_is_fused(case["shape"][-2], case["shape"][-1], *case["grid"], 9, 9, 9, 9) == 2
    for kind in ("P", "U", "K"):
        x = images(kind, case["shape"], dtype, seed=7)
        border, clip = case.get("border", "reflect"), case.get("clip", 2.0)
        cfg = M.ChainConfig(grid_size=case["grid"], border_type=border, clip_limit=clip)
        ref, st = O.chain_gauss_clahe_unsharp(x, 9, 1.0, clip, case["grid"], 9, 1.0, border, return_stages=True)
        # stages = ALL | PREFER_MARCH: small batches take the tile kernels by default (latency heuristic)
        got = cpu(M.enhance_chain(gpu(x, dev), cfg, stages=3 | 4))
        assert np.array_equal(got, ref), (kind, int((got != ref).sum()))
        assert np.array_equal(cpu(M.enhance_chain(gpu(x, dev), cfg)), ref)               # default schedule
        assert np.array_equal(cpu(M.enhance_chain(gpu(x, dev), cfg, stages=3 | 8)), ref)  # tile kernels
        gf = cpu(M.enhance_chain(gpu(x, dev), cfg, out_dtype=torch.float32, stages=3 | 4))
        assert np.array_equal(gf, st["unsharp"]), kind


def test_chain_plan_graph_replay(dev):
    """ChainPlan (CUDA-graph replay on fixed buffers) == enhance_chain, and follows new input contents."""
    import mie_b200 as M

    x = gpu(images("P", (5, 1, 512, 512), np.uint16, seed=1), dev)
    plan = M.ChainPlan(x)
    a = cpu(plan.replay()).copy()
    assert np.array_equal(a, cpu(M.enhance_chain(x)))
    x.copy_(gpu(images("U", (5, 1, 512, 512), np.uint16, seed=2), dev))
    b = cpu(plan.replay())
    assert np.array_equal(b, cpu(M.enhance_chain(x)))
    assert not np.array_equal(a, b)


@pytest.mark.parametrize("case", [(np.int16, (-1024.0, 3071.0)), (np.uint16, (0.0, 4095.0)), (np.uint8, (10.0, 200.0))])
def test_chain_integer_windows_on_marching_kernels(dev, case):
    """value_range windows with integer bounds (HU window, 12-bit data) on the marching chain kernels: the windowed
    conversion on the way in, range-checked bins / lookups, windowed quantisation on the way out — bit-exact against
    the oracle for every border, on the marching kernels and on the tile kernels (the default for small jobs)."""
    import mie_b200 as M
    import oracle as O

    dtype, vr = case
    for shape, grid in [((2, 1, 128, 256), (2, 4)), ((3, 1, 192, 384), (3, 6)), ((1, 1, 512, 512), (8, 8))]:
        x = images("U", shape, dtype, seed=61)            # full dtype range: pixels on both sides of the window
        x[..., : shape[-2] // 2, :] = images("P", shape, dtype, seed=62)[..., : shape[-2] // 2, :]
        xt = gpu(x, dev)
        for border in ("reflect", "replicate", "constant"):
            cfg = M.ChainConfig(grid_size=grid, border_type=border, value_range=vr)
            ref = O.chain_gauss_clahe_unsharp(x, 9, 1.0, 2.0, grid, 9, 1.0, border, value_range=vr)
            got = cpu(M.enhance_chain(xt, cfg, stages=3 | 4))
            assert np.array_equal(got, ref), (shape, border, int((got != ref).sum()))
            assert np.array_equal(cpu(M.enhance_chain(xt, cfg)), ref), (shape, border)            # default: tile kernels
            assert np.array_equal(cpu(M.enhance_chain(xt, cfg, stages=3 | 8)), ref), (shape, border)
        cfg = M.ChainConfig(grid_size=grid, value_range=vr)
        reff = O.chain_gauss_clahe_unsharp(x, 9, 1.0, 2.0, grid, 9, 1.0, "reflect", value_range=vr, out_dtype=np.float32)
        assert np.array_equal(cpu(M.enhance_chain(xt, cfg, out_dtype=torch.float32, stages=3 | 4)), reff), shape
    # a full-size windowed batch takes the marching kernels by itself
    x = images("P", (32, 1, 512, 512), dtype, seed=63)
    cfg = M.ChainConfig(value_range=vr)
    ref = O.chain_gauss_clahe_unsharp(x, value_range=vr)
    assert np.array_equal(cpu(M.enhance_chain(gpu(x, dev), cfg)), ref)


def test_chain_ring_round_robin(dev):
    """ChainRing: independent batches replayed round-robin on their own streams give the same bits as
    direct calls, for one, two and three slots, also when the inputs are refilled between rounds."""
    import mie_b200 as M

    xs = [gpu(images("P", (6, 1, 512, 512), np.uint16, seed=40 + i), dev) for i in range(3)]
    refs = [M.enhance_chain(x) for x in xs]
    for nslots in (1, 2, 3):
        ring = M.ChainRing(xs[:nslots])
        ring.begin()
        for k in range(7):
            ring.replay(k)
        ring.join()
        torch.cuda.synchronize()
        for i in range(nslots):
            assert torch.equal(ring.plans[i].out, refs[i]), (nslots, i)
    ring = M.ChainRing([xs[0].clone(), xs[1].clone()])
    ring.plans[0].input.copy_(xs[2])      # refill slot 0 on the caller's stream, then replay
    ring.begin()
    out0 = ring.replay(0)
    ring.join()
    torch.cuda.synchronize()
    assert torch.equal(out0, refs[2])


def test_chain_full_config2_batch(dev):
    """BASELINE.json config 2 at full size (256 x 512 x 512 uint16): bit-exact against the oracle, and
    batch-independent (a checksum of per-slice checksums equals the one from slice-at-a-time calls)."""
    import mie_b200 as M
    import oracle as O

    x = images("P", (256, 1, 512, 512), np.uint16, seed=0)
    xt = gpu(x, dev)
    got = cpu(M.enhance_chain(xt))
    ref = O.chain_gauss_clahe_unsharp(x)
    assert np.array_equal(got, ref), int((got != ref).sum())
    whole = hashlib.sha256(b"".join(hashlib.sha256(got[i].tobytes()).digest() for i in range(256))).hexdigest()
    parts = []
    for i in range(0, 256, 37):  # ragged chunks
        parts.append(cpu(M.enhance_chain(xt[i:i + 37])))
    chunks = np.concatenate(parts)
    again = hashlib.sha256(b"".join(hashlib.sha256(chunks[i].tobytes()).digest() for i in range(256))).hexdigest()
    assert whole == again


def test_empty_batch(dev):
    import mie_b200 as M

    x = torch.empty((0, 1, 64, 64), dtype=torch.uint16, device=dev)
    assert M.enhance_chain(x).shape == (0, 1, 64, 64)
    assert M.gaussian_blur2d(x, 9, 1.0).shape == (0, 1, 64, 64)
