"""GPU parity tests for median (2-D / 3-D), bilateral and global equalisation, and the committed
golden vectors.  Integer / selection work is bit-exact; bilateral follows the oracle's fp32
operation order (including its fixed-sequence exp) and is asserted bit-exact as well, with the
north star's rel 1e-5 checked against the torch-CPU kornia twin (true exp)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gpu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def cpu(t):
    return t.cpu().numpy()


def rand(dtype, shape, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max + 1, shape).astype(dtype)


# ---------------------------------------------------------------------------- median 2-D
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
@pytest.mark.parametrize("k", [3, 5, 7, (3, 5), (5, 3), (1, 3), 1])
def test_median_blur_bit_exact(dev, dtype, k):
    import mie_b200 as M
    import oracle as O

    for shape in [(2, 1, 70, 50), (1, 2, 33, 129), (1, 1, 5, 4)]:
        x = rand(dtype, shape, 3)
        if dtype != np.float32:
            x[..., : shape[-2] // 2, :] //= 16  # many ties
        for border in ("constant", "replicate"):
            got = cpu(M.median_blur(gpu(x, dev), k, border_type=border))
            assert np.array_equal(got, O.median_blur(x, k, border)), (shape, k, border)


@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8, np.float32])
def test_median_blur_3x3_packed_kernel(dev, dtype):
    """16-bit 3x3 on 16-byte-aligned rows takes the marching packed kernel (two pixels per lane,
    neighbours by shuffle): partial warps, several 256-column strips, short bands, every border rule."""
    import mie_b200 as M
    import oracle as O

    for shape in [(2, 1, 70, 48), (1, 1, 33, 512), (1, 2, 9, 264), (1, 1, 2, 8), (3, 1, 100, 1032), (1, 1, 67, 256),
                  (1, 1, 4, 8), (1, 1, 5, 16)]:
        x = rand(dtype, shape, 5)
        if dtype == np.float32:
            x[..., : shape[-2] // 2, :] = np.round(x[..., : shape[-2] // 2, :] * 8) / 8  # many ties
        else:
            x[..., : shape[-2] // 2, :] //= (64 if dtype != np.uint8 else 8)  # many ties
        for k in (3, 5):   # both packed marching kernels
            for border in ("constant", "replicate", "reflect", "symmetric"):
                if border == "reflect" and k // 2 >= min(shape[-2:]):
                    continue
                got = cpu(M.median_blur(gpu(x, dev), k, border_type=border))
                assert np.array_equal(got, O.median_blur(x, k, border)), (shape, k, border)
    # config-2 sized batch: first / last slices against the oracle, all slices against scipy on a sample
    from mie_b200 import synthetic

    x = synthetic.phantom((256, 1, 512, 512), np.uint16 if dtype != np.uint8 else np.uint8, seed=2).astype(dtype)
    if dtype == np.float32:
        x = (x / np.float32(4095.0)).astype(np.float32)
    got = cpu(M.median_blur(gpu(x, dev), 3))
    for i in (0, 1, 127, 255):
        assert np.array_equal(got[i], O.median_blur(x[i:i + 1], 3, "constant")[0]), i
    got5 = cpu(M.median_blur(gpu(x, dev), 5))
    for i in (0, 200):
        assert np.array_equal(got5[i], O.median_blur(x[i:i + 1], 5, "constant")[0]), i


def test_median_blur_matches_scipy_and_cv2(dev):
    cv2 = pytest.importorskip("cv2")
    ndi = pytest.importorskip("scipy.ndimage")
    import mie_b200 as M

    x = rand(np.uint16, (300, 400), 4)
    xt = gpu(x, dev)
    assert np.array_equal(cpu(M.median_blur(xt, 3)), ndi.median_filter(x, size=3, mode="constant", cval=0))
    assert np.array_equal(cpu(M.median_blur(xt, 5, border_type="replicate")), cv2.medianBlur(x, 5))
    assert np.array_equal(cpu(M.median(xt)), ndi.median_filter(x, size=3, mode="nearest"))  # skimage default


def test_median_unsupported_kernel(dev):
    import mie_b200 as M

    with pytest.raises(ValueError):
        M.median_blur(gpu(np.zeros((8, 8), np.uint16), dev), 9)
    with pytest.raises(ValueError):
        M.median_blur(gpu(np.zeros((8, 8), np.uint16), dev), 4)


# ---------------------------------------------------------------------------- median 3-D
@pytest.mark.parametrize("dtype", [np.int16, np.uint16, np.uint8, np.float32])
def test_median3d_bit_exact_against_oracle_and_scipy(dev, dtype):
    ndi = pytest.importorskip("scipy.ndimage")
    import mie_b200 as M
    import oracle as O

    for shape in [(12, 20, 24), (70, 33, 65), (1, 9, 9), (3, 8, 40), (20, 21, 130), (9, 17, 66), (2, 1, 2)]:
        vol = rand(dtype, shape, 6)
        for mode in ("nearest", "constant"):
            got = cpu(M.median(gpu(vol, dev), mode=mode))
            assert np.array_equal(got, O.median3d(vol, mode)), (shape, mode)
            if mode == "nearest":
                assert np.array_equal(got, ndi.median_filter(vol, size=3, mode="nearest"))


def test_median3d_slab_halos(dev):
    """z-slabs with neighbour halo planes reproduce the unsharded volume bit for bit."""
    import mie_b200 as M
    from mie_b200 import synthetic

    vol = synthetic.phantom_volume((40, 64, 96), np.int16, seed=1)
    vt = gpu(vol, dev)
    full = cpu(M.median(vt))
    for nslab in (2, 4, 5):
        bounds = np.linspace(0, 40, nslab + 1).astype(int)
        parts = []
        for i in range(nslab):
            lo, hi = bounds[i], bounds[i + 1]
            halo_lo = vt[lo - 1].contiguous() if lo > 0 else None
            halo_hi = vt[hi].contiguous() if hi < 40 else None
            parts.append(cpu(M.median(vt[lo:hi], halo_lo=halo_lo, halo_hi=halo_hi)))
        assert np.array_equal(np.concatenate(parts), full), nslab


# ---------------------------------------------------------------------------- bilateral
@pytest.mark.parametrize("dtype", [np.float32, np.uint16, np.uint8])
def test_bilateral_bit_exact_and_within_tolerance_of_true_exp(dev, dtype):
    import kornia_twin as K
    import mie_b200 as M
    import oracle as O

    for shape, k, sc, ss, border in [((2, 1, 70, 50), 9, 0.1, 1.5, "reflect"), ((1, 1, 33, 65), (5, 7), 0.2, (1.5, 1.2), "replicate"),
                                     ((1, 1, 40, 40), 3, 0.05, 0.8, "constant"), ((1, 1, 64, 64), 15, 0.3, 3.0, "reflect"),
                                     ((1, 1, 45, 67), 5, 0.15, 1.0, "circular"), ((2, 1, 33, 31), 7, 0.08, 2.0, "replicate")]:
        x = rand(dtype, shape, 8)
        x01 = O.to01(x)
        ref = O.bilateral_blur(x01, k, sc, ss, border)
        twin = K.bilateral_blur(torch.from_numpy(x01), k, sc, ss, border).numpy()
        # reproducible mode (polynomial 2^t, the oracle's operation order): bit for bit
        with M.kernel_policy("bilateral_exact_exp"):
            got = cpu(M.bilateral_blur(gpu(x, dev), k, sc, ss, border, out_dtype=torch.float32))
            assert np.array_equal(got, ref), float(np.abs(got - ref).max())
            assert (np.abs(got - twin) / np.maximum(np.abs(twin), 1e-6)).max() <= 1e-5
            if dtype != np.float32:
                q = cpu(M.bilateral_blur(gpu(x, dev), k, sc, ss, border))
                assert np.array_equal(q, O.from01(ref, dtype))
        # default mode (MUFU.EX2 colour weights on the square 3..9 windows): the north star's tolerance for floating-point
        # filters, rel 1e-5 before quantisation and 1 LSB after
        fast = cpu(M.bilateral_blur(gpu(x, dev), k, sc, ss, border, out_dtype=torch.float32))
        assert (np.abs(fast - ref) / np.maximum(np.abs(ref), 1e-3)).max() <= 1e-5
        assert (np.abs(fast - twin) / np.maximum(np.abs(twin), 1e-3)).max() <= 1e-5
        if dtype != np.float32:
            q = cpu(M.bilateral_blur(gpu(x, dev), k, sc, ss, border)).astype(np.int64)
            assert np.abs(q - O.from01(ref, dtype).astype(np.int64)).max() <= 1


# ---------------------------------------------------------------------------- equalisation: the three dispatch paths
@pytest.mark.parametrize("shape,dtype", [((3, 1, 64, 128), np.uint16), ((2, 1, 200, 256), np.int16), ((5, 1, 512, 512), np.uint16),
                                         ((1, 1, 100, 64), np.uint8), ((1, 1, 1024, 1024), np.uint16), ((2, 1, 37, 512), np.uint16),
                                         ((1, 1, 9, 8), np.uint16), ((2, 1, 640, 2048), np.uint16)])
def test_equalize_cluster_paths_agree_with_three_pass_and_oracle(dev, shape, dtype):
    """One-launch cluster kernel without a slab (small planes: second read through L2), with a shared-memory slab (TMA;
    larger planes, or forced by policy), and the three-pass path: same bits, and the oracle's."""
    import mie_b200 as M
    import oracle as O

    x = rand(dtype, shape, 21)
    x[0, 0, : shape[2] // 2] //= 9            # a skewed histogram in the first plane
    xt = gpu(x, dev)
    a = cpu(M.equalize(xt))
    with M.kernel_policy("equalize_slab"):
        b = cpu(M.equalize(xt))
    with M.kernel_policy("equalize_three_pass"):
        c = cpu(M.equalize(xt))
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert np.array_equal(a, O.from01(O.equalize(O.to01(x)), dtype))
    k = gpu(np.full(shape, 77, dtype), dev)   # step == 0: v / 255 goes back unchanged
    assert np.array_equal(cpu(M.equalize(k)), np.full(shape, 77, dtype))


# ---------------------------------------------------------------------------- global equalisation
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_equalize_bit_exact(dev, dtype):
    import mie_b200 as M
    import oracle as O

    for shape in [(3, 1, 64, 80), (1, 2, 300, 500), (1, 1, 7, 5)]:
        x = rand(dtype, shape, 10)
        ref = O.equalize(O.to01(x))
        got = cpu(M.equalize(gpu(x, dev), out_dtype=torch.float32))
        assert np.array_equal(got, ref)
        if dtype != np.float32:
            assert np.array_equal(cpu(M.equalize(gpu(x, dev))), O.from01(ref, dtype))
    if dtype == np.float32:   # values outside [0, 1] and NaN: ignored by the histogram, clamped by the lookup
        x = rand(dtype, (2, 1, 64, 80), 13) * 1.6 - 0.3
        x[0, 0, 5, 7:20] = np.nan
        ref = O.equalize(x)
        got = cpu(M.equalize(gpu(x, dev)))
        assert np.array_equal(got, ref, equal_nan=True)
    # constant image: step == 0 -> returned unchanged
    c = np.full((1, 1, 32, 32), 77, np.uint8)
    assert np.array_equal(cpu(M.equalize(gpu(c, dev))), c)


@pytest.mark.parametrize("case", [(np.int16, (-1024.0, 3071.0)), (np.uint16, (0.0, 4095.0)), (np.uint8, (10.0, 200.0))])
def test_equalize_integer_windows(dev, case):
    """value_range windows on the tuned kernels (windowed conversion, csrc/window.cuh): pixels outside the
    window are ignored by the histogram and clamped by the lookup."""
    import mie_b200 as M
    import oracle as O

    dtype, vr = case
    for shape in [(3, 1, 64, 80), (2, 1, 512, 512)]:
        x = rand(dtype, shape, 12)
        ref = O.equalize(O.to01(x, vr))
        assert np.array_equal(cpu(M.equalize(gpu(x, dev), value_range=vr, out_dtype=torch.float32)), ref)
        assert np.array_equal(cpu(M.equalize(gpu(x, dev), value_range=vr)), O.from01(ref, dtype, vr))


def test_equalize_matches_torchvision_uint8(dev):
    tvf = pytest.importorskip("torchvision.transforms.v2.functional")
    import mie_b200 as M

    x = rand(np.uint8, (2, 3, 120, 90), 11)
    ref = tvf.equalize(torch.from_numpy(x)).numpy()
    assert np.array_equal(cpu(M.equalize(gpu(x, dev))), ref)


# ---------------------------------------------------------------------------- golden vectors
def test_golden_hashes_on_gpu(dev):
    import make_golden as G
    import mie_b200 as M

    with open(os.path.join(GOLDEN, "golden.json")) as f:
        want = json.load(f)
    x = G.inputs()
    p, f32, u8, u16 = gpu(x["phantom_u16"], dev), gpu(x["noise_f32"], dev), gpu(x["noise_u8"], dev), gpu(x["uniform_u16"], dev)
    with M.kernel_policy("bilateral_exact_exp"):   # the committed digest is the reproducible mode's
        bil = G.sha(cpu(M.bilateral_blur(f32, 9, 0.1, 1.5)))
    got = {
        "chain_c2_phantom_u16": G.sha(cpu(M.enhance_chain(p))),
        "chain_c2_uniform_u16": G.sha(cpu(M.enhance_chain(u16))),
        "clahe_luts_phantom_u16_8x8_clip2": G.sha(cpu(M.clahe_luts(p, 2.0, (8, 8)))),
        "clahe_phantom_u16_8x8_clip2": G.sha(cpu(M.equalize_clahe(p, 2.0, (8, 8)))),
        "clahe_f32_4x6_clip2": G.sha(cpu(M.equalize_clahe(f32, 2.0, (4, 6)))),
        "gauss_f32_k9_s1_reflect": G.sha(cpu(M.gaussian_blur2d(f32, 9, 1.0))),
        "gauss_f32_k5_s1.2_replicate": G.sha(cpu(M.gaussian_blur2d(f32, 5, 1.2, "replicate"))),
        "unsharp_f32_k9_s1_reflect": G.sha(cpu(M.unsharp_mask(f32, 9, 1.0))),
        "opencv_clahe_u8_8x8_clip2": G.sha(cpu(M.equalize_clahe(u8, 2.0, (8, 8), semantics="opencv"))),
        "median3x3_u16_zero": G.sha(cpu(M.median_blur(u16, 3))),
        "median5x5_u16_zero": G.sha(cpu(M.median_blur(u16, 5))),
        "median3d_i16_nearest": G.sha(cpu(M.median(gpu(x["phantom_i16_vol"], dev)))),
        "bilateral_f32_k9_sc0.1_ss1.5": bil,
        "equalize_u8": G.sha(cpu(M.equalize(u8))),
        "equalize_f32": G.sha(cpu(M.equalize(f32))),
    }
    assert got == want


def test_committed_fixture(dev):
    import mie_b200 as M

    fx = np.load(os.path.join(GOLDEN, "chain_128_grid2.npz"))
    cfg = M.ChainConfig(grid_size=(2, 2))
    assert np.array_equal(cpu(M.enhance_chain(gpu(fx["input"], dev), cfg)), fx["output"])


# ---------------------------------------------------------------------------- batching loader
def test_host_pipeline_equals_direct_call(dev):
    import mie_b200 as M
    from mie_b200 import synthetic

    x = synthetic.phantom((70, 1, 256, 256), np.uint16, seed=12)       # ragged: 70 = 2 * 32 + 6
    cfg = M.ChainConfig(grid_size=(4, 4))
    ref = cpu(M.enhance_chain(gpu(x, dev), cfg))
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    pipe = M.HostSlicePipeline(dev, (256, 256), torch.uint16, chunk=32, config=cfg)
    for _ in range(4):                              # buffers are reused: runs 2+ capture / replay a CUDA graph
        yh.zero_()
        pipe.run(xh, yh)
        assert np.array_equal(yh.numpy(), ref)
    assert pipe._graph is not None
    x2 = synthetic.phantom((70, 1, 256, 256), np.uint16, seed=13)      # new contents, same staging buffers
    xh.copy_(torch.from_numpy(x2))
    pipe.run(xh, yh)
    assert np.array_equal(yh.numpy(), cpu(M.enhance_chain(gpu(x2, dev), cfg)))
    assert np.array_equal(M.enhance_chain_host(torch.from_numpy(x), cfg, device=dev).numpy(), ref)


def test_host_volume_pipeline_equals_device_volume(dev):
    """Streaming a host volume in z-chunks (with the neighbouring planes as median halos) is bit-identical to
    processing the whole volume on the device; ragged last chunk, chunk of one plane, pageable input."""
    import mie_b200 as M
    from mie_b200 import synthetic

    vol = synthetic.phantom_volume((45, 128, 192), np.int16, seed=9)
    ref = cpu(M.median3d_clahe_slab(gpu(vol, dev), 2.0, (2, 3)))
    host = torch.from_numpy(vol)
    for chunk in (16, 7, 1, 64):
        got = M.median3d_clahe_host(host, 2.0, (2, 3), device=dev, chunk=chunk)
        assert np.array_equal(got.numpy(), ref), chunk
    out = torch.empty_like(host).pin_memory()
    pipe = M.HostVolumePipeline(dev, (128, 192), torch.int16, chunk=16, grid_size=(2, 3),
                                value_range=(-1024.0, 3071.0))
    pipe.run(host.pin_memory(), out)
    ref2 = cpu(M.median3d_clahe_slab(gpu(vol, dev), 2.0, (2, 3), value_range=(-1024.0, 3071.0)))
    assert np.array_equal(out.numpy(), ref2)
    with pytest.raises(ValueError):
        M.median3d_clahe_host(gpu(vol, dev))


# ---------------------------------------------------------------------------- non-local means (config 5)
@pytest.mark.parametrize("case", [
    dict(shape=(2, 1, 64, 80), ps=7, pd=11, h=0.1, sigma=0.0),       # skimage defaults on a ragged tile grid
    dict(shape=(1, 1, 96, 96), ps=7, pd=5, h=0.08, sigma=0.04),
    dict(shape=(1, 2, 40, 50), ps=5, pd=3, h=0.2, sigma=0.0),
    dict(shape=(1, 1, 33, 47), ps=3, pd=2, h=0.05, sigma=0.01),
    dict(shape=(1, 1, 64, 64), ps=8, pd=4, h=0.1, sigma=0.0),        # even patch size -> 9
])
def test_nlm_within_tolerance_of_float64_oracle(dev, case):
    """fp32 kernel (ex2.approx weights) vs the float64 oracle: rel/abs 1e-5 on [0,1] data (the north
    star's bar for floating-point filters) and <= 1 LSB after quantisation."""
    import mie_b200 as M
    import oracle as O
    from mie_b200 import synthetic

    x = synthetic.phantom(case["shape"], np.uint16, seed=5)
    x = (x.astype(np.float64) * (65535.0 / 4095.0)).clip(0, 65535).astype(np.uint16)    # use the full range
    kw = dict(patch_size=case["ps"], patch_distance=case["pd"], h=case["h"], sigma=case["sigma"])
    ref = O.denoise_nl_means(O.to01(x).astype(np.float64), case["ps"], case["pd"], case["h"], case["sigma"])
    gf = cpu(M.denoise_nl_means(gpu(x, dev), out_dtype=torch.float32, **kw)).astype(np.float64)
    assert np.abs(gf - ref).max() <= 1e-5, float(np.abs(gf - ref).max())
    gq = cpu(M.denoise_nl_means(gpu(x, dev), **kw)).astype(np.int64)
    rq = np.rint(np.clip(ref, 0, 1) * 65535.0).astype(np.int64)
    assert np.abs(gq - rq).max() <= 1
    xf = O.to01(x)
    gff = cpu(M.denoise_nl_means(gpu(xf, dev), **kw)).astype(np.float64)
    assert np.abs(gff - ref).max() <= 1e-5


@pytest.mark.parametrize("case", [
    dict(shape=(2, 1, 40, 52), ps=7, pd=11, h=0.1, sigma=0.0),       # skimage defaults; windows clipped on every side
    dict(shape=(1, 1, 48, 48), ps=5, pd=6, h=0.08, sigma=0.04),
    dict(shape=(1, 2, 33, 21), ps=3, pd=2, h=0.05, sigma=0.01),      # ragged 16x16 tile grid
    dict(shape=(1, 1, 32, 40), ps=8, pd=4, h=0.3, sigma=0.0),        # even patch size -> 9; large h: few cut-offs
])
def test_nlm_slow_mode_matches_float64_oracle(dev, case):
    """fast_mode=False: float64 kernel in upstream's operation order (Gaussian patch weights, clipped search window,
    cut-off test before every patch row) vs the float64 oracle — only exp() may differ, by an ulp."""
    import mie_b200 as M
    import oracle as O
    from mie_b200 import synthetic

    x = synthetic.phantom(case["shape"], np.uint16, seed=7)
    x = (x.astype(np.float64) * (65535.0 / 4095.0)).clip(0, 65535).astype(np.uint16)
    kw = dict(patch_size=case["ps"], patch_distance=case["pd"], h=case["h"], sigma=case["sigma"], fast_mode=False)
    ref = O.denoise_nl_means_slow(O.to01(x).astype(np.float64), case["ps"], case["pd"], case["h"], case["sigma"])
    g64 = cpu(M.denoise_nl_means(gpu(x, dev), out_dtype=torch.float64, **kw))
    assert g64.dtype == np.float64 and np.abs(g64 - ref).max() <= 1e-13, float(np.abs(g64 - ref).max())
    g32 = cpu(M.denoise_nl_means(gpu(x, dev), out_dtype=torch.float32, **kw))
    assert np.array_equal(g32, ref.astype(np.float32)) or np.abs(g32 - ref).max() <= 6e-8
    gq = cpu(M.denoise_nl_means(gpu(x, dev), **kw)).astype(np.int64)
    rq = np.rint(np.clip(ref.astype(np.float32), 0, 1) * np.float32(65535.0)).astype(np.int64)
    assert np.abs(gq - rq).max() <= 1 and (gq != rq).mean() < 1e-3
    xf = O.to01(x)
    gff = cpu(M.denoise_nl_means(gpu(xf, dev), **kw))
    assert gff.dtype == np.float32 and np.abs(gff - ref).max() <= 6e-8
    # the two modes are different filters: make sure the slow one is not silently the fast one
    fast = O.denoise_nl_means(O.to01(x).astype(np.float64), case["ps"], case["pd"], case["h"], case["sigma"])
    assert np.abs(fast - ref).max() > 1e-4


def test_nlm_argument_errors(dev):
    import mie_b200 as M

    x = torch.zeros((1, 1, 64, 64), dtype=torch.float32, device=dev)
    with pytest.raises(NotImplementedError):
        M.denoise_nl_means(x, channel_axis=-1)
    with pytest.raises((ValueError, RuntimeError)):
        M.denoise_nl_means(x, patch_size=11)
    with pytest.raises((ValueError, RuntimeError)):
        M.denoise_nl_means(x, patch_size=17, fast_mode=False)            # slow mode: patch size <= 15
    with pytest.raises(TypeError):
        M.denoise_nl_means(x.to(torch.float32), fast_mode=False, out_dtype=torch.int16)
    with pytest.raises((ValueError, RuntimeError)):
        M.denoise_nl_means(torch.zeros((1, 1, 12, 12), device=dev))      # padding would exceed the image


# ---------------------------------------------------------------------------- edge cases of the newer entry points
def test_edge_cases_metrics_windows_host_volume(dev):
    import mie_b200 as M
    from mie_b200 import skimage_compat as S

    # single-pixel-high / single-plane inputs
    a = gpu(rand(np.uint16, (1, 1, 1, 64), 1), dev)
    assert M.mse(a, a) == 0.0 and M.psnr(a, a) == float("inf")
    with pytest.raises(ValueError):
        M.ssim(a, a)                                     # 11x11 window does not fit a 1x64 image
    assert M.ssim(a, a, ws=1)[0] == pytest.approx(1.0, abs=1e-12)
    # value_range windows that the tuned kernels cannot take (non-integer bounds) still work (generic kernels)
    x = gpu(rand(np.int16, (2, 1, 64, 64), 2), dev)
    import oracle as O
    vr = (-1000.5, 2999.25)
    ref = O.from01(O.equalize_clahe(O.to01(cpu(x), vr), 2.0, (2, 2)), np.int16, vr)
    assert np.array_equal(cpu(M.equalize_clahe(x, 2.0, (2, 2), value_range=vr)), ref)
    ref = O.from01(O.equalize(O.to01(cpu(x), vr)), np.int16, vr)
    assert np.array_equal(cpu(M.equalize(x, value_range=vr)), ref)
    # host volume of a single plane, and of fewer planes than a chunk
    for d in (1, 2, 5):
        vol = rand(np.int16, (d, 64, 64), 3)
        got = M.median3d_clahe_host(torch.from_numpy(vol), 2.0, (2, 2), device=dev, chunk=4)
        assert np.array_equal(got.numpy(), cpu(M.median3d_clahe_slab(gpu(vol, dev), 2.0, (2, 2)))), d
    # skimage signatures on (H, W) input keep the rank
    img = gpu(rand(np.uint16, (40, 56), 4), dev)
    assert S.gaussian(img, 1.0).shape == (40, 56) and S.unsharp_mask(img, 1.0, 1.0).shape == (40, 56)
    # median(out=) must not alias the input
    with pytest.raises(ValueError):
        M.median(img, out=img)
