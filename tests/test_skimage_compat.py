"""scikit-image call signatures: skimage.filters.gaussian / unsharp_mask delegate to
scipy.ndimage, which IS installed — so the oracle restatement is pinned against the real thing on the CPU,
and the CUDA path is compared with the oracle (bit-exact: same fp32 operation order) on the GPU."""
import numpy as np
import pytest
import torch

MODES = ["nearest", "reflect", "mirror", "constant", "wrap"]


def test_oracle_skimage_gaussian_matches_scipy_ndimage():
    ndi = pytest.importorskip("scipy.ndimage")
    import oracle as O

    x = np.random.default_rng(0).random((2, 37, 41), dtype=np.float32)
    for mode in MODES:
        for sigma in (0.7, 1.0, 2.3):
            ref = np.stack([ndi.gaussian_filter(p.astype(np.float64), sigma, mode=mode) for p in x])
            assert np.abs(O.skimage_gaussian(x, sigma, mode) - ref).max() < 5e-7, (mode, sigma)
    # symmetric = scipy 'reflect' also on images smaller than the kernel radius
    tiny = np.random.default_rng(1).random((1, 3, 5), dtype=np.float32)
    ref = ndi.gaussian_filter(tiny[0].astype(np.float64), 2.0, mode="reflect")
    assert np.abs(O.skimage_gaussian(tiny, 2.0, "reflect")[0] - ref).max() < 5e-7


def test_oracle_skimage_unsharp_matches_formula_on_scipy():
    ndi = pytest.importorskip("scipy.ndimage")
    import oracle as O

    x = np.random.default_rng(2).random((2, 40, 33), dtype=np.float32)
    for radius, amount in [(1.0, 1.0), (1.5, 0.8), (2.5, 2.0), (0.5, -0.5)]:
        raw = np.stack([p + amount * (p - ndi.gaussian_filter(p.astype(np.float64), radius, mode="reflect")) for p in x])
        assert np.abs(O.skimage_unsharp_mask(x, radius, amount, preserve_range=True) - raw).max() < 1e-6
        assert np.abs(O.skimage_unsharp_mask(x, radius, amount) - raw.clip(0, 1)).max() < 1e-6
    # amount == 1 on the kornia border is kornia's unsharp_mask, bit for bit (fma(1, d, x) == x + d)
    k = O.unsharp_mask(x, 9, 1.0, "symmetric")
    assert np.array_equal(O.skimage_unsharp_mask(x, 1.0, 1.0, preserve_range=True), k)


def _rand(dtype, shape, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max + 1, shape).astype(dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_gaussian_skimage_signature(dev, dtype):
    import oracle as O
    from mie_b200 import skimage_compat as S

    for shape in [(2, 1, 70, 50), (1, 1, 64, 128), (5, 9)]:
        x = _rand(dtype, shape, 3)
        xt = torch.from_numpy(x).to(dev)
        for mode in MODES:
            for sigma in (1, 2.3, (0.8, 1.6)):
                ky, kx = (2 * int(4.0 * s + 0.5) + 1 for s in (sigma if isinstance(sigma, tuple) else (sigma, sigma)))
                if mode == "mirror" and (ky // 2 >= shape[-2] or kx // 2 >= shape[-1]):
                    continue  # kornia-style reflect needs halo < image
                if mode == "wrap" and (ky // 2 > shape[-2] or kx // 2 > shape[-1]):
                    continue
                ref = O.gaussian_blur2d(O.to01(x), (ky, kx), sigma if isinstance(sigma, tuple) else (sigma, sigma),
                                        O.SCIPY_MODES[mode])
                got = S.gaussian(xt, sigma, mode=mode, out_dtype=torch.float32).cpu().numpy()
                assert np.array_equal(got, ref), (shape, mode, sigma)
                if dtype != np.float32:
                    assert np.array_equal(S.gaussian(xt, sigma, mode=mode).cpu().numpy(), O.from01(ref, dtype))
    with pytest.raises(ValueError):
        S.gaussian(xt, 1.0, mode="bogus")
    with pytest.raises(ValueError):
        S.gaussian(xt, 10.0)  # 81 taps


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_unsharp_mask_skimage_signature(dev, dtype):
    import mie_b200 as M
    import oracle as O
    from mie_b200 import skimage_compat as S

    for shape in [(2, 1, 70, 50), (1, 1, 128, 256), (1, 1, 512, 512)]:
        x = _rand(dtype, shape, 4)
        xt = torch.from_numpy(x).to(dev)
        for radius, amount, keep in [(1.0, 1.0, False), (1.5, 0.8, False), (2.5, 2.0, True), (0.6, -0.5, True)]:
            ref = O.skimage_unsharp_mask(O.to01(x), radius, amount, preserve_range=keep)
            got = S.unsharp_mask(xt, radius, amount, preserve_range=keep, out_dtype=torch.float32).cpu().numpy()
            assert np.array_equal(got, ref), (shape, radius, amount)
            if dtype != np.float32:
                assert np.array_equal(S.unsharp_mask(xt, radius, amount, preserve_range=keep).cpu().numpy(),
                                      O.from01(ref, dtype))
    # amount = 1, no clipping, kornia border: the tuned kornia kernel gives the same bits
    x = _rand(np.uint16, (2, 1, 128, 256), 5)
    xt = torch.from_numpy(x).to(dev)
    a = M.unsharp_mask(xt, 9, 1.0, "symmetric", out_dtype=torch.float32)
    b = S.unsharp_mask(xt, 1.0, 1.0, preserve_range=True, out_dtype=torch.float32)
    assert torch.equal(a, b)
