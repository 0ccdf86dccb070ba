"""scikit-image exposure / restoration wrappers: equalize_adapthist, equalize_hist,
denoise_bilateral.  RECALLED semantics (the package is not on disk; tests/test_live_pins.py activates when it is).

CPU: the array-level numpy twin (oracle/skimage_twin.py) and the per-pixel C restatement (oracle/mie_oracle.c
orc_sk_*) are structurally independent and must agree BIT FOR BIT, stage by stage.
GPU: the CUDA path through the C ABI must reproduce the oracle bit for bit (float64 outputs compared as bits)."""
import numpy as np
import pytest
import torch


def _phantom(shape, dtype, seed=0):
    from mie_b200 import synthetic

    return synthetic.phantom(shape, dtype, seed)


def _adapt_cases():
    rng = np.random.default_rng(0)
    P = _phantom((1, 1, 512, 512), np.uint16, 0)[0, 0]
    return [
        ("phantom 512 u16 default (C1 shape)", P, None, 0.01, 256),
        ("phantom 512 i16 default", (P.astype(np.int32) - 1024).astype(np.int16), None, 0.01, 256),
        ("phantom heavy clip", P, None, 0.003, 256),
        ("phantom light clip", P, None, 0.5, 256),
        ("uniform 100x130 ragged", rng.integers(0, 65536, (100, 130)).astype(np.uint16), None, 0.01, 256),
        ("uniform kernel (17, 23) nbins 128", rng.integers(0, 65536, (100, 130)).astype(np.uint16), (17, 23), 0.03, 128),
        ("uint8 64x61 kernel 8", rng.integers(0, 256, (64, 61)).astype(np.uint8), 8, 0.02, 256),
        ("constant image", np.full((64, 64), 1000, np.uint16), None, 0.01, 256),
        ("clip_limit 0 (AHE)", rng.integers(0, 4096, (96, 96)).astype(np.uint16), None, 0.0, 256),
        ("float32 image", rng.random((80, 72)).astype(np.float32), None, 0.01, 256),
        ("tiny 5x7 kernel (4, 3): reflect pad longer than a region", rng.integers(0, 65536, (5, 7)).astype(np.uint16), (4, 3), 0.5, 64),
        ("two grey levels", (rng.integers(0, 2, (48, 40)) * 3000).astype(np.uint16), None, 0.01, 256),
    ]


@pytest.mark.parametrize("case", _adapt_cases(), ids=lambda c: c[0])
def test_adapthist_twin_and_per_pixel_restatement_agree(case):
    import oracle as O
    import skimage_twin as S

    _, x, ks, cl, nb = case
    a, sa = S.equalize_adapthist(x, ks, cl, nb, return_stages=True)
    b, sb = O.sk_equalize_adapthist(x, ks, cl, nb, return_stages=True)
    assert np.array_equal(sa["grey"], sb["grey"])
    assert np.array_equal(sa["clahe"], sb["clahe"])
    assert a.dtype == b.dtype and np.array_equal(a, b)
    assert a.min() >= 0.0 and a.max() <= 1.0


def test_adapthist_semantics_known_answers():
    import skimage_twin as S

    # clip_histogram (hand-evaluated): excess 40 -> increment 5 per bin, `upper` = 5.  The mid mask is taken on the
    # UPDATED histogram (upstream's order), so bins raised into [upper, clim) are raised again to the limit:
    h = np.array([50, 0, 0, 10, 0, 0, 0, 4], np.int64)
    assert np.array_equal(S.clip_histogram(h, 10), np.full(8, 10))
    # 16 bins: excess 40 -> increment 2, upper 8; nothing lands in [8, 10); the remainder 40 - 13*2 = 14 > 0 is handed
    # out with stride max(1, under // excess) = 1 from bin 0: every bin under the limit gets one more count
    h = np.array([50, 0, 0, 3, 0, 0, 0, 4, 1, 1, 0, 0, 0, 0, 0, 5], np.int64)
    assert np.array_equal(S.clip_histogram(h, 10), np.array([10, 3, 3, 6, 3, 3, 3, 7, 4, 4, 3, 3, 3, 3, 3, 8]))
    # a constant image has one grey level -> every mapping equal -> flat output
    out = S.equalize_adapthist(np.full((32, 32), 7, np.uint8))
    assert np.all(out == out.flat[0])
    # kernel_size default is max(dim // 8, 1) in pixels
    assert S.adapthist_kernel_size((512, 300), None) == [64, 37]
    with pytest.raises(ValueError):
        S.adapthist_kernel_size((512, 300), (8, 8, 8))


def test_equalize_hist_twin_and_restatement_agree():
    import oracle as O
    import skimage_twin as S

    rng = np.random.default_rng(1)
    P = _phantom((1, 1, 512, 512), np.uint16, 0)[0, 0]
    for x in (P, (P.astype(np.int32) - 1024).astype(np.int16), rng.integers(0, 256, (64, 61)).astype(np.uint8),
              np.full((16, 16), 9, np.uint16)):
        a, b = S.equalize_hist(x), O.sk_equalize_hist(x)
        assert a.dtype == np.float64 and np.array_equal(a, b)
        assert b.max() == 1.0   # the brightest level maps to cdf == 1
    # known answer: four levels with counts 1, 2, 3, 2
    x = np.array([[0, 1, 1, 2], [2, 2, 5, 5]], np.uint8)
    ref = np.array([[1, 3, 3, 6], [6, 6, 8, 8]]) / 8.0
    assert np.array_equal(O.sk_equalize_hist(x), ref)


BIL_CASES = [
    ("phantom crop u16 defaults", lambda r: _phantom((1, 1, 512, 512), np.uint16, 0)[0, 0, 100:228, 100:260], {}),
    ("phantom crop i16 (negative: shifted)", lambda r: (_phantom((1, 1, 512, 512), np.uint16, 0)[0, 0, 100:200, 90:200].astype(np.int32) - 1024).astype(np.int16), {}),
    ("u8 edge sigma 2", lambda r: r.integers(0, 256, (40, 33)).astype(np.uint8), dict(mode="edge", sigma_color=0.1, sigma_spatial=2)),
    ("u16 reflect win 5", lambda r: r.integers(0, 65536, (30, 31)).astype(np.uint16), dict(mode="reflect", win_size=5, sigma_color=0.2)),
    ("u16 wrap 100 bins", lambda r: r.integers(0, 65536, (30, 31)).astype(np.uint16), dict(mode="wrap", sigma_color=0.2, bins=100)),
    ("u16 symmetric", lambda r: r.integers(0, 65536, (30, 31)).astype(np.uint16), dict(mode="symmetric", sigma_color=0.05)),
    ("u16 constant cval 0.3", lambda r: r.integers(0, 65536, (30, 31)).astype(np.uint16), dict(cval=0.3, sigma_color=0.05)),
]


@pytest.mark.parametrize("case", BIL_CASES, ids=lambda c: c[0])
def test_denoise_bilateral_twin_and_restatement_agree(case):
    import oracle as O
    import skimage_twin as S

    _, make, kw = case
    x = make(np.random.default_rng(3))
    a, b = S.denoise_bilateral(x, **kw), O.sk_denoise_bilateral(x, **kw)
    assert a.dtype == np.float64 and np.array_equal(a, b)


def test_denoise_bilateral_semantics_known_answers():
    import skimage_twin as S

    # win_size rule, constant images returned unchanged, smoothing keeps the range
    x = np.full((9, 9), 500, np.uint16)
    assert np.array_equal(S.denoise_bilateral(x), np.full((9, 9), 500 / 65535.0))
    rng = np.random.default_rng(5)
    y = rng.integers(1000, 1100, (24, 24)).astype(np.uint16)
    out = S.denoise_bilateral(y, sigma_color=0.5, sigma_spatial=1, mode="edge")
    f = y / 65535.0
    assert out.min() >= f.min() and out.max() <= f.max()
    # mode='constant': the zero border takes part in the mean, so an all-bright image darkens at its edge
    z = np.full((16, 16), 60000, np.uint16)
    z[8, 8] = 59000
    o = S.denoise_bilateral(z, sigma_color=10.0)
    assert o[0, 0] < o[8, 4]
    with pytest.raises(ValueError):
        S.denoise_bilateral(z, mode="nearest")


# ---------------------------------------------------------------------------------------------------- GPU parity
def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64 if a.dtype == np.float64 else np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("case", _adapt_cases(), ids=lambda c: c[0])
def test_gpu_equalize_adapthist_matches_oracle(dev, case):
    import oracle as O
    from mie_b200 import skimage_compat as K

    _, x, ks, cl, nb = case
    ref = O.sk_equalize_adapthist(x, ks, cl, nb)
    got = K.equalize_adapthist(torch.from_numpy(x).to(dev), ks, cl, nb).cpu().numpy()
    assert got.dtype == ref.dtype and got.shape == ref.shape
    assert np.array_equal(_bits(got), _bits(ref)), int((got != ref).sum())
    if ref.dtype == np.float64:
        got32 = K.equalize_adapthist(torch.from_numpy(x).to(dev), ks, cl, nb, out_dtype=torch.float32).cpu().numpy()
        assert np.array_equal(got32, ref.astype(np.float32))


@pytest.mark.gpu
def test_gpu_adapthist_batches_are_per_plane_and_c4_shape(dev):
    import oracle as O
    from mie_b200 import skimage_compat as K

    x = _phantom((3, 1, 512, 512), np.uint16, 4)
    got = K.equalize_adapthist(torch.from_numpy(x).to(dev)).cpu().numpy()
    for i in range(3):
        assert np.array_equal(_bits(got[i, 0]), _bits(O.sk_equalize_adapthist(x[i, 0])))
    big = _phantom((1, 1, 4096, 4096), np.uint16, 1)[0, 0]      # config-4 radiograph, 512-pixel regions
    ref = O.sk_equalize_adapthist(big)
    gb = K.equalize_adapthist(torch.from_numpy(big).to(dev)).cpu().numpy()
    assert np.array_equal(_bits(gb), _bits(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8])
def test_gpu_equalize_hist_matches_oracle(dev, dtype):
    import oracle as O
    from mie_b200 import skimage_compat as K

    for shape in [(2, 1, 512, 512), (1, 1, 37, 53)]:
        x = _phantom(shape, dtype, 2)
        got = K.equalize_hist(torch.from_numpy(x).to(dev)).cpu().numpy()
        for i in range(shape[0]):
            assert np.array_equal(_bits(got[i, 0]), _bits(O.sk_equalize_hist(x[i, 0])))
    k = np.full((1, 1, 16, 16), 7, dtype)
    assert np.all(K.equalize_hist(torch.from_numpy(k).to(dev)).cpu().numpy() == 1.0)
    with pytest.raises(NotImplementedError):
        K.equalize_hist(torch.zeros(4, 4, device=dev))


@pytest.mark.gpu
@pytest.mark.parametrize("case", BIL_CASES, ids=lambda c: c[0])
def test_gpu_denoise_bilateral_matches_oracle(dev, case):
    import oracle as O
    from mie_b200 import skimage_compat as K

    _, make, kw = case
    x = make(np.random.default_rng(3))
    kw = dict(kw)
    if "sigma_color" not in kw:   # image.std() is summed in a different order on the device: pin it for bit-exactness
        info = np.iinfo(x.dtype)
        f = x / float(info.max) if x.dtype.kind == "u" else (x * 2.0 + 1.0) / 65535.0
        kw["sigma_color"] = float(f.std())
    ref = O.sk_denoise_bilateral(x, **kw)
    got = K.denoise_bilateral(torch.from_numpy(x).to(dev), **kw).cpu().numpy()
    assert np.array_equal(_bits(got), _bits(ref)), float(np.abs(got - ref).max())


@pytest.mark.gpu
def test_gpu_denoise_bilateral_default_sigma_and_batch(dev):
    import oracle as O
    from mie_b200 import skimage_compat as K

    x = _phantom((2, 1, 256, 256), np.uint16, 6)
    got = K.denoise_bilateral(torch.from_numpy(x).to(dev)).cpu().numpy()
    for i in range(2):
        ref = O.sk_denoise_bilateral(x[i, 0])
        assert np.abs(got[i, 0] - ref).max() < 1e-9      # sigma_color = image.std(): summation order differs in the last ulp
    flat = np.full((1, 1, 20, 20), 321, np.uint16)
    assert np.array_equal(K.denoise_bilateral(torch.from_numpy(flat).to(dev)).cpu().numpy(), flat / 65535.0)
