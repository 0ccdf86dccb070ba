"""Size-independent properties: hypothesis-driven on the CPU oracle, and on the CUDA path at the
FULL sizes of BASELINE.json configs 3 and 4, where running the oracle over the whole input would take too
long — flips and offsets commute with selection filters, local filters agree with the oracle on crops, a
constant image stays constant, histograms conserve mass."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

SMALL = settings(max_examples=25, deadline=None, derandomize=True)   # fixed examples: the suite is a gate, not a fuzzer


# ---------------------------------------------------------------------------- CPU: oracle
@SMALL
@given(st.integers(5, 40), st.integers(5, 40), st.integers(0, 2 ** 31 - 1), st.sampled_from([3, 5]),
       st.sampled_from(["constant", "replicate", "reflect", "symmetric"]))
def test_oracle_median_commutes_with_flips_and_offsets(h, w, seed, k, border):
    import oracle as O

    x = np.random.default_rng(seed).integers(0, 3000, (1, h, w)).astype(np.int16)
    m = O.median_blur(x, k, border)
    assert np.array_equal(O.median_blur(x[:, ::-1, ::-1].copy(), k, border), m[:, ::-1, ::-1])
    if border != "constant":  # zero padding does not move with the offset
        assert np.array_equal(O.median_blur((x + 1234).astype(np.int16), k, border), m + 1234)
    c = np.full((1, h, w), 77, np.int16)
    if border != "constant":
        assert np.array_equal(O.median_blur(c, k, border), c)


@SMALL
@given(st.integers(8, 70), st.integers(8, 70), st.integers(1, 4), st.integers(1, 4), st.floats(0.0, 8.0),
       st.integers(0, 2 ** 31 - 1))
def test_oracle_clahe_invariants(h, w, gh, gw, clip, seed):
    import oracle as O

    try:
        th, tw = O.kornia_tile_size(h, w, (gh, gw))
    except ValueError:
        return
    if th * gh - h >= h or tw * gw - w >= w:   # kornia: padding must stay below the image size
        return
    x = np.random.default_rng(seed).random((1, h, w), dtype=np.float32)
    hist = O.clahe_hist(x, (gh, gw))
    assert (hist.sum(-1) == th * tw).all()                      # mass conserved (reflect padding included)
    luts = O.clahe_luts(x, float(clip), (gh, gw))
    assert (np.diff(luts.astype(np.int32), axis=-1) >= 0).all()  # monotone
    y = O.equalize_clahe(x, float(clip), (gh, gw))
    assert y.min() >= 0.0 and y.max() <= 1.0
    k = O.equalize_clahe(np.full((1, h, w), 0.5, np.float32), float(clip), (gh, gw))
    assert k.min() == k.max()                                   # a constant image stays constant


@SMALL
@given(st.integers(10, 60), st.integers(10, 60), st.sampled_from([3, 5, 9]), st.floats(0.5, 2.5),
       st.sampled_from(["reflect", "replicate", "symmetric", "circular"]), st.integers(0, 2 ** 31 - 1))
def test_oracle_gaussian_invariants(h, w, k, sigma, border, seed):
    import oracle as O

    c = np.full((1, h, w), 0.625, np.float32)
    assert np.abs(O.gaussian_blur2d(c, k, sigma, border) - 0.625).max() <= 2e-7   # weights sum to 1
    x = np.random.default_rng(seed).random((1, h, w), dtype=np.float32)
    g = O.gaussian_blur2d(x, k, sigma, border)
    assert g.min() >= x.min() - 1e-6 and g.max() <= x.max() + 1e-6                 # convex combination
    f = O.gaussian_blur2d(x[:, ::-1, ::-1].copy(), k, sigma, border)[:, ::-1, ::-1]
    assert np.abs(f - g).max() <= 1e-6                                             # symmetric kernel
    u = O.unsharp_mask(x, k, sigma, border)
    assert np.array_equal(u, x + (x - g))


# ---------------------------------------------------------------------------- GPU: full BASELINE sizes
@pytest.mark.gpu
def test_config3_full_volume_properties(dev):
    """512^3 int16 (BASELINE.json config 3): the oracle on z-crops (with their halo planes), flips and an
    offset over the whole volume, per-slice CLAHE on the first / last median planes."""
    import mie_b200 as M
    import oracle as O
    from mie_b200 import synthetic

    vol = synthetic.phantom_volume((512, 512, 512), np.int16, seed=0)
    vt = torch.from_numpy(vol).to(dev)
    med = M.median(vt)
    for z0 in (0, 250, 500):                               # 12-plane crops incl. both volume faces
        z1 = z0 + 12
        ref = O.median3d(vol[z0:z1], "nearest", halo_lo=vol[z0 - 1] if z0 > 0 else None,
                         halo_hi=vol[z1] if z1 < 512 else None)
        assert np.array_equal(med[z0:z1].cpu().numpy(), ref), z0
    assert torch.equal(M.median(vt.flip(0, 1, 2).contiguous()), med.flip(0, 1, 2))       # selection commutes with flips
    assert torch.equal(M.median((vt + 1000).to(torch.int16)), (med + 1000).to(torch.int16))   # ... and offsets
    out = M.median3d_clahe_slab(vt, 2.0, (8, 8))
    for z in (0, 511):
        ref = O.from01(O.equalize_clahe(O.to01(med[z].cpu().numpy()[None]), 2.0, (8, 8)), np.int16)[0]
        assert np.array_equal(out[z].cpu().numpy(), ref), z


@pytest.mark.gpu
def test_config4_full_image_properties(dev):
    """4096 x 4096 uint16 (BASELINE.json config 4: 9x9 bilateral + CLAHE 16x16): bilateral against the oracle
    on crops (a local filter: the crop plus its 4-pixel margin decides), CLAHE of the full image against the
    oracle, a constant image stays constant."""
    import mie_b200 as M
    import oracle as O
    from mie_b200 import synthetic

    x = synthetic.phantom((2, 1, 4096, 4096), np.uint16, seed=0)
    xt = torch.from_numpy(x).to(dev)
    bfast = M.bilateral_blur(xt, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)       # default: MUFU.EX2 colour weights
    with M.kernel_policy("bilateral_exact_exp"):                                     # reproducible: the oracle's bits
        b = M.bilateral_blur(xt, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)
    assert float(((bfast - b).abs() / b.abs().clamp_min(1e-3)).max()) <= 1e-5
    x01 = O.to01(x)
    for (i, y0, x0) in [(0, 0, 0), (1, 2000, 1900), (0, 3896, 3896), (1, 0, 3900)]:
        ya, yb, xa, xb = max(y0 - 4, 0), min(y0 + 200 + 4, 4096), max(x0 - 4, 0), min(x0 + 200 + 4, 4096)
        crop = O.bilateral_blur(x01[i:i + 1, :, ya:yb, xa:xb], 9, 0.1, 1.5, "reflect")
        # interior of the crop only (the crop's own border rule differs from the image's unless it is the image edge)
        iy0, ix0 = (0 if ya == 0 else 4), (0 if xa == 0 else 4)
        iy1, ix1 = (yb - ya if yb == 4096 else yb - ya - 4), (xb - xa if xb == 4096 else xb - xa - 4)
        got = b[i, 0, ya + iy0:ya + iy1, xa + ix0:xa + ix1].cpu().numpy()
        assert np.array_equal(got, crop[0, 0, iy0:iy1, ix0:ix1]), (i, y0, x0)
    c = M.equalize_clahe(b, 2.0, (16, 16))
    ref = O.equalize_clahe(b[:1].cpu().numpy(), 2.0, (16, 16))
    assert np.array_equal(c[:1].cpu().numpy(), ref)
    k = torch.full((1, 1, 4096, 4096), 1234, dtype=torch.uint16, device=dev)
    kb = M.bilateral_blur(k, 9, 0.1, (1.5, 1.5))
    assert bool((kb.view(torch.int16) == 1234).all())        # torch has no min / max for uint16 on CUDA
    kc = M.equalize_clahe(k, 2.0, (16, 16)).view(torch.int16)
    assert bool((kc == kc.flatten()[0]).all())
