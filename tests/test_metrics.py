"""Quality metrics: the numpy/scipy restatement of sewar's mse / rmse / psnr / ssim
(oracle/oracle.py, CPU tests) and the device kernels against it (GPU tests).  Integer sums are exact, so
mse / psnr must agree to the last bit whenever numpy's own float64 sum is exact (sums < 2^53); ssim is
float64 on both sides with different summation orders: rel 1e-9."""
import numpy as np
import pytest
import torch


def _pair(dtype, shape, seed, noise=40):
    rng = np.random.default_rng(seed)
    if dtype == np.float32:
        a = rng.random(shape, dtype=np.float32)
        return a, np.clip(a + rng.normal(0, 0.05, shape).astype(np.float32), 0, 1)
    info = np.iinfo(dtype)
    a = rng.integers(info.min, info.max + 1, shape).astype(np.int64)
    b = np.clip(a + rng.integers(-noise, noise + 1, shape), info.min, info.max)
    return a.astype(dtype), b.astype(dtype)


# ---------------------------------------------------------------------------- CPU: the oracle itself
def test_oracle_metrics_known_answers():
    import oracle as O

    a = np.zeros((8, 8), np.uint8)
    b = np.full((8, 8), 3, np.uint8)
    assert O.sewar_mse(a, b) == 9.0 and O.sewar_rmse(a, b) == 3.0
    assert O.sewar_psnr(a, a) == float("inf")
    assert O.sewar_psnr(a, b) == pytest.approx(10 * np.log10(255.0 ** 2 / 9.0), rel=1e-15)
    x, y = _pair(np.uint16, (2, 40, 52), 0)
    s, c = O.sewar_ssim(x, x)
    assert s == pytest.approx(1.0, abs=1e-12) and c == pytest.approx(1.0, abs=1e-12)
    s, c = O.sewar_ssim(x, y)
    assert 0.0 < s <= c <= 1.0
    # the 'valid' uniform window against a direct sliding-window evaluation
    from numpy.lib.stride_tricks import sliding_window_view as swv

    g, p = x[0].astype(np.float64), y[0].astype(np.float64)
    m = lambda v: swv(v, (11, 11)).mean((-1, -2))  # noqa: E731
    C1, C2 = (0.01 * 65535) ** 2, (0.03 * 65535) ** 2
    va, vb, cab = m(g * g) - m(g) ** 2, m(p * p) - m(p) ** 2, m(g * p) - m(g) * m(p)
    direct = np.mean(((2 * m(g) * m(p) + C1) * (2 * cab + C2)) / ((m(g) ** 2 + m(p) ** 2 + C1) * (va + vb + C2)))
    assert O.sewar_ssim(x[:1], y[:1])[0] == pytest.approx(direct, rel=1e-9)


def test_metrics_reject_cpu_tensors_and_mismatches():
    import mie_b200 as M

    a = torch.zeros((8, 8), dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        M.mse(a, a)  # no CPU path


# ---------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_mse_rmse_psnr_mae(dev, dtype):
    import mie_b200 as M
    import oracle as O

    for shape in [(3, 1, 64, 80), (2, 2, 33, 129), (7, 5), (1, 1, 512, 512)]:
        a, b = _pair(dtype, shape, 1)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        kw = {"MAX": 1.0} if dtype == np.float32 else {}
        if dtype == np.float32:
            assert M.mse(ta, tb) == pytest.approx(O.sewar_mse(a, b), rel=1e-12)
            assert M.psnr(ta, tb, **kw) == pytest.approx(O.sewar_psnr(a, b, **kw), rel=1e-12)
        else:  # exact integer sums
            assert M.mse(ta, tb) == pytest.approx(O.sewar_mse(a, b), rel=1e-15)
            assert M.rmse(ta, tb) == pytest.approx(O.sewar_rmse(a, b), rel=1e-15)
            assert M.psnr(ta, tb) == pytest.approx(O.sewar_psnr(a, b), rel=1e-15)
            planes_a = a.reshape((-1,) + a.shape[-2:]).astype(np.int64)
            planes_b = b.reshape((-1,) + b.shape[-2:]).astype(np.int64)
            per = M.mse(ta, tb, per_plane=True).numpy()
            exact = ((planes_a - planes_b) ** 2).sum((1, 2)) / (a.shape[-1] * a.shape[-2])
            assert np.array_equal(per, exact)
            assert M.mae(ta, tb) == pytest.approx(float(np.abs(planes_a - planes_b).mean()), rel=1e-15)
        assert M.psnr(ta, ta, **kw) == float("inf")
    with pytest.raises(ValueError):
        M.psnr(torch.zeros((4, 4), device=dev), torch.zeros((4, 4), device=dev))  # float needs MAX
    with pytest.raises(AssertionError):
        M.mse(torch.zeros((4, 4), device=dev), torch.zeros((4, 5), device=dev))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.float32])
def test_ssim_against_oracle(dev, dtype):
    import mie_b200 as M
    import oracle as O

    for shape, ws in [((2, 1, 64, 80), 11), ((1, 2, 45, 131), 11), ((1, 1, 16, 16), 8), ((1, 1, 40, 33), 16),
                      ((1, 1, 11, 11), 11), ((1, 1, 70, 70), 1)]:
        a, b = _pair(dtype, shape, 2)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        kw = {"MAX": 1.0} if dtype == np.float32 else ({"MAX": 32767} if dtype == np.int16 else {})
        got = M.ssim(ta, tb, ws=ws, **kw)
        ref = O.sewar_ssim(a, b, ws=ws, **kw)
        assert got[0] == pytest.approx(ref[0], rel=1e-9, abs=1e-12), (shape, ws)
        assert got[1] == pytest.approx(ref[1], rel=1e-9, abs=1e-12), (shape, ws)
        s, c = M.ssim(ta, ta, ws=ws, **kw)
        assert s == pytest.approx(1.0, abs=1e-9) and c == pytest.approx(1.0, abs=1e-9)
    # reproducible run to run (fixed reduction order)
    a, b = _pair(np.uint16, (4, 1, 256, 256), 3)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    r = [M.ssim(ta, tb, per_plane=True)[0].numpy() for _ in range(3)]
    assert np.array_equal(r[0], r[1]) and np.array_equal(r[0], r[2])
    with pytest.raises(ValueError):
        M.ssim(ta, tb, ws=300)
    with pytest.raises(NotImplementedError):
        M.ssim(ta, tb, mode="same")


@pytest.mark.gpu
def test_metrics_on_the_enhancement_chain(dev):
    """The evaluation loop the reference's dependency set implies: enhance, then score against the input."""
    import mie_b200 as M
    import oracle as O
    from mie_b200 import synthetic

    x = synthetic.phantom((4, 1, 512, 512), np.uint16, seed=4)
    xt = torch.from_numpy(x).to(dev)
    y = M.enhance_chain(xt)
    yn = y.cpu().numpy()
    assert M.psnr(xt, y) == pytest.approx(O.sewar_psnr(x, yn), rel=1e-15)
    assert M.ssim(xt, y)[0] == pytest.approx(O.sewar_ssim(x, yn)[0], rel=1e-9)
