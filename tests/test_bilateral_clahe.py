"""Fused bilateral -> CLAHE chain (BASELINE.json config 4; include/mie.h mie_bilateral_clahe): bit-exact against the
oracle composition from01(equalize_clahe(bilateral_blur(to01(x)))) and against the unfused CUDA operators in the
reproducible mode (kernel policy bilateral_exact_exp); statistically close in the default MUFU.EX2 mode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    dict(shape=(2, 1, 256, 256), grid=(4, 4), k=9),                       # 64-px tiles
    dict(shape=(1, 1, 128, 192), grid=(2, 3), k=5, border="replicate"),   # 64-px tiles, other window / border
    dict(shape=(1, 1, 96, 128), grid=(3, 4), k=3, border="constant"),     # 32-px tiles: one block per tile
    dict(shape=(1, 1, 256, 256), grid=(1, 1), k=7, clip=0.0),             # single 256-px tile, no clipping
    dict(shape=(1, 1, 128, 128), grid=(2, 2), k=9, border="circular"),
    dict(shape=(1, 1, 512, 512), grid=(2, 2), k=9),                       # config-4 tile size (256 px)
]


def _images(kind, shape, dtype, seed):
    from mie_b200 import synthetic

    if dtype == np.float32:
        return (synthetic.phantom(shape, np.uint16, seed).astype(np.float32) / np.float32(4095.0)).astype(np.float32)
    return synthetic.make(kind, shape, dtype, seed)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8, np.float32])
def test_fused_bilateral_clahe_matches_oracle(dev, case, dtype):
    import mie_b200 as M
    import oracle as O

    k, grid = case["k"], case["grid"]
    border, clip = case.get("border", "reflect"), case.get("clip", 2.0)
    for kind in ("P", "U"):
        x = _images(kind, case["shape"], dtype, seed=3)
        b = O.bilateral_blur(O.to01(x), k, 0.1, (1.5, 1.5), border)
        ref01 = O.equalize_clahe(b, clip, grid)
        ref = O.from01(ref01, dtype)
        xt = torch.from_numpy(x).to(dev)
        with M.kernel_policy("bilateral_exact_exp"):   # reproducible colour weights: every stage bit for bit
            got = M.bilateral_clahe(xt, k, 0.1, (1.5, 1.5), clip, grid, border).cpu().numpy()
            assert np.array_equal(got, ref), (kind, int((got != ref).sum()))
            gf = M.bilateral_clahe(xt, k, 0.1, (1.5, 1.5), clip, grid, border, out_dtype=torch.float32).cpu().numpy()
            assert np.array_equal(gf, ref01)
            # and the unfused CUDA operators give the same float image
            u = M.equalize_clahe(M.bilateral_blur(xt, k, 0.1, (1.5, 1.5), border, out_dtype=torch.float32), clip, grid)
            assert np.array_equal(u.cpu().numpy(), ref01)
        # default colour weights (MUFU.EX2, rel 6e-7): CLAHE is discontinuous — a blurred value that crosses a bin or
        # lookup boundary moves a pixel by whole LUT steps — so the statement is statistical:
        # very few pixels differ, and the mean deviation is a small fraction of one LUT step
        fast = M.bilateral_clahe(xt, k, 0.1, (1.5, 1.5), clip, grid, border, out_dtype=torch.float32).cpu().numpy()
        d = np.abs(fast.astype(np.float64) - ref01.astype(np.float64)) * 255.0     # in LUT steps
        assert (d > 1e-3).mean() <= 5e-3 and d.mean() <= 5e-3, (kind, float((d > 1e-3).mean()), float(d.mean()))
        fq = M.bilateral_clahe(xt, k, 0.1, (1.5, 1.5), clip, grid, border).cpu().numpy()
        assert (fq != ref).mean() <= 5e-3


def test_fused_path_refuses_what_it_does_not_cover(dev):
    import mie_b200 as M

    L = M._lib()
    assert L.mie_bilateral_clahe_is_fused(4096, 4096, 16, 16, 9, 1) == 1     # config 4
    assert L.mie_bilateral_clahe_is_fused(500, 300, 8, 8, 9, 1) == 0          # CLAHE padding
    assert L.mie_bilateral_clahe_is_fused(512, 512, 8, 8, 11, 1) == 0         # 11 taps
    assert L.mie_bilateral_clahe_is_fused(160, 160, 4, 4, 9, 1) == 0          # 40-px tiles
    x = torch.zeros(1, 1, 160, 160, dtype=torch.uint16, device=dev)
    with pytest.raises(NotImplementedError):
        M.bilateral_clahe(x, 9, 0.1, (1.5, 1.5), 2.0, (4, 4))
    with pytest.raises(NotImplementedError):
        M.bilateral_clahe(x, (9, 7), 0.1, (1.5, 1.5), 2.0, (5, 5))
    ws = torch.empty(M.chain.bilateral_clahe_workspace_bytes(1, 128, 128, (2, 2)) + 512, dtype=torch.uint8, device=dev)
    y = torch.zeros(1, 1, 128, 128, dtype=torch.uint16, device=dev)
    with pytest.raises(ValueError):
        M.bilateral_clahe(y, 9, 0.1, (1.5, 1.5), 2.0, (2, 2), workspace=ws[1:])


def test_plan_reuses_buffers_and_stages_compose(dev):
    import mie_b200 as M
    from mie_b200 import synthetic

    x = torch.from_numpy(synthetic.phantom((2, 1, 512, 512), np.uint16, 8)).to(dev)
    ref = M.bilateral_clahe(x, grid_size=(2, 2))
    plan = M.BilateralClahePlan(x, grid_size=(2, 2))
    assert torch.equal(plan.run().view(torch.int16), ref.view(torch.int16))
    plan.out.zero_()
    for mask in (1, 2, 4):
        plan.run(mask)
    assert torch.equal(plan.out.view(torch.int16), ref.view(torch.int16))
    ms = plan.stage_ms()
    assert set(ms) == {"bilateral_index_hist", "hist_to_lut", "pack_cells+apply_index"} and all(v > 0 for v in ms.values())
