"""mie_set_kernel_policy (include/mie.h): the explicit verification hook that replaced round 1's MIE_* environment
switches.  CPU part: the policy word is plain host state.  GPU part: every operator that has a tuned and a generic
kernel must return the SAME BITS from both on inputs the tuned kernel covers."""
import numpy as np
import pytest
import torch


def test_policy_word_is_host_state_and_rejects_unknown_bits():
    import mie_b200 as M
    from mie_b200 import _ffi

    L = M._lib()
    assert L.mie_get_kernel_policy() == 0
    assert L.mie_set_kernel_policy(1 | 32) == 0 and L.mie_get_kernel_policy() == 33
    assert L.mie_set_kernel_policy(1 << 20) == -11 and L.mie_get_kernel_policy() == 33   # unknown bit: unchanged
    assert L.mie_set_kernel_policy(0) == 0
    with M.kernel_policy("generic_median", "generic_nlm"):
        assert L.mie_get_kernel_policy() == 32 | 128
        with M.kernel_policy("generic_gauss"):
            assert L.mie_get_kernel_policy() == 32 | 128 | 1
        assert L.mie_get_kernel_policy() == 32 | 128
    assert L.mie_get_kernel_policy() == 0
    assert sum(_ffi.POLICY.values()) == 16383      # MIE_POLICY_ALL: every bit has a Python name


def test_no_getenv_left_in_the_library_sources():
    import glob
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "medical-image-enhancement-system_b200", "csrc", "*.cu*")):
        assert "getenv" not in open(f).read(), f


def _x(dev, shape, dtype, seed=0, kind="P"):
    from mie_b200 import synthetic

    return torch.from_numpy(synthetic.make(kind, shape, dtype, seed)).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8])
def test_tuned_and_generic_kernels_agree_bit_for_bit(dev, dtype):
    import mie_b200 as M

    x = _x(dev, (6, 1, 256, 512), dtype, seed=11)
    ops = [
        ("generic_gauss", lambda: M.gaussian_blur2d(x, 9, 1.0)),
        ("generic_gauss", lambda: M.unsharp_mask(x, 9, 1.0)),
        ("generic_clahe", lambda: M.equalize_clahe(x, 2.0, (4, 8))),
        ("clahe_float_rules", lambda: M.equalize_clahe(x, 2.0, (4, 8))),
        ("generic_equalize", lambda: M.equalize(x)),
        ("equalize_float_rules", lambda: M.equalize(x)),
        ("equalize_three_pass", lambda: M.equalize(x)),
        ("equalize_slab", lambda: M.equalize(x)),
        ("equalize_slab", lambda: M.equalize(x[:, :, :37].contiguous(), out_dtype=torch.float32)),
        ("generic_median", lambda: M.median_blur(x, 3)),
        ("generic_median", lambda: M.median_blur(x, 5)),
        ("generic_bilateral", lambda: M.bilateral_blur(x[:2], 5, 0.1, (1.5, 1.5))),
    ]
    for policy, fn in ops:
        # the bilateral's default colour weight is MUFU.EX2 (tolerance mode); its two reproducible kernels are compared
        with M.kernel_policy(*(["bilateral_exact_exp"] if policy == "generic_bilateral" else [])):
            tuned = fn().cpu()
            with M.kernel_policy(policy):
                generic = fn().cpu()
        assert torch.equal(tuned.view(torch.uint8), generic.view(torch.uint8)), policy
    # non-local means is a float filter whose two kernels sum the patch distances in different orders: both are held
    # to the north star's tolerance against the float64 oracle (tests/test_gpu_ops.py), i.e. <= 1 LSB between them
    xn = x[0, 0, :96, :128].contiguous()
    tuned = M.denoise_nl_means(xn, 5, 4, 0.1).cpu().to(torch.int32)
    with M.kernel_policy("generic_nlm"):
        generic = M.denoise_nl_means(xn, 5, 4, 0.1).cpu().to(torch.int32)
    assert int((tuned - generic).abs().max()) <= 1
    assert M._lib().mie_get_kernel_policy() == 0


@pytest.mark.gpu
def test_volume_median_and_clahe16_variants_agree(dev):
    import mie_b200 as M

    v = _x(dev, (40, 96, 128), np.int16, seed=5)
    a = M.median(v).cpu()
    with M.kernel_policy("generic_median"):
        b = M.median(v).cpu()
    assert torch.equal(a, b)
    x = _x(dev, (3, 1, 256, 256), np.uint16, seed=9)
    ref = M.equalize_clahe(x, 2.0, (4, 4), semantics="opencv").cpu()
    for pol in ("clahe16_no_cluster", "clahe16_two_sweep", "clahe16_full_luts"):
        with M.kernel_policy(pol):
            got = M.equalize_clahe(x, 2.0, (4, 4), semantics="opencv").cpu()
        assert torch.equal(ref.view(torch.int16), got.view(torch.int16)), pol
    # LUTs bounded by the batch's largest pixel value (default) against full LUTs: bounds at 0, inside a thread's 64 bins,
    # at the half boundary, above it, at 65535; geometries with padding
    # (the bound is applied from 296 tiles on: batches of 5 .. 50 images here)
    for shape, grid, clip, top in [((5, 1, 300, 500), (8, 8), 2.0, 4095), ((6, 1, 256, 250), (8, 8), 40.0, 32767),
                                   ((50, 1, 64, 61), (2, 3), 3.0, 32768), ((20, 1, 128, 128), (4, 4), 2.0, 65535),
                                   ((20, 1, 200, 136), (5, 3), 0.0, 100), ((5, 1, 128, 128), (8, 8), 4.0, 0),
                                   ((19, 1, 128, 128), (4, 4), 4.0, 40000), ((5, 1, 128, 128), (8, 8), 2.0, 63)]:
        rng = np.random.default_rng(top)
        y = torch.from_numpy(rng.integers(0, top + 1, shape, dtype=np.uint16)).to(dev)
        a = M.equalize_clahe(y, clip, grid, semantics="opencv").cpu()
        with M.kernel_policy("clahe16_full_luts"):
            b = M.equalize_clahe(y, clip, grid, semantics="opencv").cpu()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (shape, grid, clip, top)
