"""CPU checks of the arithmetic identities the tuned kernels rely on (csrc/chain_fast.cuh) and of host-side
helpers.  fp32 fma is emulated in float64, which is exact here: every product / sum involved fits in 53 bits
before the single rounding to fp32."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def test_u16_to_unit_without_divide_is_exact_for_all_codes():
    v = np.arange(65536, dtype=np.uint32)
    bits = (np.uint32(0x43000000) | v).astype(np.uint32)           # OR v into the mantissa of 128.0f
    t = (bits.view(np.float32) - np.float32(128.0)).astype(np.float32)   # v * 2^-16, exact
    assert np.array_equal(t.astype(np.float64), v / 65536.0)
    c = np.float32(1.5259021893143654e-05)                          # RN(1/65535)
    assert c == np.float32(1.0 / 65535.0)
    x = fma32(t, c, t)
    assert np.array_equal(x, v.astype(np.float32) / np.float32(65535.0))


def test_i16_and_u8_variants():
    s = np.arange(-32768, 32768, dtype=np.int32)
    u = (s.astype(np.uint32) & 0xFFFF) ^ 0x8000                      # v + 32768
    t = ((np.uint32(0x43000000) | u).astype(np.uint32).view(np.float32) - np.float32(128.0)).astype(np.float32)
    x = fma32(t, np.float32(1.0 / 65535.0), t)
    assert np.array_equal(x, (s.astype(np.float32) + np.float32(32768.0)) / np.float32(65535.0))
    b = np.arange(256, dtype=np.uint32)
    t8 = ((np.uint32(0x47000000) | b).astype(np.uint32).view(np.float32) - np.float32(32768.0)).astype(np.float32)
    x8 = fma32(t8, np.float32(0.003921568859368563), t8)
    assert np.array_equal(x8, b.astype(np.float32) / np.float32(255.0))


def test_markstein_division_by_255_is_correctly_rounded():
    rng = np.random.default_rng(0)
    o = np.concatenate([(rng.random(4_000_000) * 255).astype(np.float32), np.arange(256, dtype=np.float32),
                        np.float32([1e-6, 254.99998, 255.0, 0.0])])
    r = np.float32(0.003921568859368563)
    q0 = (o * r).astype(np.float32)
    q = fma32(fma32(np.float32(-255.0), q0, o), r, q0)
    assert np.array_equal(q, (o / np.float32(255.0)).astype(np.float32))


def test_magic_number_floor_trunc_rint():
    rng = np.random.default_rng(1)
    g = np.concatenate([rng.random(1_000_000, dtype=np.float32), np.float32([0.0, 1.0, 0.5, 255 / 256.0])])
    magic = np.float32(8388608.0)
    # floor(g*256): the fma is exact before its round-down, so floor of the exact product
    n = np.floor(g.astype(np.float64) * 256.0 + float(magic)).astype(np.int64) - 0x800000
    assert np.array_equal(np.minimum(n, 255), np.minimum(np.floor(g * np.float32(256.0)), 255).astype(np.int64))
    # trunc(fl(g*255)): round the product to fp32 first, then add 2^23 rounding down
    f = (g * np.float32(255.0)).astype(np.float32)
    idx = np.floor(f.astype(np.float64) + float(magic)).astype(np.int64) - 0x800000
    assert np.array_equal(idx, f.astype(np.int64))
    # rint(c*65535) by adding 2^23 in round-to-nearest-even
    p = (g * np.float32(65535.0)).astype(np.float32)
    q = (p + magic).astype(np.float32).view(np.uint32) & 0xFFFF
    assert np.array_equal(q.astype(np.int64), np.rint(p).astype(np.int64))


def test_axis_weight_table_matches_kornia_axis():
    """The per-position weight table of chain_b (csrc/chain_fast.cu: launch_chain_b_fast) equals
    (T-1-r)/(T-1) of the oracle's kornia_axis for every interior tile position."""
    T = 64
    for k in range(T + 8):
        p = k - 4
        r = p + T // 2 if p < T // 2 else p - T // 2
        w = np.float32(T - 1 - r) / np.float32(T - 1)
        for ty0 in (64, 192):                       # interior tiles of a 512-row image
            y = ty0 + p
            rel = y - T // 2
            j0 = rel // T
            assert rel - j0 * T == r
            assert w == np.float32(T - 1 - (rel - j0 * T)) / np.float32(T - 1)


def test_gaussian_kernel_host_function_matches_oracle():
    import mie_b200 as M
    import oracle as O

    for k, s in [(9, 1.0), (7, 1.0), (5, 1.2), (33, 5.0), (4, 1.0), (1, 1.0)]:
        a, b = M.get_gaussian_kernel1d(k, s), O.gaussian_kernel1d(k, s)
        assert a.dtype == np.float32 and np.array_equal(a, b)
        assert abs(float(a.astype(np.float64).sum()) - 1.0) < 1e-6


def test_committed_fixture_matches_oracle():
    import oracle as O

    fx = np.load(os.path.join(GOLDEN, "chain_128_grid2.npz"))
    out, st = O.chain_gauss_clahe_unsharp(fx["input"], grid_size=(2, 2), return_stages=True)
    assert np.array_equal(out, fx["output"]) and np.array_equal(st["luts"], fx["luts"])


def test_synthetic_phantom_properties():
    from mie_b200 import synthetic

    p = synthetic.phantom((4, 1, 512, 512), np.uint16, seed=0)
    assert p.dtype == np.uint16 and p.max() <= 4095
    assert (p == 0).mean() >= 0.40                      # >= 40 % air, exactly zero
    v = synthetic.phantom_volume((16, 64, 64), np.int16, seed=0)
    assert v.min() == -1024 and v.max() <= 3071
    assert np.array_equal(synthetic.phantom((2, 1, 64, 64), np.uint16, seed=3), synthetic.phantom((2, 1, 64, 64), np.uint16, seed=3))


def test_chain_config_validation_without_gpu():
    import torch

    import mie_b200 as M

    x = torch.zeros(1, 1, 64, 64, dtype=torch.uint16)
    with pytest.raises(TypeError):
        M.enhance_chain(x, M.ChainConfig(clip_limit=2))
    with pytest.raises(ValueError):
        M.enhance_chain(x, M.ChainConfig(denoise_kernel_size=8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.enhance_chain(x)


# ---------------------------------------------------------------------------- median selection networks
SORT9 = [(0, 3), (1, 7), (2, 5), (4, 8), (0, 7), (2, 4), (3, 8), (5, 6), (0, 2), (1, 3), (4, 5), (7, 8), (1, 4),
         (3, 6), (5, 7), (0, 1), (2, 4), (3, 5), (6, 8), (2, 3), (4, 5), (6, 7), (1, 2), (3, 4), (5, 6)]
MERGE44 = [(0, 4), (1, 5), (2, 6), (3, 7), (2, 4), (3, 5), (1, 2), (3, 4), (5, 6)]


def _cs(a, i, j):
    lo, hi = np.minimum(a[i], a[j]), np.maximum(a[i], a[j])
    a[i], a[j] = lo, hi


def _select27_sorted(A, B, C):
    """Python twin of select27_sorted (csrc/median.cu): rank 13 of 27 from three sorted 9-lists."""
    W, Y = [None] * 8, [None] * 5
    for c in range(9):
        lo, hi = np.minimum(A[c], B[c]), np.maximum(A[c], B[c])
        if c >= 5:
            W[c - 5] = np.minimum(lo, C[c])
        if c <= 3:
            W[4 + c] = np.maximum(hi, C[c])
        if 2 <= c <= 6:
            Y[c - 2] = np.maximum(lo, np.minimum(hi, C[c])) if c <= 4 else np.minimum(hi, np.maximum(lo, C[c]))
    for i, j in MERGE44:
        _cs(W, i, j)
    m = W[6]
    for j in range(5):
        m = np.minimum(m, np.maximum(W[5 - j], Y[j]))
    return m


def test_median27_selection_network():
    # the 9-input sorting network, exhaustively on 0/1 inputs (0-1 principle)
    bits = ((np.arange(512)[:, None] >> np.arange(9)[None, :]) & 1).T.copy()
    a = [bits[i].copy() for i in range(9)]
    for i, j in SORT9:
        _cs(a, i, j)
    assert np.array_equal(np.stack(a), np.sort(bits, axis=0))
    # the selection from three sorted planes: every 0/1 pattern class (counts of ones per list) ...
    cnt = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(10), indexing="ij"), -1).reshape(-1, 3)
    lists = [(np.arange(9)[:, None] >= 9 - cnt[None, :, k]).astype(np.int64) for k in range(3)]
    got = _select27_sorted(list(lists[0]), list(lists[1]), list(lists[2]))
    assert np.array_equal(got, (cnt.sum(1) >= 14).astype(np.int64))
    # ... and random values with many ties
    rng = np.random.default_rng(0)
    for hi in (2, 3, 7, 1000):
        v = rng.integers(0, hi, (27, 20000))
        s = [np.sort(v[9 * k:9 * k + 9], axis=0) for k in range(3)]
        got = _select27_sorted(list(s[0]), list(s[1]), list(s[2]))
        assert np.array_equal(got, np.sort(v, axis=0)[13])


# ---------------------------------------------------------------------------- integer bin / index rules
def test_integer_bin_and_index_rules_for_default_ranges():
    """csrc/window.cuh (Codes<T>): for default-range integer pixels the kornia histogram bin floor(x*256), the
    lookup index trunc(x*255) and equalize's bin floor(RN(RN(x*255)/255)*256) are integer functions of the code:
    16-bit: bin = u >> 8, index = u // 257 = (u * 65281) >> 24;  8-bit: bin = index = u.  Exhaustive check
    against the float32 formulas (the library repeats it in C before the first launch)."""
    f32 = np.float32
    for bits, rng in ((16, 65535), (8, 255)):
        u = np.arange(rng + 1, dtype=np.int64)
        x = (u.astype(f32) / f32(rng)).astype(f32)
        b = np.minimum(np.floor((x * f32(256)).astype(f32)), 255).astype(np.int64)
        v = (x * f32(255)).astype(f32)
        idx = np.floor(v).astype(np.int64)
        eb = np.minimum(np.floor(((v / f32(255)).astype(f32) * f32(256)).astype(f32)), 255).astype(np.int64)
        if bits == 16:
            assert np.array_equal(b, u >> 8) and np.array_equal(eb, u >> 8)
            assert np.array_equal(idx, u // 257) and np.array_equal(idx, (u * 65281) >> 24)
        else:
            assert np.array_equal(b, u) and np.array_equal(eb, u) and np.array_equal(idx, u)


SORT5 = [(0, 3), (1, 4), (0, 2), (1, 3), (0, 1), (2, 4), (1, 2), (3, 4), (2, 3)]
M22 = [(0, 2), (1, 3), (1, 2)]
M33 = [(0, 3), (1, 4), (1, 3), (2, 5), (2, 3), (3, 4)]
M43 = [(0, 4), (1, 4), (1, 5), (2, 5), (3, 6), (3, 4), (2, 3), (4, 5)]


def _net(v, net):
    v = list(v)
    for i, j in net:
        _cs(v, i, j)
    return v


def test_median25_selection_network():
    """Python twin of select25_sorted_columns (csrc/median.cu): rank 12 of 25 from five vertically sorted
    columns — row sorts, pruning to 13 candidates, three merges, rank_6(W u Y)."""
    bits = ((np.arange(32)[:, None] >> np.arange(5)[None, :]) & 1).T.copy()
    assert np.array_equal(np.stack(_net([bits[i].copy() for i in range(5)], SORT5)), np.sort(bits, axis=0))
    rng = np.random.default_rng(1)
    for hi in (2, 3, 7, 1000):
        v = rng.integers(0, hi, (25, 20000))
        cols = [np.sort(v[5 * c:5 * c + 5], axis=0) for c in range(5)]          # cols[c][r]
        R = [_net([cols[c][r] for c in range(5)], SORT5) for r in range(5)]    # rows sorted horizontally
        xz = _net([R[0][3], R[0][4], R[4][0], R[4][1]], M22)
        y = _net([R[1][2], R[1][3], R[1][4], R[3][0], R[3][1], R[3][2]], M33)
        w = _net(xz + [R[2][1], R[2][2], R[2][3]], M43)
        m = w[6]
        for i in range(6):
            m = np.minimum(m, np.maximum(w[i], y[5 - i]))
        assert np.array_equal(m, np.sort(v, axis=0)[12]), hi


def test_loader_chunk_schedule_is_a_partition_with_tapered_ends():
    """HostSlicePipeline.schedule: contiguous chunks covering [0, n) exactly once, none larger than `chunk`; with taper the
    first and last chunks are a quarter / half of it (the exposed ends of the H2D / D2H overlap)."""
    from mie_b200.loader import HostSlicePipeline

    class P(HostSlicePipeline):
        def __init__(self, chunk, taper):   # schedule() needs nothing else
            self.chunk, self.taper = chunk, taper

    for n in (0, 1, 7, 31, 32, 33, 64, 95, 96, 100, 256, 1000):
        for chunk in (1, 4, 8, 32, 64):
            for taper in (False, True):
                spans = P(chunk, taper).schedule(n)
                assert [a for a, _ in spans] == [0] * (1 if spans else 0) + [b for _, b in spans[:-1]]
                assert (spans[-1][1] if spans else 0) == n
                assert all(0 < b - a <= chunk for a, b in spans)
    assert P(32, True).schedule(256)[0] == (0, 8) and P(32, True).schedule(256)[-1] == (248, 256)
    assert P(32, False).schedule(256)[0] == (0, 32)
    assert P(32, True).schedule(64) == [(0, 32), (32, 64)]          # too short to taper


def test_loader_rejects_wrong_dtype_shape_and_device_before_touching_the_gpu():
    """Round-1 advice: a dtype mismatch used to be converted silently by copy_ (a blocking CPU cast, wrong value range)."""
    import torch
    from mie_b200.loader import HostSlicePipeline

    class P(HostSlicePipeline):
        def __init__(self):
            self.h, self.w, self.dtype, self.out_dtype = 16, 24, torch.uint16, torch.uint16

    p = P()
    good = torch.zeros(3, 16, 24, dtype=torch.uint16)
    p._validate(good, good.clone())
    p._validate(good.reshape(3, 1, 16, 24), good.clone())
    with pytest.raises(TypeError):
        p._validate(good.to(torch.int16), good)
    with pytest.raises(TypeError):
        p._validate(good, good.to(torch.float32))
    with pytest.raises(ValueError):
        p._validate(torch.zeros(3, 16, 25, dtype=torch.uint16), good)
    with pytest.raises(ValueError):
        p._validate(good, torch.zeros(2, 16, 24, dtype=torch.uint16))
    with pytest.raises(ValueError):
        p._validate(good.transpose(0, 1).transpose(0, 1)[:, ::2], good)
    with pytest.raises(TypeError):
        p._validate(good.numpy(), good)


def test_two_operation_division_by_255_matches_the_quotient_on_a_sample():
    """div255_x2 (csrc/chain_fast.cuh): q = fma(x, r, RN(x * r2)), r = RN(1/255), r2 = RN(1/255 - r).  The exhaustive
    check over every float in [0, 256] is profiles/microbench/div255_check.c (log beside it); here a sample of blend-like
    values is re-evaluated with exact rational arithmetic."""
    from fractions import Fraction

    r = np.float32(1.0) / np.float32(255.0)
    r2 = np.float32(1.0 / 255.0 - float(r))
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.arange(0, 256, dtype=np.float32), rng.random(3000).astype(np.float32) * np.float32(255.0),
                         np.float32([2.0 ** -60, 2.0 ** -30, 1e-8, 254.99998, 255.0, 256.0])])
    for x in xs:
        p = np.float32(x * r2)                                                       # RN(x * r2)
        q = np.float32(Fraction(float(x)) * Fraction(float(r)) + Fraction(float(p)))  # fma: one rounding
        assert q == np.float32(x) / np.float32(255.0), float(x)
