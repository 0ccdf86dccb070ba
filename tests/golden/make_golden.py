"""Generates tests/golden/golden.json (+ small .npz fixtures): seeded inputs -> sha256 of the CPU
oracle's outputs.  The reference repository ships no golden vectors (0 lines of code, no tests),
and kornia / scikit-image cannot be imported in this image, so these vectors are produced by the
oracle restatement itself, after it has been pinned against cv2 / scipy / torchvision / the kornia
twin by tests/test_oracle.py.  Their job is to detect drift (of the oracle or of the CUDA path)
between sessions.

    python tests/golden/make_golden.py          # rewrite golden.json and the fixtures
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def inputs():
    """Seeded inputs shared with the GPU parity tests (tests/test_gpu_golden.py)."""
    from mie_b200 import synthetic

    rng = np.random.default_rng(1234)
    return {
        "phantom_u16": synthetic.phantom((2, 1, 512, 512), np.uint16, seed=1),
        "uniform_u16": synthetic.uniform((1, 1, 256, 256), np.uint16, seed=2),
        "phantom_i16_vol": synthetic.phantom_volume((24, 96, 80), np.int16, seed=3),
        "noise_f32": rng.random((2, 1, 100, 130), dtype=np.float32),
        "noise_u8": rng.integers(0, 256, (1, 1, 300, 500), dtype=np.uint8),
    }


def compute() -> dict:
    import oracle as O

    x = inputs()
    out = {}
    p = x["phantom_u16"]
    out["chain_c2_phantom_u16"] = sha(O.chain_gauss_clahe_unsharp(p))
    out["chain_c2_uniform_u16"] = sha(O.chain_gauss_clahe_unsharp(x["uniform_u16"]))
    out["clahe_luts_phantom_u16_8x8_clip2"] = sha(O.clahe_luts(O.to01(p), 2.0, (8, 8)))
    out["clahe_phantom_u16_8x8_clip2"] = sha(O.from01(O.equalize_clahe(O.to01(p), 2.0, (8, 8)), np.uint16))
    f = x["noise_f32"]
    out["clahe_f32_4x6_clip2"] = sha(O.equalize_clahe(f, 2.0, (4, 6)))
    out["gauss_f32_k9_s1_reflect"] = sha(O.gaussian_blur2d(f, 9, 1.0))
    out["gauss_f32_k5_s1.2_replicate"] = sha(O.gaussian_blur2d(f, 5, 1.2, "replicate"))
    out["unsharp_f32_k9_s1_reflect"] = sha(O.unsharp_mask(f, 9, 1.0))
    out["opencv_clahe_u8_8x8_clip2"] = sha(O.opencv_clahe(x["noise_u8"], 2.0, (8, 8)))
    out["median3x3_u16_zero"] = sha(O.median_blur(x["uniform_u16"], 3))
    out["median5x5_u16_zero"] = sha(O.median_blur(x["uniform_u16"], 5))
    out["median3d_i16_nearest"] = sha(O.median3d(x["phantom_i16_vol"]))
    out["bilateral_f32_k9_sc0.1_ss1.5"] = sha(O.bilateral_blur(f, 9, 0.1, 1.5))
    out["equalize_u8"] = sha(O.from01(O.equalize(O.to01(x["noise_u8"])), np.uint8))
    out["equalize_f32"] = sha(O.equalize(f))
    return out


def small_fixture():
    """A fixture small enough to commit as data: 1 x 128 x 128 uint16 phantom -> chain output + LUTs."""
    import oracle as O
    from mie_b200 import synthetic

    x = synthetic.phantom((1, 1, 128, 128), np.uint16, seed=9)
    y, st = O.chain_gauss_clahe_unsharp(x, grid_size=(2, 2), return_stages=True)
    return {"input": x, "output": y, "luts": st["luts"]}


if __name__ == "__main__":
    with open(os.path.join(HERE, "golden.json"), "w") as fjson:
        json.dump(compute(), fjson, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "chain_128_grid2.npz"), **small_fixture())
    print("wrote golden.json, chain_128_grid2.npz")
