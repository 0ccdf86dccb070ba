"""Generates tests/golden/cv2_clahe.json + cv2_clahe16_128.npz: outputs of cv2.createCLAHE (OpenCV 4.13,
an implementation independent of this repository) on seeded inputs.  cv::CLAHE is the verification
anchor of SURVEY.md §8(a) A1': the oracle (CPU suite) and the CUDA path (GPU suite) must both reproduce
these vectors bit for bit, so this part of the parity claim is pinned to a third-party binary rather
than to the restatement itself.

    python tests/golden/make_cv2_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
import sys  # noqa: E402

sys.path.insert(0, HERE)


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


from cv2_inputs import image  # noqa: E402  (same directory)


# (name, h, w, dtype, bits, (grid rows, grid cols), clip limit)
CASES = [
    ("u16_512_8x8_clip2", 512, 512, "uint16", 16, (8, 8), 2.0),          # BASELINE.json config 1 in 16-bit-native mode
    ("u16_512_8x8_clip40", 512, 512, "uint16", 16, (8, 8), 40.0),
    ("u16_ct12bit_512_8x8_clip2", 512, 512, "uint16", 12, (8, 8), 2.0),
    ("u16_300x500_8x8_clip2", 300, 500, "uint16", 16, (8, 8), 2.0),      # padding (reflect-101)
    ("u16_1024_16x16_clip2", 1024, 1024, "uint16", 12, (16, 16), 2.0),
    ("u16_512_8x8_noclip", 512, 512, "uint16", 16, (8, 8), 0.0),
    ("u16_512_2x2_clip2", 512, 512, "uint16", 8, (2, 2), 2.0),           # tiles of 65 536 pixels
    ("u16_256_1x1_clip2_16levels", 256, 256, "uint16", 4, (1, 1), 2.0),
    ("u16_37x53_3x5_clip1.5", 37, 53, "uint16", 16, (3, 5), 1.5),
    ("u8_512_8x8_clip2", 512, 512, "uint8", 8, (8, 8), 2.0),
    ("u8_300x500_8x8_clip40", 300, 500, "uint8", 8, (8, 8), 40.0),
    # one axis divisible, the other not: cv::CLAHE then pads the divisible axis by a full `tiles` pixels
    ("u8_512x500_8x8_clip2", 512, 500, "uint8", 8, (8, 8), 2.0),
    ("u8_300x512_8x8_clip2", 300, 512, "uint8", 8, (8, 8), 2.0),
    ("u16_512x500_8x8_clip2", 512, 500, "uint16", 12, (8, 8), 2.0),
    ("u16_300x512_8x8_clip2", 300, 512, "uint16", 16, (8, 8), 2.0),
    ("u8_64x61_4x4_clip3", 64, 61, "uint8", 8, (4, 4), 3.0),
]


def main():
    out = {"cv2_version": cv2.__version__, "cases": {}}
    for name, h, w, dt, bits, grid, clip in CASES:
        img = image(h, w, np.dtype(dt), seed=len(name), bits=bits)
        ref = cv2.createCLAHE(clip, (grid[1], grid[0])).apply(img)
        out["cases"][name] = {"h": h, "w": w, "dtype": dt, "bits": bits, "grid": list(grid), "clip": clip,
                              "seed": len(name), "input_sha256": sha(img), "output_sha256": sha(ref)}
    with open(os.path.join(HERE, "cv2_clahe.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    x = image(128, 128, np.uint16, seed=77, bits=12)
    np.savez_compressed(os.path.join(HERE, "cv2_clahe16_128.npz"), input=x,
                        output=cv2.createCLAHE(2.0, (4, 4)).apply(x), grid=np.array([4, 4]), clip=np.array(2.0))
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
