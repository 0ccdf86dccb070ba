"""Seeded inputs of the cv2 golden vectors (shared by make_cv2_golden.py and the tests; no cv2 import)."""
import numpy as np


def image(h, w, dtype, seed, bits):
    """Smooth-plus-noise image using `bits` bits of the container (12 = CT-like occupancy of uint16)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.5 + 0.23 * np.sin(xx / 37.0) + 0.2 * np.cos(yy / 23.0) + rng.normal(0, 0.05, (h, w))
    return (np.clip(img, 0, 1) * (2 ** bits - 1)).astype(dtype)
