"""Generates tests/golden/upstream_*.npz from the REAL kornia / scikit-image / sewar — for a person with network access.

The builder image has no network and no wheels for these packages, so every restatement in oracle/ is RECALLED and
parity is unpinned.  This script is the way out that does not depend on the builder: run it on any
machine where the packages the reference pins can be installed,

    pip download --no-deps kornia==0.8.2 scikit-image==0.26.0 sewar==0.4.6 -d /tmp/w   # or point --wheels at your cache
    python tests/golden/make_upstream_golden.py --wheels /tmp/w

It (1) verifies every artefact it finds in --wheels against the sha256 digests of the reference's lock file
(/root/reference/uv.lock:227-229 kornia sdist + wheel, :632 scikit-image sdist, :700 sewar sdist — transcribed below,
because the lock file does not travel with this repository), (2) checks the imported versions, (3) writes the outputs of
the real functions on this repository's seeded synthetic inputs (mie_b200.synthetic, BASELINE.json configs) to
tests/golden/upstream_kornia.npz / upstream_skimage.npz / upstream_sewar.npz.  tests/test_upstream_golden.py then holds
the oracle (CPU suite) and the CUDA path (GPU suite) to those vectors on machines that have neither network nor wheels.
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# sha256 digests transcribed from the reference's uv.lock (file:line in the docstring above)
PINNED = {
    "kornia-0.8.2.tar.gz": "5411b2ce0dd909d1608016308cd68faeef90f88c47f47e8ecd40553fd4d8b937",
    "kornia-0.8.2-py2.py3-none-any.whl": "32dfe77c9c74a87a2de49395aa3c2c376a1b63c27611a298b394d02d13905819",
    "scikit_image-0.26.0.tar.gz": "f5f970ab04efad85c24714321fcc91613fcb64ef2a892a13167df2f3e59199fa",
    "sewar-0.4.6.tar.gz": "342cfd007a7ae99b252a6459d6e586744e8787c1b1ec51dae88f179916db3b83",
}
VERSIONS = {"kornia": "0.8.2", "skimage": "0.26.0"}


def verify_artifacts(wheel_dir: str) -> list[str]:
    """sha256 of every pinned artefact present in wheel_dir; raises on a mismatch.  Platform wheels of scikit-image are
    not in PINNED (the lock lists cp314 wheels only): they are reported as unverified."""
    seen = []
    for path in sorted(glob.glob(os.path.join(wheel_dir, "*"))):
        name = os.path.basename(path)
        if name in PINNED:
            h = hashlib.sha256(open(path, "rb").read()).hexdigest()
            if h != PINNED[name]:
                raise SystemExit(f"{name}: sha256 {h} does not match the reference lock file ({PINNED[name]})")
            seen.append(f"{name}: sha256 verified against uv.lock")
        elif name.endswith((".whl", ".tar.gz")):
            seen.append(f"{name}: not pinned by the lock file (unverified)")
    return seen


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wheels", default=None, help="directory holding the downloaded artefacts to verify")
    args = ap.parse_args()
    notes = verify_artifacts(args.wheels) if args.wheels else ["no --wheels directory given: artefacts unverified"]
    for n in notes:
        print(n)
    import torch

    from mie_b200 import synthetic

    x2 = synthetic.phantom((4, 1, 512, 512), np.uint16, seed=0)
    t = torch.from_numpy(x2.astype(np.float32) / np.float32(65535.0))
    meta = dict(notes=np.array(notes))
    try:
        import kornia as K

        if K.__version__ != VERSIONS["kornia"]:
            print(f"warning: kornia {K.__version__} imported, the reference pins {VERSIONS['kornia']}")
        g = K.filters.gaussian_blur2d(t, (9, 9), (1.0, 1.0))
        c = K.enhance.equalize_clahe(g, 2.0, (8, 8))
        u = K.filters.unsharp_mask(c, (9, 9), (1.0, 1.0))
        np.savez_compressed(
            os.path.join(HERE, "upstream_kornia.npz"), version=np.array(K.__version__), seed=np.array(0),
            clahe_c1=K.enhance.equalize_clahe(t[:1], 2.0, (8, 8)).numpy(), gauss=g.numpy()[:1], clahe_of_gauss=c.numpy()[:1],
            chain_u16=torch.round(u.clamp(0, 1) * 65535.0).numpy().astype(np.uint16),
            median3=K.filters.median_blur(t[:1], (3, 3)).numpy(), equalize=K.enhance.equalize(t[:1]).numpy(),
            bilateral=K.filters.bilateral_blur(t[:1, :, :256, :256], (9, 9), 0.1, (1.5, 1.5)).numpy(), **meta)
        print("wrote upstream_kornia.npz")
    except ImportError as e:
        print("kornia not importable:", e)
    try:
        import skimage
        from skimage import exposure, filters, restoration

        if skimage.__version__ != VERSIONS["skimage"]:
            print(f"warning: scikit-image {skimage.__version__} imported, the reference pins {VERSIONS['skimage']}")
        img = x2[0, 0]
        vol = synthetic.phantom_volume((16, 128, 128), np.int16, seed=0)
        np.savez_compressed(
            os.path.join(HERE, "upstream_skimage.npz"), version=np.array(skimage.__version__),
            adapthist=exposure.equalize_adapthist(img), equalize_hist=exposure.equalize_hist(img),
            bilateral=restoration.denoise_bilateral(np.ascontiguousarray(img[128:256, 128:256])),
            median3d=filters.median(vol),
            nlm=restoration.denoise_nl_means(img[:128, :128] / 65535.0, 7, 11, 0.1, fast_mode=True), **meta)
        print("wrote upstream_skimage.npz")
    except ImportError as e:
        print("scikit-image not importable:", e)
    try:
        import sewar

        a, b = x2[0, 0], x2[1, 0]
        np.savez_compressed(os.path.join(HERE, "upstream_sewar.npz"), mse=np.array(sewar.full_ref.mse(a, b)),
                            psnr=np.array(sewar.full_ref.psnr(a, b)), ssim=np.array(sewar.full_ref.ssim(a, b)[0]), **meta)
        print("wrote upstream_sewar.npz")
    except ImportError as e:
        print("sewar not importable:", e)


if __name__ == "__main__":
    main()
