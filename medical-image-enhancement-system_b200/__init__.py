"""B200-native enhancement hot path (CLAHE / equalisation, Gaussian, median,
bilateral, unsharp mask) behind kornia-compatible call signatures.

Importable as `mie_b200` (the directory name carries the reference repository's
hyphenated name, so a small shim package provides the import name).  Every
operator calls hand-written sm_100a CUDA kernels through the C ABI declared in
include/mie.h; there is no CPU or PyTorch fallback.
"""
from ._ffi import lib as _lib  # noqa: F401
from ._ffi import kernel_policy, value_range_mode
from .chain import (BilateralClahePlan, ChainConfig, ChainPlan, ChainRing, bilateral_clahe, chain_workspace_bytes,
                    enhance_chain)
from .enhance import clahe16_luts, clahe_apply, clahe_histograms, clahe_luts, equalize, equalize_clahe
from .filters import bilateral_blur, denoise_nl_means, gaussian_blur2d, get_gaussian_kernel1d, median, median_blur, unsharp_mask
from .loader import HostSlicePipeline, HostVolumePipeline, enhance_chain_host, median3d_clahe_host
from .metrics import mae, mse, psnr, rmse, ssim
from . import volume
from .volume import PeerSlabPlan, SlabPlan, exchange_z_halos, map_peer_halos, median3d_clahe_slab, shard_range, start_z_halo_exchange

__version__ = "0.1.0"

__all__ = [
    "equalize_clahe", "equalize", "clahe_histograms", "clahe_luts", "clahe_apply", "clahe16_luts",
    "gaussian_blur2d", "unsharp_mask", "median_blur", "bilateral_blur", "median", "get_gaussian_kernel1d",
    "denoise_nl_means",
    "ChainConfig", "ChainPlan", "ChainRing", "enhance_chain", "chain_workspace_bytes", "bilateral_clahe", "BilateralClahePlan",
    "HostSlicePipeline", "enhance_chain_host", "HostVolumePipeline", "median3d_clahe_host",
    "mse", "rmse", "psnr", "ssim", "mae", "value_range_mode", "kernel_policy",
    "shard_range", "exchange_z_halos", "start_z_halo_exchange", "median3d_clahe_slab", "SlabPlan", "PeerSlabPlan", "map_peer_halos",
]
