"""dif_pan_b200 — B200-native (sm_100a) implementation of the DDIF / Dif-PAN denoising hot path.

Drop-in classes with the reference's module signatures:
    UNetSR3                      (models/sr3_dwt.py)
    GaussianDiffusion            (diffusion/diffusion_ddpm_pan.py)
    NoiseScheduleVP, model_wrapper, DPM_Solver   (solver/dpm_solver.py)
plus the Haar DWT / fused cond assembly (dataset/*.py, diffusion_engine.py:221-228), the scene driver with tiling / stitching
(diffusion_engine.py:351-505), on-device validation metrics (utils/_metric_legacy.py) and the patch sharding helper.
The inference path executes hand-written CUDA kernels through the C ABI in include/ddif_b200.h; there is no fallback.
Training (first slice): `training` (autograd graph with all dense convolutions forward / dgrad / wgrad on the CUDA kernels) and `ddp`
(bucketed gradient all-reduce overlapped with backward).
"""
from .unet import UNetSR3  # noqa: F401
from .diffusion import GaussianDiffusion, make_beta_schedule, fuse_output, device_randn  # noqa: F401
from .dpm_solver import NoiseScheduleVP, model_wrapper, DPM_Solver, interpolate_fn  # noqa: F401
from .wavelet import haar_dwt2, haar_idwt2, wavelet_channels, assemble_cond, make_cond  # noqa: F401
from .scene import tile_scene, stitch_tiles, sample_cond, fuse_scene  # noqa: F401
from . import metrics, optim, training, ddp  # noqa: F401

__all__ = ["UNetSR3", "GaussianDiffusion", "make_beta_schedule", "fuse_output", "device_randn", "NoiseScheduleVP",
           "model_wrapper", "DPM_Solver", "interpolate_fn", "haar_dwt2", "haar_idwt2", "wavelet_channels", "assemble_cond", "make_cond",
           "tile_scene", "stitch_tiles", "sample_cond", "fuse_scene", "metrics", "optim", "training", "ddp"]
