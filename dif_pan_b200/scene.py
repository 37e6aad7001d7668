"""Scene driver: the caller of the denoising path — `test_fn` of /root/reference/diffusion_engine.py:351-505
without the file I/O (h5py / savemat are not part of the path).

    raw lms, pan (digital numbers)  ->  cond (one fused kernel, wavelet.make_cond)  ->  sampling  ->  clip(sample + lms, 0, 1) * division

Two ways through the sampler:
  * whole scene (what the reference does, :441-447): the UNet is fully convolutional, so a 256x256 / 512x512 scene is ONE sample whose
    GroupNorm statistics and self-attention span the scene; the CUDA plan is built for that (B, H, W).  Parity is against the reference
    run on the same scene.
  * tiled (BASELINE configs[2]: "512x512 scenes tiled into patches"): the scene's cond is cut into `patch` x `patch` tiles with stride
    `patch - overlap`, the tiles are flattened into the batch (and can be sharded over GPUs like any patch batch, sharding.py), sampled,
    and stitched back with uniform averaging of the overlaps (`ddif_tile_f32`).  Tiling changes the GroupNorm statistics (per tile
    instead of per scene), so tiled output differs from whole-scene output by construction; its parity is against the reference run
    on the same tiles.
No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .diffusion import GaussianDiffusion, make_beta_schedule, fuse_output
from .dpm_solver import NoiseScheduleVP, model_wrapper, DPM_Solver
from .wavelet import make_cond


def _grid(size: int, patch: int, stride: int) -> int:
    if patch > size or stride < 1 or stride > patch or (size - patch) % stride:
        raise ValueError(f"tiles of {patch} with stride {stride} do not cover {size} exactly")
    return (size - patch) // stride + 1


def tile_scene(x: torch.Tensor, patch: int = 64, overlap: int = 0) -> torch.Tensor:
    """[B, C, H, W] -> [B*ny*nx, C, patch, patch] (row-major tile order), one gather kernel."""
    if not x.is_cuda:
        raise RuntimeError("dif_pan_b200.scene runs on CUDA only (no CPU fallback)")
    x = x.to(torch.float32).contiguous()
    B, C, H, W = x.shape
    stride = patch - overlap
    ny, nx = _grid(H, patch, stride), _grid(W, patch, stride)
    tiles = torch.empty(B * ny * nx, C, patch, patch, dtype=torch.float32, device=x.device)
    _lib.launch("ddif_tile_t", _lib.stream_of(x.device), scene=x.data_ptr(), tiles=tiles.data_ptr(), batch=B, c=C,
                h=H, w=W, ph=patch, pw=patch, sy=stride, sx=stride, ny=ny, nx=nx, dir=0)
    return tiles


def stitch_tiles(tiles: torch.Tensor, scene_hw: Tuple[int, int], overlap: int = 0) -> torch.Tensor:
    """Inverse of tile_scene; pixels covered by several tiles get their uniform average."""
    if not tiles.is_cuda:
        raise RuntimeError("dif_pan_b200.scene runs on CUDA only (no CPU fallback)")
    tiles = tiles.to(torch.float32).contiguous()
    n, C, ph, pw = tiles.shape
    H, W = scene_hw
    ny, nx = _grid(H, ph, ph - overlap), _grid(W, pw, pw - overlap)
    if n % (ny * nx):
        raise ValueError(f"{n} tiles are not a whole number of {ny}x{nx} scenes")
    B = n // (ny * nx)
    scene = torch.empty(B, C, H, W, dtype=torch.float32, device=tiles.device)
    _lib.launch("ddif_tile_t", _lib.stream_of(tiles.device), scene=scene.data_ptr(), tiles=tiles.data_ptr(), batch=B,
                c=C, h=H, w=W, ph=ph, pw=pw, sy=ph - overlap, sx=pw - overlap, ny=ny, nx=nx, dir=1)
    return scene


def sample_cond(net, cond: torch.Tensor, channels: int, sampler: str = "ddim25", n_timestep: int = 500, noise=None, x_T=None,
                seed: int = 0) -> torch.Tensor:
    """One sampling of a cond batch -> residual sample [B, channels, H, W].
    sampler: 'ddpm' | 'ddimN' (GaussianDiffusion, diffusion_engine.py:441-444) | 'dpmN' (DPM-Solver++ 2M, N steps, time_uniform)."""
    betas = make_beta_schedule("cosine", n_timestep, cosine_s=8e-3)
    dif = GaussianDiffusion(net, image_size=cond.shape[-1], channels=channels, pred_mode="x_start", loss_type="l1", device=str(cond.device),
                            clamp_range=(0, 1))
    dif.set_new_noise_schedule(betas=betas, device=cond.device)
    dif.seed = seed
    if sampler == "ddpm":
        return dif(cond, mode="ddpm_sample", noise=noise)
    if sampler.startswith("ddim"):
        return dif(cond, mode="ddim_sample", section_counts=sampler, noise=noise)
    if sampler.startswith("dpm"):
        steps = int(sampler[3:])
        ns = NoiseScheduleVP("discrete", betas=dif.betas.cpu())
        mfn = model_wrapper(net, ns, model_type="x_start", guidance_type="classifier-free", condition=cond, guidance_scale=1.0)
        if x_T is None:
            from .diffusion import device_randn
            x_T = device_randn((cond.shape[0], channels, cond.shape[2], cond.shape[3]), cond.device, seed, 0)
        return DPM_Solver(mfn, ns, algorithm_type="dpmsolver++").sample(x_T, steps=steps, order=2, skip_type="time_uniform", method="multistep")
    raise ValueError(f"unknown sampler {sampler!r}")


def fuse_scene(net, lms_dn: torch.Tensor, pan_dn: torch.Tensor, division: float, order: str = "pan", sampler: str = "ddim25",
               n_timestep: int = 500, patch: Optional[int] = None, overlap: int = 0, tile_batch: int = 256, seed: int = 0,
               x_T=None) -> torch.Tensor:
    """raw lms [B,C,H,W], pan [B,P,H,W] -> fused image in digital numbers, clipped to [0, division] (diffusion_engine.py:441-466).
    patch=None: whole-scene sampling like the reference; patch=64: tiled (see module docstring), `tile_batch` tiles per UNet batch."""
    C = lms_dn.shape[1]
    cond = make_cond(lms_dn, pan_dn, division, order)
    if patch is None:
        sample = sample_cond(net, cond, C, sampler, n_timestep, seed=seed, x_T=x_T)
        sr = fuse_output(sample, cond)
    else:
        tiles = tile_scene(cond, patch, overlap)
        outs = []
        for i in range(0, tiles.shape[0], tile_batch):
            c = tiles[i:i + tile_batch].contiguous()
            xt = None if x_T is None else x_T[i:i + tile_batch].contiguous()
            outs.append(fuse_output(sample_cond(net, c, C, sampler, n_timestep, seed=seed + i, x_T=xt), c))
        sr = stitch_tiles(torch.cat(outs, 0), tuple(lms_dn.shape[-2:]), overlap)
    return (sr * float(division)).clamp_(0, float(division))
