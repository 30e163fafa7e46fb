"""Shared test inputs: configurations, seeded weights and tensors.

Used by tests/ AND by oracle/make_golden.py, so the golden vectors (generated once from the real
reference in the authoring container) and the tests see byte-identical weights and inputs.  Weights
come from the package's own module constructors under torch.manual_seed (bit-identical to the
reference's init, checked by make_golden.py) plus a seeded perturbation of every 1-D parameter so
GroupNorm/LayerNorm gains and biases are exercised.
"""
from __future__ import annotations

import torch

BASE = dict(T=1000, loss_type="simple", loss_flat="sum", beta_schedule="linear", unet_dropout=0.0,
            t_rec_max=100, force_latent=True, d_mode="convolutional_res", u_mode="convolutional_res",
            d_chans=64, d_dropout=0, d_n_blocks=3, u_n_blocks=3, ae_loss=True)

# C1: standard DDPM on 1x28x28 (default dims cannot run at 28x28)
C1 = dict(BASE, image_size=28, unet_chan=64, unet_in=1, unet_dims=(1, 2, 2), n_downsamples=0, color_channels=1)
# C2: dDDPM x2, 3x64x64 -> latent 8x16x16
C2 = dict(BASE, image_size=64, unet_chan=128, unet_in=8, unet_dims=(1, 2, 2, 2), n_downsamples=2, color_channels=3)
# C3/C4: dDDPM x3, 3x256x256 -> latent 8x32x32
C3 = dict(BASE, image_size=256, unet_chan=128, unet_in=8, unet_dims=(1, 2, 2, 2), n_downsamples=3, color_channels=3)
# a small dDDPM used for fast full-model checks: 3x32x32 -> latent 8x8x8, 64-channel U-Net
CS = dict(BASE, image_size=32, unet_chan=64, unet_in=8, unet_dims=(1, 2), n_downsamples=2, color_channels=3)


def build_model(cfg: dict, pkg, kind: str, seed: int = 0, device: str = "cpu"):
    """kind: 'unet' | 'ddpm' | 'dddpm' | 'dddpm_ae'; pkg: module exposing Unet/DDPM/... (ours or the reference)."""
    torch.manual_seed(seed)
    net = pkg.Unet(cfg)
    if kind == "unet":
        model = net
    elif kind == "ddpm":
        model = pkg.DDPM(cfg, net, device, cfg["color_channels"])
    elif kind == "dddpm":
        model = pkg.DownsampleDDPM(cfg, net, device, cfg["color_channels"])
    elif kind == "dddpm_ae":
        model = pkg.DownsampleDDPMAutoencoder(cfg, net, device, cfg["color_channels"])
    else:
        raise ValueError(kind)
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 or name.endswith(".g") or name.endswith(".b"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
    return model


def randn(seed: int, *shape) -> torch.Tensor:
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def rand_pm1(seed: int, *shape) -> torch.Tensor:
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed)) * 2 - 1


def eval_images(seed: int, *shape) -> torch.Tensor:
    """Evaluation inputs: values on the uint8 grid rescaled to [-1, 1] (what losses.py:66-109 assumes), so that the
    open bins beyond +-0.999 of the discretised likelihood occur."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, shape, generator=g).float() / 127.5 - 1.0


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double().cpu() - b.double().cpu()).abs().max())


# ---- full T-step chains at the BASELINE sizes (tests/test_gpu_chain_full.py, oracle/make_golden_chain.py) ----------
CHAIN_ROWS = 3                                   # rows of the benchmarked batch that the CPU side follows
CHAIN_SEEDS = {"c1": 7101, "c2": 7102, "c3": 7103}
C1_CHAIN_BATCH = 16          # BASELINE configs[0]: the standard DDPM on 1x28x28, T = 1000, batch 16


def chain_noise(tag: str, T: int, rows: int, C: int, H: int, W: int) -> torch.Tensor:
    """Pre-drawn chain noise (T+1, rows, C, H, W) of the checked rows: entry 0 is the start image, entry 1+k the
    k-th step's z.  CPU generator, so the GPU test, the oracle and the reference-side generator see the same bits."""
    g = torch.Generator().manual_seed(CHAIN_SEEDS[tag])
    return torch.randn(T + 1, rows, C, H, W, generator=g)
