"""Downsampled DDPM: diffusion on a learned low-resolution latent.

Drop-in for models/diffusion/dddpm.py:11-177 of the reference: `DownsampleDDPM` and
`DownsampleDDPMAutoencoder(config, denoise_model, device, color_channels=3)` with the same methods
(`sample -> (x, z)`, `losses -> (obj, {'latent','recon'})`, `rescaled_downsample/upsample`,
`loss_recon`, `reconstruct`) and attributes (`downsample`, `upsample`, `dim_reduc`, `x_shape`,
`t_rec_max`, `force_latent`).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .ddpm import DDPM
from .downsampled import _ResampleNet, get_downsampling, get_upsampling


class DownsampleDDPM(DDPM):
    def __init__(self, config: dict, denoise_model: nn.Module, device: str, color_channels: int = 3):
        super().__init__(config, denoise_model, device, color_channels)
        self.t_rec_max = int(self.timesteps - 1) if config["t_rec_max"] == -1 else config["t_rec_max"]
        self.x_shape = [self.in_channels, self.image_size, self.image_size]
        self.force_latent = config["force_latent"]
        unet_in = config["unet_in"]
        self.dim_reduc = np.power(2, config["n_downsamples"]).astype(int)
        z_size = int(self.image_size / self.dim_reduc)
        self.sample_shape = [unet_in, z_size, z_size]
        assert unet_in >= self.in_channels, \
            f"Input channels to DDPM-Unet {unet_in} should be equal or larger to data color channels {self.in_channels}."
        self.downsample = get_downsampling(config, self.x_shape)
        self.upsample = get_upsampling(config, self.x_shape)
        for net in (self.downsample, self.upsample):         # 'fp32' = validation mode for the resampling nets as well
            if isinstance(net, _ResampleNet) and "precision" in config:
                net.precision = config["precision"]

    # ---- latent <-> image ----------------------------------------------------------------------
    def _resample(self, net, x: torch.Tensor) -> torch.Tensor:
        # tanh fused into the last 1x1 conv ('convolutional*' modes) / applied by the bicubic resampler ('deterministic')
        return net(x, tanh=bool(self.force_latent))

    def rescaled_downsample(self, x: torch.Tensor) -> torch.Tensor:
        """dddpm.py:92-101."""
        z = self._resample(self.downsample, x)
        assert list(z.shape)[1:] == self.sample_shape, f"mismatch between {list(z.shape)[1:]} and {self.sample_shape}"
        return z

    def rescaled_upsample(self, z: torch.Tensor) -> torch.Tensor:
        """dddpm.py:103-112."""
        x = self._resample(self.upsample, z)
        assert list(x.shape)[1:] == self.x_shape, f"mismatch between {list(x.shape)[1:]} and {self.x_shape}"
        return x

    # ---- sampling --------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, batch_size: int = 16, every: int = 1, early_stop: int = None, noise=None) -> tuple:
        """dddpm.py:76-90: latent chain, then one up-net pass.  Returns (x_sample, z_sample)."""
        z_sample = self.p_sample_loop((batch_size, *self.sample_shape), every, early_stop, noise=noise)
        x_sample = self.rescaled_upsample(z_sample)
        assert list(z_sample.shape)[1:] == self.sample_shape
        assert list(x_sample.shape)[1:] == self.x_shape
        return x_sample, z_sample

    @torch.no_grad()
    def reconstruct(self, x: torch.Tensor, n: int) -> tuple:
        """dddpm.py:35-74."""
        assert x.shape[0] >= n, f"batch size ({x.shape[0]}) is below {n}"
        x = x[:n]
        t = torch.linspace(0, self.timesteps - 1, n, device=x.device, dtype=torch.long)
        z = self.rescaled_downsample(x)
        eps = torch.randn_like(z)
        z_t = self.q_sample(z, t, eps)
        eps_hat = self.latent_model(z_t, t)
        z_recon = self.predict_x_from_eps(z_t, t, eps_hat, clip=False)
        x_recon = self.rescaled_upsample(z_recon)
        assert list(x_recon.shape)[1:] == self.x_shape
        return x_recon, z_recon

    @torch.no_grad()
    def test_losses(self, x: torch.Tensor, noise=None) -> dict:
        """dddpm.py:145-148: the variational bound of the latent chain, evaluated on z = downsample(x)."""
        return self.test_losses_(self.rescaled_downsample(x), noise=noise)

    # ---- training objective ------------------------------------------------------------------------
    def loss_recon(self, x: torch.Tensor, z_hat: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """dddpm.py:114-120: per-sample reconstruction error, zeroed for t >= t_rec_max."""
        x_hat = self.rescaled_upsample(z_hat)
        assert x_hat.shape == x.shape, f"mismatch between {x_hat.shape} and {x.shape}"
        loss = self.mse_rows(x, x_hat)
        return torch.where(t < self.t_rec_max, loss, torch.zeros_like(loss))

    def losses(self, x: torch.Tensor, t: torch.Tensor, eps: torch.Tensor = None) -> tuple:
        """dddpm.py:122-143 (recon loss through the U-Net's x0 prediction)."""
        z = self.rescaled_downsample(x)
        if eps is None:
            eps = torch.randn_like(z)
        z_t = self.q_sample(z, t, eps)
        eps_hat = self.latent_model(z_t, t)
        L_ddpm = self.loss_ddpm(eps, eps_hat, t)
        z_hat = self.predict_x_from_eps(z_t, t, eps_hat, clip=False)
        L_rec = self.loss_recon(x, z_hat, t)
        obj = (L_ddpm + L_rec).mean()
        return obj, {"latent": L_ddpm.mean(), "recon": L_rec.mean()}


class DownsampleDDPMAutoencoder(DownsampleDDPM):
    def losses(self, x: torch.Tensor, t: torch.Tensor, eps: torch.Tensor = None) -> tuple:
        """dddpm.py:155-177: recon loss on the (detached-afterwards) latent; U-Net trained on L_ddpm only."""
        z = self.rescaled_downsample(x)
        L_rec = self.loss_recon(x, z, t)
        z = z.detach()
        if eps is None:
            eps = torch.randn_like(z)
        z_t = self.q_sample(z, t, eps)
        eps_hat = self.latent_model(z_t, t)
        L_ddpm = self.loss_ddpm(eps, eps_hat, t)
        obj = (L_ddpm + L_rec).mean()
        return obj, {"latent": L_ddpm.mean(), "recon": L_rec.mean()}

