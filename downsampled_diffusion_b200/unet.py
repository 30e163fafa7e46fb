"""U-Net epsilon-predictor: parameter container with the reference's module tree and state_dict keys.

Mirrors models/unet/unet.py:9-104 and models/unet/blocks.py of the reference: same constructor
(`Unet(config)`), same parameter names / shapes / registration (and therefore RNG-init) order, so a
reference checkpoint loads with `load_state_dict` and `model.parameters()` zips with the reference's.
The modules below hold parameters only; `Unet.forward` runs on the GPU through libddb200
(engine.UnetEngine) -- none of the nn.Conv2d / nn.GroupNorm forwards is ever called.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch import nn

from . import engine as _engine


class SinusoidalPosEmb(nn.Module):
    """blocks.py:17-29 (no parameters; evaluated inside dd_time_bias)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim


class Block(nn.Module):
    """blocks.py:74-84: Conv2d(3x3, pad 1) -> GroupNorm(groups) -> Mish."""

    def __init__(self, dim: int, dim_out: int, groups: int = 8):
        super().__init__()
        self.block = nn.Sequential(nn.Conv2d(dim, dim_out, 3, padding=1), nn.GroupNorm(groups, dim_out), nn.Mish())


class ResnetBlock(nn.Module):
    """blocks.py:87-115."""

    def __init__(self, dim: int, dim_out: int, *, time_emb_dim: int, dropout: float = 0.0, groups: int = 8):
        super().__init__()
        self.mlp = nn.Sequential(nn.Mish(), nn.Linear(time_emb_dim, dim_out))
        self.block1 = Block(dim, dim_out, groups)
        self.block2 = Block(dim_out, dim_out, groups)
        self.dropout = nn.Dropout(p=dropout)
        self.res_conv = nn.Conv2d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class LayerNorm(nn.Module):
    """blocks.py:50-60: channel LayerNorm with (1,C,1,1) gain/bias, eps added to the std."""

    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1, 1))


class LinearAttention(nn.Module):
    """blocks.py:118-134."""

    def __init__(self, dim: int, heads: int = 4, dim_head: int = 32):
        super().__init__()
        self.heads = heads
        self.dim_head = dim_head
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


class PreNorm(nn.Module):
    """blocks.py:63-71."""

    def __init__(self, dim: int, fn: nn.Module):
        super().__init__()
        self.fn = fn
        self.norm = LayerNorm(dim)


class Residual(nn.Module):
    """blocks.py:8-14."""

    def __init__(self, fn: nn.Module):
        super().__init__()
        self.fn = fn


class Downsample(nn.Module):
    """blocks.py:41-47: Conv2d(dim, dim, 3, stride 2, pad 1)."""

    def __init__(self, dim: int):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, 2, 1)


class Upsample(nn.Module):
    """blocks.py:32-38: ConvTranspose2d(dim, dim, 4, stride 2, pad 1)."""

    def __init__(self, dim: int):
        super().__init__()
        self.conv = nn.ConvTranspose2d(dim, dim, 4, 2, 1)


class Unet(nn.Module):
    """`Unet(config)` with keys unet_chan, unet_in, unet_dims, unet_dropout (unet.py:19-22).

    forward(x:(B,unet_in,h,w) fp32 CUDA, time:(B,) int64) -> (B,unet_in,h,w) fp32.
    `precision`: 'bf16' (tcgen05 implicit-GEMM path, default) or 'fp32' (validation mode, CUDA-core kernels).
    """

    def __init__(self, config: dict):
        super().__init__()
        dim = config["unet_chan"]
        in_channels = config["unet_in"]
        mults = tuple(config["unet_dims"])
        dropout = config["unet_dropout"]
        self.dim, self.in_channels, self.dim_mults, self.p_dropout = dim, in_channels, mults, dropout
        self.precision = config.get("precision", "bf16")

        dims = [in_channels] + [dim * m for m in mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        n_res = len(in_out)

        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, dim * 4), nn.Mish(), nn.Linear(dim * 4, dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        for i, (d_in, d_out) in enumerate(in_out):
            last = i >= n_res - 1
            self.downs.append(nn.ModuleList([
                ResnetBlock(d_in, d_out, time_emb_dim=dim, dropout=dropout),
                ResnetBlock(d_out, d_out, time_emb_dim=dim, dropout=dropout),
                Residual(PreNorm(d_out, LinearAttention(d_out))),
                Downsample(d_out) if not last else nn.Identity(),
            ]))
        mid = dims[-1]
        self.mid_block1 = ResnetBlock(mid, mid, time_emb_dim=dim)
        self.mid_attn = Residual(PreNorm(mid, LinearAttention(mid)))
        self.mid_block2 = ResnetBlock(mid, mid, time_emb_dim=dim)
        for i, (d_in, d_out) in enumerate(reversed(in_out[1:])):
            last = i >= n_res - 1          # never true (unet.py:60): every up stage ends in an Upsample
            self.ups.append(nn.ModuleList([
                ResnetBlock(d_out * 2, d_in, time_emb_dim=dim),
                ResnetBlock(d_in, d_in, time_emb_dim=dim),
                Residual(PreNorm(d_in, LinearAttention(d_in))),
                Upsample(d_in) if not last else nn.Identity(),
            ]))
        self.final_conv = nn.Sequential(Block(dim, dim), nn.Conv2d(dim, in_channels, 1))
        self._engines = _engine.EngineCache()

    # ---- execution -------------------------------------------------------------------------
    def engine(self, B: int, H: int, W: int, precision: str = None) -> "_engine.UnetEngine":
        """The compiled step program for this (batch, resolution, precision); built once and cached."""
        precision = precision or self.precision
        key = (B, H, W, precision)
        eng = self._engines.get(key)
        if eng is None:
            eng = _engine.UnetEngine(self, B, H, W, precision)
            self._engines[key] = eng
        return eng

    def invalidate(self) -> None:
        """Drop packed-weight caches (call after an in-place parameter update outside autograd's view)."""
        for e in self._engines.values():
            e.weights_version = None

    def forward(self, x: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
        assert x.dim() == 4 and x.shape[1] == self.in_channels, f"expected (B,{self.in_channels},h,w), got {tuple(x.shape)}"
        if not x.is_cuda:
            raise RuntimeError("downsampled_diffusion_b200.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, _, H, W = x.shape
        eng = self.engine(B, H, W)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import unet_apply
            return unet_apply(self, eng, x, time)
        if self.training and self.p_dropout > 0:
            raise RuntimeError("train-mode dropout is only available through the autograd path")
        return eng.forward(x.contiguous().float(), time)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate()
