"""Down/up-sampling networks of dDDPM: parameter containers + their launch programs.

Drop-in for models/downsampled/convblocks.py:70-159 and models/downsampled/wrapper.py:6-59 of the
reference (`get_downsampling(config, shape)`, `get_upsampling(config, shape)`, same module tree and
state_dict keys: `conv.0.weight`, `conv.1.c1.weight`, ...).  The modules hold parameters; forward
runs a fixed list of libddb200 launches (engine.Program).
"""
from __future__ import annotations


import numpy as np
import torch
from torch import nn

from . import _lib as L
from .engine import Act, EngineCache, Program, _is_pow2, ensure_lazy


class ConvResBlock(nn.Module):
    """convblocks.py:92-130: x + c4(mish(c3(mish(c2(mish(c1(mish(x)))))))), then avg-pool / nearest x2."""

    def __init__(self, dim: int, in_channels: int, out_channels: int = None, upsample: bool = False,
                 downsample: bool = False, dropout: float = 0, residual: bool = False):
        super().__init__()
        assert not (upsample and downsample), "Does not make sense to both down- and upsample."
        self.upsample, self.downsample, self.residual = upsample, downsample, residual
        self.c1 = nn.Conv2d(in_channels, dim, 1)
        self.c2 = nn.Conv2d(dim, dim, 3, padding=1)
        self.c3 = nn.Conv2d(dim, dim, 3, padding=1)
        self.c4 = nn.Conv2d(dim, out_channels, 1)
        self.drop = nn.Dropout2d(p=dropout)


class _ResampleProgram(Program):
    """Launch list of a ConvResNet / SimpleDownConv / SimpleUpConv for a fixed (B, H, W)."""

    def __init__(self, net: nn.Module, B: int, C: int, H: int, W: int, precision: str, tanh: bool):
        # 'bf16' (the sampling default) runs the 32/64-channel block convolutions on the tensor cores in TF32 over fp32
        # activations (dd_conv_tc32, Mish copies written by the producing epilogue); 'fp32' is the CUDA-core validation mode.
        self.tc = precision != "fp32"
        super().__init__(net, B, "fp32")
        ensure_lazy()
        self.x_in = self.empty(B, C, H, W, dtype=torch.float32)
        layers = list(net.conv)
        x = None
        self.out = None
        for i, m in enumerate(layers):
            last = i == len(layers) - 1
            if isinstance(m, ConvResBlock):
                if m.drop.p > 0 and net.training:
                    raise RuntimeError("Dropout2d in the resampling nets is only supported with p=0 (reference default)")
                if self.tc and self._tc_ok(x, m):
                    nxt = (i + 1 < len(layers) and isinstance(layers[i + 1], ConvResBlock) and not (m.upsample or m.downsample))
                    h = self.tc_conv(x, m.c1, "1x1")
                    h = self.tc_conv(h, m.c2, "3x3")
                    h = self.tc_conv(h, m.c3, "3x3")
                    x = self.tc_conv(h, m.c4, "1x1", residual=x if m.residual else None, emit_mish=nxt)
                else:
                    h, _ = self.conv(x, m.c1, kind="1x1", pre_mish=True)
                    h, _ = self.conv(h, m.c2, kind="3x3", pre_mish=True)
                    h, _ = self.conv(h, m.c3, kind="3x3", pre_mish=True)
                    x, _ = self.conv(h, m.c4, kind="1x1", pre_mish=True, residual=x if m.residual else None)
                if m.upsample:
                    y = self.act(x.H * 2, x.W * 2, x.C, B)
                    self.add("dd_upsample_nearest2", L.ptr(x.t), L.ptr(y.t), self.dcode, B, x.H, x.W, x.C)
                    x = y
                elif m.downsample:
                    y = self.act(x.H // 2, x.W // 2, x.C, B)
                    self.add("dd_avgpool2", L.ptr(x.t), L.ptr(y.t), self.dcode, B, x.H, x.W, x.C)
                    x = y
                continue
            # plain nn.Conv2d / nn.ConvTranspose2d layer
            transposed = isinstance(m, nn.ConvTranspose2d)
            ks = m.kernel_size[0]
            stride = m.stride[0]
            Cin = m.in_channels
            Cout = m.out_channels
            first = x is None
            if first:
                Hi, Wi = H, W
            else:
                Hi, Wi = x.H, x.W
            if transposed:
                Ho, Wo, mode, pad = Hi * 2, Wi * 2, 1, 1
                wd = self.packed((16, Cin, Cout), torch.float32,
                                 lambda buf, m=m, Cin=Cin, Cout=Cout: buf.copy_(m.weight.detach().permute(2, 3, 0, 1).reshape(16, Cin, Cout)))
            else:
                pad = m.padding[0]
                Ho = (Hi + 2 * pad - ks) // stride + 1
                Wo = (Wi + 2 * pad - ks) // stride + 1
                mode = 0
                wd = self.packed((ks * ks, Cin, Cout), torch.float32,
                                 lambda buf, m=m, ks=ks, Cin=Cin, Cout=Cout: buf.copy_(m.weight.detach().permute(2, 3, 1, 0).reshape(ks * ks, Cin, Cout)))
            b_t = self.f32(m.bias) if m.bias is not None else None
            flags = (L.CONV_IN_NCHW if first else 0)
            if last:
                self.out = self.empty(B, Cout, Ho, Wo, dtype=torch.float32)
                flags |= L.CONV_OUT_NCHW | (L.CONV_TANH if tanh else 0)
                dst = self.out
                y = None
            else:
                y = self.act(Ho, Wo, Cout, B)
                dst = y.t
            thin_w = lambda c: c % 4 == 0 and 4 <= c <= 128 and (c & (c - 1)) == 0
            plain1x1 = ks == 1 and stride == 1 and not transposed
            if plain1x1 and first and not last and Cin <= 8 and thin_w(Cout):            # 3 -> 64 input layer: HBM streaming kernel
                wt = self.packed((Cin, Cout), torch.float32, lambda buf, m=m, Cin=Cin, Cout=Cout: buf.copy_(m.weight.detach().reshape(Cout, Cin).t()))
                self.add("dd_conv1x1_thin_in", L.ptr(self.x_in), L.ptr(wt), L.ptr(b_t) if b_t is not None else None, L.ptr(dst), B, Hi * Wi,
                         Cin, Cout, 0)
            elif plain1x1 and last and not first and Cout <= 8 and thin_w(Cin) and Cout <= Cin // 4:     # 64 -> 3 (+tanh) output layer
                wt = self.packed((Cout, Cin), torch.float32, lambda buf, m=m, Cin=Cin, Cout=Cout: buf.copy_(m.weight.detach().reshape(Cout, Cin)))
                self.add("dd_conv1x1_thin_out", L.ptr(x.t), L.ptr(wt), L.ptr(b_t) if b_t is not None else None, L.ptr(dst), B, Hi * Wi,
                         Cin, Cout, 1 if tanh else 0)
            else:
                self.add("dd_conv_direct", L.ptr(self.x_in) if first else L.ptr(x.t), None, Cin, 0,
                         L.DD_F32 if first else self.dcode, L.ptr(wd), L.ptr(b_t) if b_t is not None else None, None,
                         L.ptr(dst), self.dcode, B, Hi, Wi, Cout, ks, stride, pad, mode, flags)
            x = y
        if self.out is None:
            raise ValueError("resampling net must end in a plain convolution")
        self.refresh_weights()

    # ---- TF32 tensor-core form of a ConvResBlock convolution (convblocks.py:112-118: conv(mish(x))) ----------------
    @staticmethod
    def _tc_ok(x, m) -> bool:
        chans = (m.c1.in_channels, m.c1.out_channels, m.c4.out_channels)
        return all(c % 32 == 0 for c in chans) and _is_pow2(x.H) and _is_pow2(x.W) and x.H * x.W >= 16

    def tc_conv(self, x, conv: nn.Conv2d, kind: str, residual=None, emit_mish: bool = True):
        Cin, Cout, ks = conv.in_channels, conv.out_channels, conv.kernel_size[0]
        xm = x.mish
        if xm is None:                      # block input that no epilogue activated (first block, after a resampling step)
            xm = self.act(x.H, x.W, x.C, x.B)
            self.add("dd_ew", 0, L.ptr(x.t), None, L.ptr(xm.t), x.t.numel(), 1.0, 0)
        K = ks * ks * Cin
        wp = self.packed((Cout, K), torch.float32, lambda buf: buf.copy_(conv.weight.detach().permute(0, 2, 3, 1).reshape(Cout, K)))
        b_t = self.f32(conv.bias) if conv.bias is not None else None
        y = self.act(x.H, x.W, Cout, x.B)
        ym = self.act(x.H, x.W, Cout, x.B) if emit_mish else None
        self.add("dd_conv_tc32", L.TC_CONV3x3 if ks == 3 else L.TC_CONV1x1, L.ptr(xm.t), None, Cin, 0, L.ptr(wp), Cout,
                 L.ptr(b_t) if b_t is not None else None, L.ptr(residual.t) if residual is not None else None, L.ptr(y.t),
                 L.ptr(ym.t) if ym is not None else None, None, x.B, x.H, x.W, Cout)
        y.mish = ym
        return y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.refresh_weights()
        self.x_in.copy_(x)
        self.run_ops()
        return self.out.clone()


class _ResampleNet(nn.Module):
    precision = "bf16"

    def _program(self, x: torch.Tensor, tanh: bool) -> _ResampleProgram:
        if not hasattr(self, "_programs"):
            self._programs = EngineCache()
        B, C, H, W = x.shape
        key = (B, C, H, W, self.precision, tanh)
        prog = self._programs.get(key)
        if prog is None:
            prog = _ResampleProgram(self, B, C, H, W, self.precision, tanh)
            self._programs[key] = prog
        return prog

    def forward(self, x: torch.Tensor, tanh: bool = False) -> torch.Tensor:
        """NCHW fp32 in / out.  `tanh=True` fuses dddpm.py:99-100 / 110-111's squashing into the last conv."""
        if not x.is_cuda:
            raise RuntimeError("downsampled_diffusion_b200 resampling nets run on CUDA only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import resample_apply
            return resample_apply(self, x, tanh)
        return self._program(x, tanh).forward(x.contiguous().float())


class ConvResNet(_ResampleNet):
    """convblocks.py:133-159."""

    def __init__(self, dim: int, in_channels: int, out_channels: int, n_downsamples: int = 1, upsample: bool = False,
                 dropout: float = 0, n_blocks: int = 1):
        super().__init__()
        downsample = not upsample
        layers = [nn.Conv2d(in_channels, dim, 1)]
        for _ in range(n_downsamples):
            layers.append(ConvResBlock(int(dim / 2), dim, dim, upsample, downsample, dropout, residual=True))
            for _ in range(max(int(n_blocks) - 1, 0)):
                layers.append(ConvResBlock(int(dim / 2), dim, dim, False, False, dropout, residual=True))
        layers.append(nn.Conv2d(dim, out_channels, 1))
        self.conv = nn.Sequential(*layers)


class SimpleDownConv(_ResampleNet):
    """convblocks.py:70-78: a stack of stride-2 3x3 convolutions."""

    def __init__(self, dim: int = 8, in_channels: int = 3, n_downsamples: int = 1):
        super().__init__()
        dims = [in_channels] + [dim] * n_downsamples
        self.in_out = list(zip(dims[:-1], dims[1:]))
        self.conv = nn.Sequential(*[nn.Conv2d(i, o, 3, stride=2, padding=1) for i, o in self.in_out])


class SimpleUpConv(_ResampleNet):
    """convblocks.py:81-89: a stack of 4x4 stride-2 transposed convolutions."""

    def __init__(self, dim: int = 8, in_channels: int = 3, n_downsamples: int = 1):
        super().__init__()
        dims = [in_channels] + [dim] * n_downsamples
        self.in_out = list(zip(dims[:-1], dims[1:]))
        self.conv = nn.Sequential(*[nn.ConvTranspose2d(o, i, 4, stride=2, padding=1) for i, o in self.in_out[::-1]])


class _BicubicFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, tanh):
        if not x.is_cuda:
            raise RuntimeError("downsampled_diffusion_b200 runs on CUDA tensors only (no CPU fallback)")
        x = x.contiguous().float()
        B, C, H, W = x.shape
        y = torch.empty(B, C, size[0], size[1], dtype=torch.float32, device=x.device)
        L.call("dd_bicubic2d", L.ptr(x), L.ptr(y), B * C, H, W, size[0], size[1], L.stream())
        if tanh:
            L.call("dd_ew", 4, L.ptr(y), None, L.ptr(y), y.numel(), 1.0, 0, L.stream())
            ctx.save_for_backward(y)
        ctx.geom, ctx.tanh = (B, C, H, W, size[0], size[1]), tanh
        return y

    @staticmethod
    def backward(ctx, gy):
        B, C, H, W, Ho, Wo = ctx.geom
        gy = gy.contiguous().float()
        if ctx.tanh:
            (y,) = ctx.saved_tensors
            g2 = torch.empty_like(gy)
            L.call("dd_ew", 2, L.ptr(y), L.ptr(gy), L.ptr(g2), gy.numel(), 1.0, 0, L.stream())
            gy = g2
        gx = torch.empty(B, C, H, W, dtype=torch.float32, device=gy.device)
        L.call("dd_bicubic2d_bwd", L.ptr(gy), L.ptr(gx), B * C, H, W, Ho, Wo, L.stream())
        return gx, None, None


class Interpolate:
    """convblocks.py:8-26 ('deterministic' mode): `F.interpolate(size, mode='bicubic', align_corners=True)` as the
    libddb200 kernel `dd_bicubic2d` (+ its input gradient); callable like the partial the reference returns.  The
    reference never asks for another mode (wrapper.py:24, 53 call get_interpolate(size) with the defaults)."""

    def __init__(self, size: tuple):
        self.size = (int(size[0]), int(size[1]))

    def __call__(self, x: torch.Tensor, tanh: bool = False) -> torch.Tensor:
        return _BicubicFn.apply(x, self.size, bool(tanh))


def get_interpolate(size: tuple, mode: str = None, align: bool = True):
    """convblocks.py:8-26."""
    if mode not in (None, "bicubic") or not align:
        raise NotImplementedError("only the reference's default (bicubic, align_corners=True) has a kernel in libddb200")
    return Interpolate(size)


def get_upsampling(config: dict, shape: tuple):
    """wrapper.py:6-30."""
    assert shape[1] == shape[2]
    assert shape[0] == 1 or shape[0] == 3
    in_channels, mode = shape[0], config["u_mode"]
    if mode == "deterministic":
        return get_interpolate((shape[1], shape[2]))
    if mode == "convolutional":
        return SimpleUpConv(config["unet_in"], in_channels, config["n_downsamples"])
    if mode == "convolutional_res":
        return ConvResNet(config["d_chans"], config["unet_in"], in_channels, config["n_downsamples"], upsample=True,
                          dropout=config["d_dropout"], n_blocks=config["u_n_blocks"])
    raise NotImplementedError(f'Upsampling method for "{mode}" not implemented!')


def get_downsampling(config: dict, shape: tuple):
    """wrapper.py:33-59."""
    assert shape[1] == shape[2]
    assert shape[0] == 1 or shape[0] == 3
    in_channels, mode = shape[0], config["d_mode"]
    if mode == "deterministic":
        scale = np.power(2, config["n_downsamples"]).astype(int)
        size = (int(shape[1] / scale), int(shape[2] / scale))
        assert size[0] % 2 == 0, "result from downsampling should have even dimensions."
        return get_interpolate(size)
    if mode == "convolutional":
        return SimpleDownConv(config["unet_in"], in_channels, config["n_downsamples"])
    if mode == "convolutional_res":
        return ConvResNet(config["d_chans"], in_channels, config["unet_in"], config["n_downsamples"], upsample=False,
                          dropout=config["d_dropout"], n_blocks=config["d_n_blocks"])
    raise NotImplementedError(f'Downsampling method for "{mode}" not implemented!')
