"""downsampled_diffusion_b200 -- B200-native (sm_100a) hot path of simonamtoft/downsampled-diffusion.

Same Python surface as the reference for the sampling chain and the training denoising step:
`Unet`, `DDPM`, `DownsampleDDPM`, `DownsampleDDPMAutoencoder`, `get_downsampling`, `get_upsampling`, `EMA`.
All arithmetic runs in libddb200.so (hand-written CUDA, C-ABI in include/ddb200.h); there is no
CPU, cuDNN/cuBLAS or Triton fallback.
"""
from .unet import Unet
from .ddpm import DDPM
from .dddpm import DownsampleDDPM, DownsampleDDPMAutoencoder
from .downsampled import ConvResNet, SimpleDownConv, SimpleUpConv, get_downsampling, get_upsampling
from .ema import EMA
from .schedule import make_beta_schedule
from .evalfmt import fix_samples, generate_samples
from .optim import Adam
from . import parallel

__all__ = ["Unet", "DDPM", "DownsampleDDPM", "DownsampleDDPMAutoencoder", "ConvResNet", "SimpleDownConv",
           "SimpleUpConv", "get_downsampling", "get_upsampling", "EMA", "make_beta_schedule", "fix_samples", "generate_samples", "Adam"]
