"""Output formatting of the sampler's caller: `fix_samples` of utils/eval_helpers.py:37-41, which
generate_model_samples.py:44-51 applies to every batch before np.save.  One kernel (per-image min / max, scale to
[0, 255], NCHW -> NHWC) and one device-to-host copy into pinned memory instead of five ATen passes + np.moveaxis."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def fix_samples(samples: torch.Tensor, out: torch.Tensor = None, non_blocking: bool = False) -> np.ndarray:
    """(B, C, H, W) device tensor -> (B, H, W, C) float32 numpy array in [0, 255], min-max normalised per image.

    out: optional pinned host tensor (B, H, W, C) to receive the copy (reused across batches by a sampling loop);
    with non_blocking=True the caller synchronises before reading the returned array."""
    dev = ops.fix_samples_raw(samples)
    if out is None:
        out = torch.empty(dev.shape, dtype=torch.float32, pin_memory=True)
    out.copy_(dev, non_blocking=non_blocking)
    return out.numpy()
