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


@torch.no_grad()
def generate_samples(model, n_samples: int, batch_size: int, every: int = 1):
    """The sampling loop of generate_model_samples.py:41-51: ceil(n_samples / batch_size) calls of `model.sample`, each
    batch passed through `fix_samples`.  Returns (sample_list, latent_list) -- lists of (B, H, W, C) float32 arrays, the
    objects the reference hands to np.save (:61, :67); latent_list is empty for a plain DDPM.

    The formatted batch goes to pinned host memory with a non-blocking copy that overlaps the next batch's chain; the host
    waits on one event per batch only when it collects the arrays at the end."""
    n_batches = -(-n_samples // batch_size)
    pending = []
    for _ in range(n_batches):
        out = model.sample(batch_size, every)
        parts = out if isinstance(out, tuple) else (out,)
        host = []
        for t in parts:
            dev = ops.fix_samples_raw(t)
            buf = torch.empty(dev.shape, dtype=torch.float32, pin_memory=True)
            buf.copy_(dev, non_blocking=True)
            host.append(buf)
        ev = torch.cuda.Event()
        ev.record()
        pending.append((ev, host))
    sample_list, latent_list = [], []
    for ev, host in pending:
        ev.synchronize()
        sample_list.append(host[0].numpy())
        if len(host) > 1:
            latent_list.append(host[1].numpy())
    return sample_list, latent_list
