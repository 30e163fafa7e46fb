"""Generate tests/golden/golden_v4.npz from the REAL reference: BASELINE configs[0] as written -- the standard DDPM
(downsample = 0) on 1x28x28 MNIST-shaped images, the full T=1000 ancestral chain at batch 16.

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:
    python oracle/make_golden_c1.py
Same recipe as oracle/make_golden_chain.py: the unmodified reference's public `sample()` (models/diffusion/ddpm.py:229-254)
with `torch.randn` handing out the pre-drawn noise of tests/common.chain_noise in call order.  Stored: the 16 final images.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference          # noqa: E402
from oracle.make_golden_chain import _Handout             # noqa: E402


def main():
    ref = import_reference()
    import downsampled_diffusion_b200 as ours
    from tests import common as tc

    torch.set_num_threads(os.cpu_count() or 8)
    cfg, B = tc.C1, tc.C1_CHAIN_BATCH
    mine, theirs = tc.build_model(cfg, ours, "ddpm", 0), tc.build_model(cfg, ref, "ddpm", 0)
    sd_m, sd_t = mine.state_dict(), theirs.state_dict()
    for k in sd_m:
        assert torch.equal(sd_m[k], sd_t[k]), f"init differs from the reference at {k}"
    theirs.eval()
    noise = tc.chain_noise("c1", cfg["T"], B, 1, 28, 28)
    h, real_randn = _Handout(noise), torch.randn
    t0 = time.time()
    torch.randn = h
    try:
        with torch.no_grad():
            x = theirs.sample(B)
    finally:
        torch.randn = real_randn
    assert h.k == cfg["T"] + 1 and tuple(x.shape) == (B, 1, 28, 28)
    print(f"chain c1: {time.time() - t0:.1f} s, range [{float(x.min()):.3f}, {float(x.max()):.3f}]", flush=True)
    path = os.path.join(ROOT, "tests", "golden", "golden_v4.npz")
    np.savez_compressed(path, **{"fullchain.c1.x": x.numpy()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
