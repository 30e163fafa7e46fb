"""Generate tests/golden/golden_v3.npz from the REAL reference: FULL T=1000 ancestral chains at the BASELINE sizes.

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:
    python oracle/make_golden_chain.py
The unmodified reference is driven through its public `sample()` (models/diffusion/ddpm.py:229-254,
dddpm.py:76-90); the only interposition is on `torch.randn`, which hands out the pre-drawn chain noise of
tests/common.chain_noise in call order (start image, then one draw per step) so the GPU path -- whose generator
cannot reproduce a CPU stream -- sees the very same tensors.  Stored per configuration: the final latents of the
checked rows and the up-sampled images (C3: every second pixel, to keep the fixture small).

Also stored: one full-size (256x256) training objective of the reference with gradient norms (C4 shape, 2 rows).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference   # noqa: E402


class _Handout:
    """torch.randn stand-in: returns the pre-drawn tensors in order, checking the requested shape."""

    def __init__(self, seq):
        self.seq, self.k = seq, 0

    def __call__(self, *shape, **kw):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        t = self.seq[self.k]
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
        self.k += 1
        return t


def main():
    ref = import_reference()
    import downsampled_diffusion_b200 as ours
    from tests import common as tc

    torch.set_num_threads(os.cpu_count() or 8)
    out = {}

    def ref_model(cfg, kind):
        mine = tc.build_model(cfg, ours, kind, 0)
        theirs = tc.build_model(cfg, ref, kind, 0)
        sd_m, sd_t = mine.state_dict(), theirs.state_dict()
        for k in sd_m:
            assert torch.equal(sd_m[k], sd_t[k]), f"init differs from the reference at {k}"
        return theirs.eval()

    real_randn = torch.randn
    for tag, cfg, hw in (("c2", tc.C2, 16), ("c3", tc.C3, 32)):
        m = ref_model(cfg, "dddpm_ae")
        rows = tc.CHAIN_ROWS
        noise = tc.chain_noise(tag, cfg["T"], rows, cfg["unet_in"], hw, hw)
        h = _Handout(noise)
        t0 = time.time()
        torch.randn = h
        try:
            with torch.no_grad():
                x, z = m.sample(rows)
        finally:
            torch.randn = real_randn
        assert h.k == cfg["T"] + 1
        out[f"fullchain.{tag}.z"] = z.numpy()
        out[f"fullchain.{tag}.x"] = (x[:, :, ::2, ::2] if tag == "c3" else x).numpy()
        print(f"chain {tag}: {time.time() - t0:.1f} s, z range [{float(z.min()):.3f}, {float(z.max()):.3f}]", flush=True)

    # ---- full-size training objective (C4 shape, 2 rows) ---------------------------------------------
    m = ref_model(tc.C3, "dddpm_ae")
    m.train()
    x = tc.rand_pm1(31, 2, 3, 256, 256)
    t = torch.tensor([50, 700])
    eps = tc.randn(32, 2, 8, 32, 32)
    torch.randn_like, real_rl = (lambda ref_t: eps), torch.randn_like
    try:
        obj, d = m.losses(x, t)
    finally:
        torch.randn_like = real_rl
    obj.backward()
    out["fulltrain.c4.obj"] = obj.detach().numpy()
    out["fulltrain.c4.latent"] = d["latent"].detach().numpy()
    out["fulltrain.c4.recon"] = d["recon"].detach().numpy()
    out["fulltrain.c4.grad_norms"] = np.asarray([0.0 if p.grad is None else float(p.grad.double().norm()) for p in m.parameters()])
    for n in ("upsample.conv.1.c2.weight", "downsample.conv.0.weight"):
        out[f"fulltrain.c4.grad.{n}"] = dict(m.named_parameters())[n].grad.numpy()
    n = "latent_model.mid_block1.block1.block.0.weight"          # (256,256,3,3): a corner keeps the fixture small
    out[f"fulltrain.c4.grad.{n}[:32,:32]"] = dict(m.named_parameters())[n].grad[:32, :32].numpy().copy()
    print("training objective done", float(obj))

    path = os.path.join(ROOT, "tests", "golden", "golden_v3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
