"""Generate tests/golden/golden_v2.npz from the REAL reference: the evaluation-side chain (SURVEY.md 8(f).3:
DDPM.vlb_terms / calc_prior / test_losses, DownsampleDDPM.test_losses), the sampler caller's output formatting
(8(f).2: utils.eval_helpers.fix_samples) and one optimizer step of the trainer (8(f).1: clip_grad_norm_ + Adam,
trainers/trainer_ddpm.py:118-144 -- torch's own implementations, which is what the reference calls).

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:
    python oracle/make_golden_eval.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference          # noqa: E402


def main():
    ref = import_reference()
    import downsampled_diffusion_b200 as ours
    from tests import common as tc
    from utils.eval_helpers import fix_samples            # the reference's (REF is first on sys.path after import_reference)

    torch.set_num_threads(8)
    out = {}

    def put(name, t):
        # a copy: .numpy() of a CPU parameter aliases it, and the optimizer below keeps writing to the parameters
        out[name] = t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.array(t, copy=True)

    def ref_model(cfg, kind, seed=0):
        mine = tc.build_model(cfg, ours, kind, seed)
        theirs = tc.build_model(cfg, ref, kind, seed)
        for (k, a), (k2, b) in zip(mine.state_dict().items(), theirs.state_dict().items()):
            assert k == k2 and torch.equal(a, b), k
        theirs.eval()
        return theirs

    # ---- vlb_terms with per-sample t (t = 0 takes the discretised-likelihood branch), calc_prior ----------
    m = ref_model(tc.C1, "ddpm")
    x = tc.eval_images(71, 4, 1, 28, 28)
    eps = tc.randn(72, 4, 1, 28, 28)
    t = torch.tensor([0, 1, 500, 999])
    with torch.no_grad():
        x_t = m.q_sample(x, t, eps)
        put("eval.c1.vlb_terms", m.vlb_terms(x, x_t, t))
        t0 = torch.zeros(4, dtype=torch.long)
        x_t0 = m.q_sample(x, t0, eps)
        put("eval.c1.vlb_terms_t0", m.vlb_terms(x, x_t0, t0))
        put("eval.c1.prior", m.calc_prior(x))

    # ---- full evaluation chains (T = 50) ------------------------------------------------------------
    m = ref_model(dict(tc.C1, T=50), "ddpm")
    x = tc.eval_images(73, 2, 1, 28, 28)
    torch.manual_seed(9)
    for k, v in m.test_losses(x).items():
        put(f"eval.c1.test_losses.{k}", v)
    m = ref_model(dict(tc.CS, T=50), "dddpm_ae")
    x = tc.eval_images(74, 2, 3, 32, 32)
    torch.manual_seed(10)
    for k, v in m.test_losses(x).items():
        put(f"eval.cs.test_losses.{k}", v)
    print("evaluation chains done")

    # ---- fix_samples -----------------------------------------------------------------------------------
    s = tc.randn(75, 3, 3, 32, 32)
    put("fix_samples.out", fix_samples(s))

    # ---- 'deterministic' (bicubic) resampler mode: wrapper.py:22-24, 49-53; loss + gradient through the up-sampler ----
    cfg = dict(tc.CS, d_mode="deterministic", u_mode="deterministic", unet_in=3)
    m = ref_model(cfg, "dddpm")
    x = tc.rand_pm1(85, 4, 3, 32, 32)
    with torch.no_grad():
        z = m.rescaled_downsample(x)
        put("det.z", z)
        put("det.xhat", m.rescaled_upsample(z))
    m.train()
    t = torch.tensor([3, 50, 99, 700])
    torch.manual_seed(12)
    obj, d = m.losses(x, t)
    obj.backward()
    put("det.loss.obj", obj)
    put("det.loss.latent", d["latent"])
    put("det.loss.recon", d["recon"])
    put("det.loss.grad_final", dict(m.named_parameters())["latent_model.final_conv.1.weight"].grad)
    put("det.loss.grad_init", dict(m.named_parameters())["latent_model.downs.0.0.block1.block.0.weight"].grad)

    # ---- one trainer step: clip_grad_norm_(1.0) + Adam(lr=2e-4), twice (trainer_ddpm.py:118-144, trainer.py:69) ----
    net = ref_model(tc.CS, "unet")
    opt = torch.optim.Adam(net.parameters(), lr=2e-4)
    names = ("final_conv.1.weight", "downs.0.0.block1.block.0.bias", "mid_attn.fn.norm.g", "time_mlp.3.weight")
    for step in range(2):
        g = torch.Generator().manual_seed(80 + step)
        for p in net.parameters():
            p.grad = (0.05 if step == 0 else 0.0005) * torch.randn(p.shape, generator=g)   # step 0 is clipped, step 1 is not
        put(f"optim.norm{step}", torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0))
        opt.step()
        sd = net.state_dict()
        for n in names:
            put(f"optim.step{step}.{n}", sd[n])

    path = os.path.join(ROOT, "tests", "golden", "golden_v2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays,", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
