"""Margin of the C1 bf16 evaluation-chain test (tests/test_gpu_eval.py): worst relative deviation per quantity, three fresh models."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
from tests.test_gpu_eval import eval_noise

g = {}
for name in ("golden_v1.npz", "golden_v2.npz"):
    with np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name)) as z:
        g.update({k: z[k] for k in z.files})
dev = torch.device("cuda:0")
for rep in range(3):
    m = tc.build_model(dict(tc.C1, T=50, precision="bf16"), dd, "ddpm", device="cuda").to(dev).eval()
    x = tc.eval_images(73, 2, 1, 28, 28)
    noise = eval_noise(9, x, 50)
    got = m.test_losses(x.to(dev), noise=noise.to(dev))
    out = []
    for k in ("vlb_t", "prior", "vlb", "L_simple_t", "L_simple"):
        ref = np.asarray(g[f"eval.c1.test_losses.{k}"]); a = got[k].detach().cpu().numpy()
        if k == "vlb_t":
            out.append("KL %.2e" % np.max(np.abs(a[:, :-1] - ref[:, :-1]) / np.abs(ref[:, :-1])))
            out.append("L0 %.2e" % np.max(np.abs(a[:, -1] - ref[:, -1]) / np.abs(ref[:, -1])))
        else:
            out.append("%s %.2e" % (k, np.max(np.abs(a - ref) / np.abs(ref))))
    print("rep", rep, " ".join(out), "(bar 6e-2)")
