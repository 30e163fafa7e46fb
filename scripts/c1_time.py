"""BASELINE configs[0] (standard DDPM on 1x28x28, T = 1000, batch 16): wall time of the full chain per precision, step time by
CUDA events.  usage: python scripts/c1_time.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc

dev = torch.device("cuda:0")
for precision in ("bf16", "fp32"):
    m = tc.build_model(dict(tc.C1, precision=precision), dd, "ddpm", device="cuda").to(dev).eval()
    with torch.no_grad():
        m.sample(16)                                    # lowering + graph capture
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x = m.sample(16)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    print(f"c1 {precision}: chain of 1000 steps at batch 16: {best:.1f} ms ({16e3 / best:.1f} samples/s, {best / 1000:.4f} ms per step), finite {bool(torch.isfinite(x).all())}")
