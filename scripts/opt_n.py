"""N fused optimizer steps (clip + Adam + EMA) over the C4 model's 22.67 M parameters, for ncu: python scripts/opt_n.py [N]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
m = tc.build_model(dict(tc.C3, precision="bf16"), dd, "dddpm_ae", device="cuda:0").to(dev).train()
for p in m.parameters():
    p.grad = torch.randn_like(p) * 1e-3
ema = dd.EMA(m, decay=0.995)
opt = dd.Adam(m.parameters(), lr=2e-4, max_grad_norm=1.0)
opt.attach_ema(ema, m)
for _ in range(N):
    opt.step(ema="update")
torch.cuda.synchronize()
print("ok", float(opt.grad_norm))
