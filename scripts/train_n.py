"""Run N training steps of C4 (dDDPM x3 256x256, B images) for ncu launch lists: python scripts/train_n.py B N"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"          # "bf16" -> TF32 tensor-core training, "fp32" -> CUDA-core validation mode
m = tc.build_model(dict(tc.C3, unet_dropout=0.1, precision=prec), dd, "dddpm_ae", device="cuda:0").to(dev).train()
ema = dd.EMA(m, decay=0.995)
opt = dd.Adam(m.parameters(), lr=2e-4, max_grad_norm=1.0)
opt.attach_ema(ema, m)
x = tc.rand_pm1(1, B, 3, 256, 256).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(N):
    if i == N - 1:
        torch.cuda.synchronize(); e0.record()
    obj, _ = m(x)
    obj.backward()
    opt.step(ema="update"); opt.zero_grad()
e1.record(); torch.cuda.synchronize()
print("last step ms", e0.elapsed_time(e1), "loss", float(obj.detach()))
