"""Run N graph replays of the C1 sampling step (standard DDPM, 1x28x28, batch 16; for ncu launch lists of the ragged-map path)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from tests import common as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
m = tc.build_model(dict(tc.C1, precision="bf16"), dd, "ddpm", device="cuda:0").to(dev).eval()
plan = m.sampling_plan((B, 1, 28, 28))
plan.prepare()
plan.t_dev.fill_(999)
plan.noise.normal_()
plan.eng.x_in.normal_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(N):
    plan.graph.replay()
e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1) / N, "finite", bool(torch.isfinite(plan.eng.x_in).all()))
