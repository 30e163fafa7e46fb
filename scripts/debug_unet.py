import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DD_DEBUG"] = "1"
import downsampled_diffusion_b200 as dd
from tests import common as tc
from oracle import ddpm_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cfg = dict(tc.CS, precision=prec)
dev = torch.device("cuda:0")
net = tc.build_model(cfg, dd, "unet").to(dev).eval()
x = tc.randn(13, 2, 8, 8, 8).to(dev)
t = torch.tensor([500, 37], device=dev)
eng = net.engine(2, 8, 8)
print("ops:", len(eng.ops))
with torch.no_grad():
    eps = net(x, t)
torch.cuda.synchronize()
sd = {k: v.cpu() for k, v in net.state_dict().items()}
taps = {}
with torch.no_grad():
    ref = O.unet_forward(sd, cfg, x.cpu(), t.cpu(), taps=taps)
print("rel l2", tc.rel_l2(eps, ref), "nan:", bool(torch.isnan(eps).any()))
# time-bias rows vs oracle
import torch.nn.functional as F
temb = taps["temb"]
rb = net.downs[0][0]
col = eng.tb_off[id(rb)]
refb = F.linear(F.mish(temb), sd["downs.0.0.mlp.1.weight"], sd["downs.0.0.mlp.1.bias"])
print("tb err", tc.max_abs(eng.tb_batch[:, col:col + refb.shape[1]], refb))
