"""N-GPU parity of the sharded paths against rank 0's single-GPU result (torchrun, NCCL):
 * sampling: sample_sharded over the global pre-drawn noise == model.sample on the whole batch (within the bf16 chain tolerance: tile shapes depend on the batch);
 * training: per-rank micro-batches + allreduce_gradients (mean) == the gradient of the full batch on one GPU."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import downsampled_diffusion_b200 as dd
from downsampled_diffusion_b200 import parallel
from tests import common as tc
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(tc.CS, T=20)
m = tc.build_model(cfg, dd, "dddpm_ae", device=str(dev)).to(dev).eval()
B = 4 * world
noise = torch.stack([tc.randn(100 + i, B, 8, 8, 8) for i in range(cfg["T"] + 1)]).to(dev)
with torch.no_grad():
    x, z = parallel.sample_sharded(m, B, noise=noise)
    if rank == 0:
        xr, zr = m.sample(B, noise=noise)
        print("sampling: sharded vs single-GPU max-abs", float((x - xr).abs().max()), float((z - zr).abs().max()))
        assert float((z - zr).abs().max()) < 5e-2 and float((x - xr).abs().max()) < 2e-2
# training: gradient of the global batch mean loss
m.train()
xs = tc.rand_pm1(7, B, 3, 32, 32).to(dev)
t = torch.arange(B, device=dev) * 3 % cfg["T"]
eps = tc.randn(8, B, 8, 8, 8).to(dev)
lo, hi = parallel.shard_range(B, rank, world)
obj, _ = m.losses(xs[lo:hi], t[lo:hi], eps=eps[lo:hi])
obj.backward()
params = list(m.parameters())
parallel.allreduce_gradients(params)
g_dp = [p.grad.clone() if p.grad is not None else None for p in params]
if rank == 0:
    m.zero_grad()
    obj_full, _ = m.losses(xs, t, eps=eps)
    obj_full.backward()
    worst = 0.0
    for gd, p in zip(g_dp, params):
        if p.grad is None: continue
        den = float(p.grad.norm()) + 1e-12
        worst = max(worst, float((gd - p.grad).norm()) / den)
    print(f"training: data-parallel (x{world}) vs full-batch gradient, worst relative L2 over {len(params)} tensors: {worst:.3e}")
    assert worst < 5e-3          # TF32 convolutions: per-shard tiles round differently
dist.barrier(); dist.destroy_process_group()
