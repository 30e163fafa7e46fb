"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the Python mirror has the
reference's module tree, the product path refuses to run without CUDA, and the multi-GPU host logic
(sharding + gather) works under gloo with world_size 2."""
import os
import re
import subprocess
import sys

import pytest
import torch

import downsampled_diffusion_b200 as dd
from downsampled_diffusion_b200 import _lib
from downsampled_diffusion_b200.parallel import shard_range
from tests import common as tc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ddb200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(dd_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"libddb200.so does not export {name}"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.dd_version() >= 100


def test_error_reporting_without_gpu():
    lib = _lib.lib()
    rc = lib.dd_q_sample(None, None, None, None, None, None, 1, 3, None)      # chw not a multiple of 4
    assert rc == -1 and b"multiple of 4" in lib.dd_last_error()


def test_module_tree_matches_reference_names():
    m = tc.build_model(tc.C2, dd, "dddpm_ae")
    sd = m.state_dict()
    assert len(sd) == 348
    for k, shape in {"latent_model.downs.0.0.block1.block.0.weight": (128, 8, 3, 3),
                     "latent_model.downs.0.2.fn.fn.to_qkv.weight": (384, 128, 1, 1),
                     "latent_model.downs.0.2.fn.norm.g": (1, 128, 1, 1),
                     "latent_model.ups.2.3.conv.weight": (128, 128, 4, 4),
                     "latent_model.final_conv.1.weight": (8, 128, 1, 1),
                     "downsample.conv.7.weight": (8, 64, 1, 1), "upsample.conv.1.c2.weight": (32, 32, 3, 3),
                     "posterior_log_variance_clipped": (1000,)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "vlb_weights" not in sd and hasattr(m, "vlb_weights")
    assert m.sample_shape == [8, 16, 16] and m.x_shape == [3, 64, 64] and int(m.dim_reduc) == 4
    assert sum(p.numel() for p in tc.build_model(tc.C3, dd, "dddpm_ae").parameters()) == 22671699
    assert callable(dd.DDPM.p_losses) and callable(dd.DownsampleDDPMAutoencoder.p_losses)


def test_no_cpu_fallback():
    m = tc.build_model(tc.CS, dd, "dddpm_ae")
    with pytest.raises(RuntimeError):
        m.latent_model(torch.zeros(1, 8, 8, 8), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError):
        m.rescaled_downsample(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError):
        m.q_sample(torch.zeros(1, 8, 8, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 8, 8, 8))
    with pytest.raises(ValueError):
        dd.DDPM(dict(tc.C1, loss_flat="max"), torch.nn.Identity(), "cpu", 1)
    with pytest.raises(ValueError):
        dd.make_beta_schedule("quadratic", 10)


def test_shard_range():
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DD_ROOT"])
from downsampled_diffusion_b200.parallel import sample_sharded, allreduce_gradients, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

class Stub:                      # stands in for DownsampleDDPM.sample: a deterministic function of the noise rows
    def sample(self, n, early_stop=None, noise=None):
        assert noise.shape[1] == n
        z = noise.sum(0)
        return z * 2.0, z
total = 5
g = torch.Generator().manual_seed(0)
noise = torch.randn(4, total, 2, 3, 3, generator=g)
x, z = sample_sharded(Stub(), total, noise=noise)
assert torch.equal(z, noise.sum(0)) and torch.equal(x, 2 * noise.sum(0)), "gathered result differs from single-process"
p = torch.nn.Parameter(torch.zeros(10)); p.grad = torch.full((10,), float(rank + 1))
q = torch.nn.Parameter(torch.zeros(3, 3)); q.grad = torch.full((3, 3), float(10 * (rank + 1)))
allreduce_gradients([p, q], bucket_bytes=16)
assert torch.allclose(p.grad, torch.full((10,), sum(range(1, world + 1)) / world))
assert torch.allclose(q.grad, torch.full((3, 3), 10 * sum(range(1, world + 1)) / world))
# gradients handed out as views of one flat buffer (what the training programs do) are reduced in place
flat = torch.arange(44, dtype=torch.float32) * (rank + 1)
a = torch.nn.Parameter(torch.zeros(2, 5)); a.grad = flat[0:10].view(2, 5)
b = torch.nn.Parameter(torch.zeros(19)); b.grad = flat[12:31]                # 2 elements of alignment gap before it, 1 after
c = torch.nn.Parameter(torch.zeros(12)); c.grad = flat[32:44]
allreduce_gradients([a, b, c, p])
mean = sum(range(1, world + 1)) / world
assert torch.allclose(flat, torch.arange(44, dtype=torch.float32) * mean), "flat gradient buffer was not reduced in place"
assert a.grad.data_ptr() == flat.data_ptr() and torch.allclose(b.grad, torch.arange(12, 31, dtype=torch.float32) * mean)
assert torch.allclose(p.grad, torch.full((10,), mean))                      # p went through the bucket path again (already equal on all ranks)
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_sharded_sampling_and_grad_allreduce_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DD_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
