"""Multi-GPU drivers: one process per GPU (torchrun), torch.distributed for the plumbing.

Sampling: every sample's chain is independent, so the batch is sharded over ranks with
NO communication during the T steps and one final gather.  To stay identical to the single-GPU result,
rank r consumes rows [lo, hi) of the *global* pre-drawn noise tensor.
Training: data-parallel replicas, one gradient all-reduce (mean) per optimizer step.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `total` items for `rank` (first total % world ranks get one more)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_noise(noise: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """noise: (n_steps+1, B_global, C, H, W) -> this rank's (n_steps+1, B_local, C, H, W) view."""
    lo, hi = shard_range(noise.shape[1], rank, world)
    return noise[:, lo:hi]


def gather_batch(local: torch.Tensor, total: int, group=None) -> Optional[torch.Tensor]:
    """Concatenate per-rank batches along dim 0 on every rank (all_gather; ragged shards are padded)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)


def sample_sharded(model, total_batch: int, noise: Optional[torch.Tensor] = None, early_stop: int = None,
                   gather: bool = True):
    """`model.sample(total_batch)` with the batch sharded over the process group.

    noise: optional GLOBAL pre-drawn chain noise (n_steps+1, total_batch, C, H, W), host or device.
    Returns what model.sample returns (a tensor, or an (x, z) tuple for dDDPM) for the global batch when
    `gather`, else for the local shard."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(total_batch, rank, world)
    local_noise = None if noise is None else noise[:, lo:hi]
    out = model.sample(hi - lo, early_stop=early_stop, noise=local_noise)
    if not gather or world == 1:
        return out
    if isinstance(out, tuple):
        return tuple(gather_batch(o, total_batch) for o in out)
    return gather_batch(out, total_batch)


def _allreduce_mean(flat: torch.Tensor, world: int) -> None:
    """In-place mean over the group: NCCL averages inside the collective (ReduceOp.AVG, no second pass over the
    buffer); gloo (the CPU tests) has no AVG, so it sums and divides."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.div_(world)


def allreduce_gradients(params: Sequence[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> None:
    """DDP-style mean all-reduce of .grad over the default group, in flat fp32 buckets (NCCL over NVLink 5;
    NVSwitch makes the cost bandwidth-bound, so buckets are sized for launch latency, not link count)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    # The training programs hand out each network's parameter gradients as views of ONE flat buffer (autograd.py,
    # param_grads): such a group is reduced in place, as it lies, with no gather / scatter copies around the collective.
    by_storage = {}
    for p in params:
        if p.grad is not None:
            by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p.grad)
    loose = []
    for gs in by_storage.values():
        lo = min(g.storage_offset() for g in gs)
        hi = max(g.storage_offset() + g.numel() for g in gs)
        dense = all(g.is_contiguous() and g.dtype == torch.float32 for g in gs) and 10 * sum(g.numel() for g in gs) >= 9 * (hi - lo)
        if len(gs) > 1 and dense:
            flat = torch.empty(0, dtype=torch.float32, device=gs[0].device).set_(gs[0].untyped_storage(), lo, (hi - lo,))
            _allreduce_mean(flat, world)
        else:
            loose += gs
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        _allreduce_mean(flat, world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in loose:
        bucket.append(g)
        size += g.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()


class GradReducer:
    """Data-parallel gradient averaging overlapped with the backward pass (the reference trains on one GPU,
    trainers/trainer_ddpm.py:219-251).

    The training programs write each network's parameter gradients into ONE flat buffer at the end of that network's backward
    (autograd.py, param_grads).  While armed, that moment launches the buffer's NCCL all-reduce (ReduceOp.AVG, in place) on a
    side stream, so the up-/down-sampling nets' and the U-Net's collectives run under whatever is still back-propagating;
    `finish()` makes the main stream wait for them and reduces what did not come through a flat buffer.

        reducer = GradReducer(model)
        reducer.arm(); loss.backward(); reducer.finish(); optimizer.step()

    Arm only for a backward that starts from cleared gradients (`zero_grad(set_to_none=True)`, the default): autograd then
    adopts the flat buffer's views as `.grad`, so the in-place result is what the optimizer reads.  When gradients are
    accumulated over micro-batches (trainer_ddpm.py:35), leave it unarmed and call `allreduce_gradients` after the last backward."""

    def __init__(self, model: torch.nn.Module):
        self.params = list(model.parameters())
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        dev = self.params[0].device
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.reduced = []

    def _on_flat(self, flat: torch.Tensor) -> None:
        main = torch.cuda.current_stream(flat.device)
        self.stream.wait_stream(main)                    # the gather launch that filled `flat` is on the main stream
        with torch.cuda.stream(self.stream):
            _allreduce_mean(flat, self.world)
        flat.record_stream(self.stream)
        self.reduced.append(flat)

    def arm(self) -> None:
        if self.world == 1 or self.stream is None:
            return
        from . import autograd as _ag
        self.reduced = []
        _ag.set_grads_ready_hook(self._on_flat)

    def finish(self) -> None:
        if self.world == 1 or self.stream is None:
            return
        from . import autograd as _ag
        _ag.set_grads_ready_hook(None)
        torch.cuda.current_stream().wait_stream(self.stream)
        done = {f.untyped_storage().data_ptr() for f in self.reduced}
        rest = [p for p in self.params if p.grad is not None and p.grad.untyped_storage().data_ptr() not in done]
        if rest:
            allreduce_gradients(rest)
        self.reduced = []
