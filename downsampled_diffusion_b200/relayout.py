"""Tabulated re-layouts for the training programs (one `dd_gather_f32` launch instead of hundreds of strided copies).

A training program derives ~600 packed buffers from the fp32 master parameters (K-major weight matrices, flipped /
transposed input-gradient forms, space-to-depth forms of the strided convolutions, padded gains and biases) and has to
refresh all of them after every optimizer step; in the other direction it maps the packed weight gradients back to
parameter layout.  Every one of these is a fixed placement of elements, written in the engine as small torch expressions
(`w.permute(...).reshape(...)`, slice assignments).  Instead of launching those expressions every step, the engine pushes
*index codes* through them once -- the parameters (or the gradient arena) temporarily hold "which tensor" / "which
element" numbers, exact in fp32 -- reads back where every destination element came from, checks the table against the
real expressions on the real data, and from then on replays the table.  Expressions that are not pure placements (a sum, a
scale) fail the check and keep running as they are.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib as L

GATHER_CHUNK = 4096          # elements per block: csrc/optim.cu
EXACT = 1 << 24              # integers below this are exact in fp32


class GatherPlan:
    """dst_k[i] = src[ord - 1][idx] (or 0) for destination buffers dst_k, codes = ord << 32 | idx."""

    def __init__(self, dsts: Sequence[torch.Tensor], codes: Sequence[torch.Tensor], first_dst_per_call: bool = False):
        dev = codes[0].device
        segs, blocks, start = [], [], 0
        for si, (d, c) in enumerate(zip(dsts, codes)):
            n = c.numel()
            if d is not None:
                assert d.numel() == n and d.is_contiguous() and d.dtype == torch.float32
            segs += [d.data_ptr() if d is not None else 0, start, n]
            blocks += [[si, k] for k in range((n + GATHER_CHUNK - 1) // GATHER_CHUNK)]
            start += n
        self.codes = torch.cat([c.reshape(-1) for c in codes]).contiguous()
        self.segs = torch.tensor(segs, dtype=torch.int64, device=dev)          # uint64 bit patterns
        self.blocks = torch.tensor(blocks, dtype=torch.int32, device=dev)
        self.n_blocks = len(blocks)
        self.keep = list(dsts)
        self.first_dst_per_call = first_dst_per_call
        self.src_key: Optional[tuple] = None
        self.src_table: Optional[torch.Tensor] = None
        self.total = start

    def run(self, src_ptrs: tuple, dst0: Optional[torch.Tensor] = None) -> None:
        if src_ptrs != self.src_key:            # sources moved (model.to(), p.data = ...): pinned, non-blocking upload
            self.src_table = torch.tensor(src_ptrs, dtype=torch.int64).pin_memory().to(self.codes.device, non_blocking=True)
            self.src_key = src_ptrs
        assert (dst0 is not None) == self.first_dst_per_call
        L.call("dd_gather_f32", L.ptr(self.segs), L.ptr(self.blocks), self.n_blocks, L.ptr(self.codes), L.ptr(self.src_table),
               L.ptr(dst0) if dst0 is not None else None, L.stream())


def codes_from_probes(ord_probe: torch.Tensor, idx_probe: torch.Tensor, numels: torch.Tensor) -> Optional[torch.Tensor]:
    """Two fp32 snapshots of one destination -- produced with the sources holding (tensor ordinal + 1) and (element index)
    -- to int64 codes; None when the values are not a clean placement (non-integers, out of range)."""
    o, x = ord_probe.reshape(-1), idx_probe.reshape(-1)
    ordi, idx = o.round().long(), x.round().long()
    lim = numels[(ordi - 1).clamp(min=0, max=numels.numel() - 1)]
    ok = (o == ordi) & (x == idx) & (ordi >= 0) & (ordi <= numels.numel()) & (idx >= 0) & \
         (((ordi > 0) & (idx < lim)) | ((ordi == 0) & (idx == 0)))
    if not bool(ok.all()):                      # one host synchronisation per destination
        return None
    return (ordi << 32) | idx


def apply_codes(codes: torch.Tensor, flat_src: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
    """Reference evaluation of a code table with torch indexing (used once, to validate a table on real data)."""
    ordi, idx = codes >> 32, codes & 0xFFFFFFFF
    pos = (offsets[(ordi - 1).clamp(min=0)] + idx).clamp(max=flat_src.numel() - 1)
    return torch.where(ordi > 0, flat_src[pos], torch.zeros((), dtype=flat_src.dtype, device=flat_src.device))
