"""Batch sharding over the GPUs of one box.

Every frame is independent through the whole forward, so the path shards over the batch with no
data-path collective; the only exchange is the final gather of outputs (`pred_0`, `lowest_cost_bhw`,
`overall_mask_bhw`, ~1.6 MB per frame) over NCCL/NVLink.  One process per GPU (torchrun)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous split of `total` frames: rank r gets [lo, hi)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(data, rank, world):
    """Slice every tensor of a (cur_data | src_data) dictionary along the batch axis."""
    out = {}
    for k, v in data.items():
        if torch.is_tensor(v):
            lo, hi = shard_range(v.shape[0], rank, world)
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_outputs(outputs, group=None):
    """all_gather of the forward's output dictionary along the batch axis (equal shard sizes use one
    all_gather_into_tensor per tensor; ragged shards fall back to all_gather of padded tensors)."""
    world = dist.get_world_size(group)
    res = {}
    for k, v in outputs.items():
        if v is None:
            res[k] = None
            continue
        t = v.contiguous()
        was_bool = t.dtype == torch.bool
        if was_bool:
            t = t.to(torch.uint8)
        n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = [int(s.item()) for s in sizes]
        if len(set(sizes)) == 1:
            full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
            dist.all_gather_into_tensor(full, t, group=group)
        else:
            mx = max(sizes)
            pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
            pad[: t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
            full = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
        res[k] = full.bool() if was_bool else full
    return res


class GatherPlan:
    """Preallocated equal-shard gather for the timed path (no size exchange, no host sync)."""

    def __init__(self, outputs, world, group=None):
        self.group = group
        self.bufs = {}
        for k, v in outputs.items():
            if v is not None:
                dt = torch.uint8 if v.dtype == torch.bool else v.dtype
                self.bufs[k] = torch.empty((world * v.shape[0],) + tuple(v.shape[1:]), device=v.device, dtype=dt)

    def run(self, outputs):
        for k, buf in self.bufs.items():
            v = outputs[k]
            v = v.to(torch.uint8) if v.dtype == torch.bool else v
            dist.all_gather_into_tensor(buf, v.contiguous(), group=self.group)
        return self.bufs
