"""Batch sharding over the GPUs of one box.

Every frame is independent through the whole forward, so the path shards over the batch with no
data-path collective; the only exchange is the final gather of outputs (`pred_0`, `lowest_cost_bhw`,
`overall_mask_bhw`, ~1.6 MB per frame) over NCCL/NVLink.  One process per GPU (torchrun).

`GatherPlan` is the timed path: the forward writes its outputs straight into ONE packed, frame-major send buffer
(`PackedOutputs`), ONE collective per step moves it (gather to rank 0 by default), on a communication stream of its
own so that step i's gather runs under step i+1's front phase, and only the root downloads the gathered batch.
`gather_outputs` is the general (ragged shards, any dictionary) convenience path.
"""
from __future__ import annotations

import torch

from . import _abi
import torch.distributed as dist

ALIGN = 256  # bytes: every field of a packed frame starts on a 256-byte boundary


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


class PackedOutputs:
    """Frame-major packing of an output dictionary: frame g of the (global) batch occupies bytes
    [g * frame_bytes, (g+1) * frame_bytes) as  key0[g] | key1[g] | ...  (each field 256-byte aligned).  Because the
    unit of interleave is the frame, the concatenation of the ranks' buffers IS the packed global batch, and every
    key is a strided view `[n_frames, ...]` of it -- one collective, one D2H copy, no re-layout anywhere."""

    def __init__(self, outputs):
        """outputs: dictionary of per-rank output tensors [B_local, ...] (None entries are skipped)."""
        self.fields, off = {}, 0
        for k, v in outputs.items():
            if v is None:
                continue
            dt = torch.uint8 if v.dtype == torch.bool else v.dtype
            n = 1
            for s in v.shape[1:]:
                n *= s
            nbytes = n * torch.empty((), dtype=dt).element_size()
            self.fields[k] = (off, tuple(v.shape[1:]), dt, v.dtype == torch.bool)
            off += (nbytes + ALIGN - 1) // ALIGN * ALIGN
        self.frame_bytes = off

    def nbytes(self, n_frames):
        return n_frames * self.frame_bytes

    def views(self, buf, n_frames):
        """buf: flat uint8 tensor (host or device) of at least nbytes(n_frames) -> {key: [n_frames, ...] view}."""
        out = {}
        for k, (off, shape, dt, is_bool) in self.fields.items():
            es = torch.empty((), dtype=dt).element_size()
            typed = buf[: self.nbytes(n_frames)].view(dt)
            strides, s = [], 1
            for d in reversed(shape):
                strides.append(s)
                s *= d
            v = typed.as_strided((n_frames,) + shape, (self.frame_bytes // es,) + tuple(reversed(strides)), off // es)
            out[k] = v.view(torch.bool) if is_bool else v
        return out


class GatherPlan:
    """Preallocated equal-shard gather for the timed path: no size exchange, no host sync, one collective.

        plan = GatherPlan(outputs_of_one_forward, world, mode="root")
        model.output_views = plan.send_views(slot)      # the forward writes its outputs into the send buffer
        out = model(...)
        done, gathered = plan.run(slot)                 # collective on the plan's own stream; `done` = its event

    mode "root": `dist.gather` to rank 0 -- `gathered` is the global batch on rank 0 and None elsewhere (only the
    root downloads);  mode "all": `all_gather_into_tensor`, every rank holds the global batch.  `slots` buffers
    rotate so that the collective / download of step i can run under the forward of step i+1."""

    def __init__(self, outputs, world, group=None, mode="root", slots=2, root=0):
        if mode not in ("root", "all"):
            raise ValueError("GatherPlan mode is 'root' or 'all'")
        self.group, self.world, self.mode, self.root = group, world, mode, root
        self.rank = dist.get_rank(group) if world > 1 else 0
        self.packed = PackedOutputs(outputs)
        first = next(v for v in outputs.values() if v is not None)
        self.B, dev = first.shape[0], first.device
        nb = self.packed.nbytes(self.B)
        self.send = [torch.zeros(nb, dtype=torch.uint8, device=dev) for _ in range(slots)]
        holds_all = mode == "all" or self.rank == root
        self.recv = [torch.zeros(nb * world, dtype=torch.uint8, device=dev) if (holds_all and world > 1) else None
                     for _ in range(slots)]
        self.cuda = dev.type == "cuda"  # (CPU tensors + gloo: host-logic tests; everything is synchronous there)
        self.stream = _abi.new_stream(dev) if self.cuda else None
        self.done = [None] * slots  # event: the collective that last used slot s has finished

    def send_views(self, slot):
        """Destination views for the forward's outputs (`B200BDModel.output_views`).  Waits (on the current stream)
        for the collective that last read this slot."""
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.done[slot])
        return self.packed.views(self.send[slot], self.B)

    def pack(self, outputs, slot):
        """For callers that did not route the forward into `send_views`: copy an output dictionary in."""
        dst = self.send_views(slot)
        for k, v in dst.items():
            v.copy_(outputs[k])

    def gathered_buffer(self, slot):
        """The flat buffer a consumer downloads: the global batch (root / mode 'all'), else None."""
        if self.world == 1:
            return self.send[slot]
        return self.recv[slot]

    def run(self, slot=0):
        """Issue the collective for `slot` on the plan's stream after everything enqueued so far on the current
        stream.  Returns (event, {key: [world*B, ...] strided view} | None)."""
        n = self.world * self.B
        if self.world == 1:
            ev = None
            if self.cuda:
                ev = torch.cuda.Event()
                ev.record()
            self.done[slot] = ev
            return ev, self.packed.views(self.send[slot], n)
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self._collective(slot)
                ev = torch.cuda.Event()
                ev.record(self.stream)
        else:
            self._collective(slot)
            ev = None
        self.done[slot] = ev
        buf = self.recv[slot]
        return ev, (self.packed.views(buf, n) if buf is not None else None)

    def _collective(self, slot):
        if self.mode == "all":
            if self.cuda:
                dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
            else:  # gloo has no all_gather_into_tensor
                dist.all_gather(list(self.recv[slot].view(self.world, -1).unbind(0)), self.send[slot], group=self.group)
        else:
            parts = list(self.recv[slot].view(self.world, -1).unbind(0)) if self.rank == self.root else None
            dist.gather(self.send[slot], parts, dst=self.root, group=self.group)
