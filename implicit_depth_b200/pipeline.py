"""Host <-> device streaming around `B200BDModel.forward`.

`FramePipeline` keeps two device-side input slots and three pinned host-side output slots and runs three CUDA
streams: while batch i is in the forward (compute stream), batch i+1 is uploaded (copy-in stream) and the outputs
of batch i-1 are downloaded (copy-out stream).  Ordering is by CUDA events only; the host blocks once per batch,
on the event of the batch whose results it hands back.  This is the path `bench.py` times as `e2e`.  Batches staged
with `staging.FrameStaging` go up as one copy and are read by the forward in place.  With a
`parallel.GatherPlan` the forward writes its outputs straight into the plan's packed send buffer, ONE collective
per step gathers the ranks' buffers on the plan's own stream (under the next forward), and the rank that holds the
gathered batch downloads it in ONE copy; the other ranks download nothing.
"""
from __future__ import annotations

import torch

from . import _abi

from .staging import FrameStaging, StagedDict, StagedFrame


class FramePipeline:
    def __init__(self, model, device, gather=None, encoder_ahead=False, encoder_priority=0, **forward_kwargs):
        """model: B200BDModel (or anything with the reference's forward signature); gather: optional
        `parallel.GatherPlan` (packed outputs, one collective per step, root-only download), or a plain callable
        applied to the output dict on the compute stream (e.g. `parallel.gather_outputs`).

        encoder_ahead (staged batches, B200BDModel with its built-in encoder): the image-prior encoder depends only
        on the current images, so the encoder of batch i+1 is launched on a fourth stream as soon as batch i has taken
        its own encoder outputs, and runs under the forward of batch i; the forward itself then is matching encoder +
        plane sweep on all SMs followed by the back phase.  Results are unchanged (same kernels, same order per
        batch)."""
        self.model, self.device, self.gather, self.kw = model, torch.device(device), gather, forward_kwargs
        self.s_in = _abi.new_stream(self.device)
        self.s_out = _abi.new_stream(self.device)
        self.encoder_ahead = bool(encoder_ahead)
        if self.encoder_ahead:
            model.encoder_ahead = True
            self.s_enc = _abi.new_stream(self.device, priority=int(encoder_priority))
        self.ev_enc = [None, None]         # encoder of the batch in slot s has finished (encoder_ahead)
        self.slots = [None, None]          # device input dictionaries
        self.HOST_SLOTS = 3                # a yielded dictionary stays valid while the next batch is being produced
        self.host_out = [None] * self.HOST_SLOTS   # pinned output dictionaries
        self.host_buf = [None] * self.HOST_SLOTS   # pinned packed buffers (GatherPlan path)
        self.ev_free = [None, None]        # forward that last read input slot s has finished
        self.ev_d2h = [None, None]         # download that last read the gather plan's device slot s has finished
        from .parallel import GatherPlan

        self.plan = gather if isinstance(gather, GatherPlan) else None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload_staged(self, slot, frame):
        """A batch staged by `staging.FrameStaging`: ONE H2D copy of the whole pinned buffer into the slot's device
        frame, which the forward then reads in place."""
        if not isinstance(self.slots[slot], StagedFrame) or self.slots[slot].staging.fields != frame.staging.fields:
            self.slots[slot] = frame.staging.device_frame(self.device)
            self.s_in.wait_stream(torch.cuda.current_stream(self.device))  # allocation-time work of the caller's stream
        with torch.cuda.stream(self.s_in):
            if self.ev_free[slot] is not None:
                self.s_in.wait_event(self.ev_free[slot])
            self.h2d_bytes = FrameStaging.upload(frame, self.slots[slot])
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def _upload(self, slot, cur, src=None):
        """Enqueue the H2D copies of one batch on the copy-in stream; returns the event that marks their end."""
        if isinstance(cur, StagedFrame):
            return self._upload_staged(slot, cur)
        if not isinstance(self.slots[slot], tuple):
            mk = lambda d: {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in d.items()}
            self.slots[slot] = (mk(cur), mk(src))
            self.s_in.wait_stream(torch.cuda.current_stream(self.device))
        dcur, dsrc = self.slots[slot]
        with torch.cuda.stream(self.s_in):
            if self.ev_free[slot] is not None:
                self.s_in.wait_event(self.ev_free[slot])
            n = 0
            for dst, srcd in ((dcur, cur), (dsrc, src)):
                for k, v in srcd.items():
                    dst[k].copy_(v, non_blocking=True)
                    n += v.numel() * v.element_size()
            self.h2d_bytes = n
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def _encoder_hook(self, main, next_slot, ev_next):
        """Called by the forward of batch i right after it copied its encoder outputs out: launch the encoder of
        batch i+1 (already being uploaded into `next_slot`) on the encoder stream."""
        def hook():
            nxt = self.slots[next_slot]
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.s_enc):
                self.s_enc.wait_event(ev)       # the encoder plan's buffers are free again
                self.s_enc.wait_event(ev_next)  # the next batch's images have arrived
                self.model.run_encoder(nxt.cur["image_b3hw"], nxt.staging.K, nxt.staging.P,
                                       bool(self.kw.get("infer_depth", False)))
                done = torch.cuda.Event()
                done.record(self.s_enc)
            self.ev_enc[next_slot] = done
        return hook

    def _forward(self, slot, ev_in, next_slot=None, ev_next=None):
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ev_in)
        staged = isinstance(self.slots[slot], StagedFrame)
        if staged:
            frame = self.slots[slot]
            dcur, dsrc = StagedDict(frame.cur), frame.src  # (the forward may add keys to cur_data)
            dcur.frame = frame
        else:
            dcur, dsrc = self.slots[slot]
            dcur = dict(dcur)
        ahead = self.encoder_ahead and staged
        if ahead:
            if self.ev_enc[slot] is not None:   # this batch's encoder ran ahead, under the previous forward
                main.wait_event(self.ev_enc[slot])
                self.ev_enc[slot] = None
            self.model.after_encoder_handoff = \
                self._encoder_hook(main, next_slot, ev_next) if ev_next is not None else None
        if self.plan is not None:
            # the forward writes into the plan's send buffer of this slot: first make sure the download that last read
            # the slot's gathered buffer is over (the collective below overwrites it), then route the outputs
            if self.ev_d2h[slot] is not None:
                if self.plan.world > 1:   # the download read recv[slot], which the collective writes
                    self.plan.stream.wait_event(self.ev_d2h[slot])
                else:                     # single rank: the download read send[slot], which the forward writes
                    main.wait_event(self.ev_d2h[slot])
            self.model.output_views = self.plan.send_views(slot)
        try:
            out = self.model("test", dcur, dsrc, **self.kw)
        finally:
            if ahead:
                self.model.after_encoder_handoff = None
            if self.plan is not None:
                self.model.output_views = None
        self.ev_free[slot] = torch.cuda.Event()
        self.ev_free[slot].record(main)
        if self.plan is not None:
            ev_ready, _ = self.plan.run(slot)  # one collective on the plan's stream, under the next forward
            return out, ev_ready
        if self.gather is not None:
            g = self.gather(out)
            out = {k: (g[k] if k in g else v) for k, v in out.items()}
            ev = torch.cuda.Event()
            ev.record(main)
            return out, ev
        return out, self.ev_free[slot]

    def _download(self, slot, hslot, out, ev_ready):
        """Enqueue the D2H copies of one batch on the copy-out stream; returns the event that marks their end.
        GatherPlan path: ONE copy of the packed (gathered) buffer, on the rank(s) that hold it."""
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_ready)
            n = 0
            if self.plan is not None:
                buf = self.plan.gathered_buffer(slot)
                if buf is not None:
                    if self.host_buf[hslot] is None:
                        self.host_buf[hslot] = torch.empty(buf.shape, dtype=torch.uint8).pin_memory()
                        self.host_out[hslot] = self.plan.packed.views(self.host_buf[hslot],
                                                                      self.plan.world * self.plan.B)
                    self.host_buf[hslot].copy_(buf, non_blocking=True)
                    n = buf.numel()
                else:
                    self.host_out[hslot] = {}  # the results of this step live on the root rank
            else:
                if self.host_out[hslot] is None:
                    self.host_out[hslot] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory()
                                            for k, v in out.items() if v is not None}
                for k, h in self.host_out[hslot].items():
                    h.copy_(out[k], non_blocking=True)
                    out[k].record_stream(self.s_out)
                    n += h.numel() * h.element_size()
            self.d2h_bytes = n
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self.ev_d2h[slot] = ev
        return ev

    def run(self, host_batches):
        """host_batches: iterable of (cur_data, src_data) dictionaries of PINNED host tensors, or of host
        `staging.StagedFrame`s (one pinned buffer per batch, one copy).  Yields one dictionary of pinned host tensors
        per batch; it stays valid until the generator has been advanced twice more (three host slots rotate).  With a
        `GatherPlan` in mode "root" the root rank's dictionaries hold the gathered global batch (strided views of one
        packed pinned buffer) and the other ranks' dictionaries are empty."""
        as_args = lambda b: (b,) if isinstance(b, StagedFrame) else b
        pending = None  # (slot, event of the D2H copy)
        it = iter(host_batches)
        nxt = next(it, None)
        i = 0
        ev_in = self._upload(0, *as_args(nxt)) if nxt is not None else None
        while nxt is not None:
            slot = i & 1
            following = next(it, None)
            ev_next = self._upload(slot ^ 1, *as_args(following)) if following is not None else None
            out, ev_ready = self._forward(slot, ev_in, slot ^ 1, ev_next)
            hslot = i % self.HOST_SLOTS
            ev_done = self._download(slot, hslot, out, ev_ready)
            if pending is not None:
                pending[1].synchronize()
                yield self.host_out[pending[0]]
            pending = (hslot, ev_done)
            nxt, ev_in = following, ev_next
            i += 1
        if pending is not None:
            pending[1].synchronize()
            yield self.host_out[pending[0]]
