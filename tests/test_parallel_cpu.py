"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of shard -> (fake) forward -> gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from implicit_depth_b200.parallel import GatherPlan, PackedOutputs, gather_outputs, shard_batch, shard_range


def test_shard_range_covers_everything():
    for total in (1, 4, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    full = {"image_b3hw": torch.randn(total, 3, 4, 6), "rendered_depth": torch.randn(total, 2, 2, 3), "tag": "x"}
    mine = shard_batch(full, rank, world)
    # stand-in for the per-GPU forward: any per-frame function
    out = {"pred_0": mine["rendered_depth"] * 2 + mine["image_b3hw"].mean((1, 2, 3)).view(-1, 1, 1, 1),
           "overall_mask_bhw": mine["rendered_depth"][:, 0] > 0, "lowest_cost_bhw": None}
    g = gather_outputs(out)
    ref_pred = full["rendered_depth"] * 2 + full["image_b3hw"].mean((1, 2, 3)).view(-1, 1, 1, 1)
    ok = torch.equal(g["pred_0"], ref_pred) and torch.equal(g["overall_mask_bhw"], full["rendered_depth"][:, 0] > 0)
    ok = ok and g["lowest_cost_bhw"] is None and g["overall_mask_bhw"].dtype == torch.bool
    # the timed path: packed frame-major outputs, one collective (gather to the root / all_gather), two rotating slots
    if total % world == 0:
        for mode in ("root", "all"):
            plan = GatherPlan(out, world, mode=mode)
            for step in range(3):
                slot = step & 1
                scaled = {k: (v * (step + 1) if v is not None and v.dtype != torch.bool else v) for k, v in out.items()}
                for k, dst in plan.send_views(slot).items():
                    dst.copy_(scaled[k])
                _, res = plan.run(slot)
                if rank == 0 or mode == "all":
                    ok = ok and torch.equal(res["pred_0"], ref_pred * (step + 1))
                    ok = ok and torch.equal(res["overall_mask_bhw"], full["rendered_depth"][:, 0] > 0)
                    ok = ok and set(res) == {"pred_0", "overall_mask_bhw"}
                else:
                    ok = ok and res is None and plan.gathered_buffer(slot) is None
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])  # equal and ragged shards
def test_shard_forward_gather_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_packed_outputs_views_roundtrip():
    """`PackedOutputs`: frame-major packing -- the concatenation of two ranks' buffers is the packed global batch."""
    torch.manual_seed(0)
    outs = [{"pred_0": torch.randn(2, 8, 6, 10), "lowest_cost_bhw": torch.randn(2, 3, 5),
             "overall_mask_bhw": torch.rand(2, 3, 5) > 0.5, "none": None} for _ in range(2)]
    pk = PackedOutputs(outs[0])
    assert pk.frame_bytes % 256 == 0 and set(pk.fields) == {"pred_0", "lowest_cost_bhw", "overall_mask_bhw"}
    bufs = []
    for o in outs:
        b = torch.zeros(pk.nbytes(2), dtype=torch.uint8)
        for k, v in pk.views(b, 2).items():
            v.copy_(o[k])
        bufs.append(b)
    glob = pk.views(torch.cat(bufs), 4)
    for k in pk.fields:
        assert torch.equal(glob[k], torch.cat([o[k] for o in outs], 0)) and glob[k].dtype == outs[0][k].dtype
