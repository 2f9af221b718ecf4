"""torchrun worker of tests/test_parallel_gpu.py (one rank per GPU, NCCL): every rank runs the forward on its own shard
of a global batch; the gathered outputs must be bit-identical to the single-GPU forward of the whole batch.  Covers both
multi-GPU paths: `GatherPlan` (packed outputs, one collective on its own stream, through `FramePipeline`) and
`gather_outputs` (general all_gather).  Prints one line `GATHER_OK {...}` from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from implicit_depth_b200 import synthetic  # noqa: E402
from implicit_depth_b200.bd_model import B200BDModel, default_options  # noqa: E402
from implicit_depth_b200.parallel import GatherPlan, gather_outputs, shard_batch  # noqa: E402
from implicit_depth_b200.pipeline import FramePipeline  # noqa: E402
from implicit_depth_b200.staging import FrameStaging  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, K, D, BL = 192, 256, 7, 16, 2  # per-rank batch 2
    opts = default_options(image_width=W, image_height=H, matching_num_depth_bins=D)
    model = B200BDModel(opts)
    synthetic.init_model_weights(model, seed=0)
    model = model.to(dev).eval()
    model.use_cuda_graph = True
    keys = ("pred_0", "lowest_cost_bhw", "overall_mask_bhw")
    n_steps = 3
    batches = [synthetic.make_frame_batch(9000 + s, world * BL, K, H, W) for s in range(n_steps)]
    to_dev = lambda d: {k: (v if torch.is_tensor(v) else torch.from_numpy(v)).to(dev) for k, v in d.items()}

    # --- reference: rank 0 runs every global batch alone (the kernels are batch-invariant) -------------------
    full = []
    if rank == 0:
        for cur, src in batches:
            o = model("test", to_dev(cur), to_dev(src), return_mask=True)
            full.append({k: o[k].cpu() for k in keys})

    # --- path 1: general all_gather of the output dictionary -----------------------------------------------
    ok_general = True
    for s, (cur, src) in enumerate(batches):
        o = model("test", to_dev(shard_batch({k: torch.from_numpy(v) for k, v in cur.items()}, rank, world)),
                  to_dev(shard_batch({k: torch.from_numpy(v) for k, v in src.items()}, rank, world)), return_mask=True)
        g = gather_outputs({k: o[k] for k in keys})
        if rank == 0:
            ok_general &= all(torch.equal(g[k].cpu(), full[s][k]) for k in keys)

    # --- path 2: the timed path -- staged inputs, packed outputs, one collective per step on its own stream --
    staging = FrameStaging(BL, K, H, W, P=8, matching_scale=opts.matching_scale)
    hosts = []
    for cur, src in batches:
        c = {k: v[rank * BL:(rank + 1) * BL] for k, v in cur.items()}
        s_ = {k: v[rank * BL:(rank + 1) * BL] for k, v in src.items()}
        hosts.append(staging.host_frame().fill(c, s_))
    ok_plan = {}
    for mode in ("root", "all"):
        probe = model("test", to_dev({k: v[:BL] for k, v in batches[0][0].items()}),
                      to_dev({k: v[:BL] for k, v in batches[0][1].items()}), return_mask=True)
        plan = GatherPlan({k: probe[k] for k in keys}, world, mode=mode)
        pipe = FramePipeline(model, dev, gather=plan, return_mask=True)
        good = True
        for s, res in enumerate(pipe.run(iter(hosts))):
            if rank == 0:
                eq = {k: bool(torch.equal(res[k], full[s][k])) for k in keys}
                good &= all(eq.values())
            elif mode == "root":
                eq = {"empty": len(res) == 0}
                good &= (len(res) == 0)
            else:
                eq = {k: tuple(res[k].shape) == (world * BL,) + tuple(probe[k].shape[1:]) for k in keys}
                good &= all(eq.values())
            if not all(eq.values()):
                print(f"[rank {rank}] mode {mode} step {s}: {eq}", file=sys.stderr, flush=True)
        flag = torch.tensor([1 if good else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_plan[mode] = bool(flag.item())
        d2h = pipe.d2h_bytes
        ok_plan[mode + "_d2h_bytes_rank0"] = d2h if rank == 0 else None
    torch.cuda.synchronize()
    if rank == 0:
        print("GATHER_OK " + json.dumps({"world": world, "general": bool(ok_general), **ok_plan}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
