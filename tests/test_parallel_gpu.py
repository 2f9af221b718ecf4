"""Multi-GPU path on hardware (BASELINE config 3): ranks shard the batch, NCCL gathers the outputs; the
gathered tensors must equal the single-GPU outputs of the same frames bit for bit.  Needs >= 2 GPUs (skipped on a
one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_parallel_gpu.py -m gpu`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_gather_plan_pipeline_single_rank_is_bit_identical():
    """The packed-output path of `FramePipeline` at world size 1 (what bench.py's e2e runs at N=1): the forward writes
    into the plan's send buffer, one D2H copy per batch, results equal to direct calls."""
    from implicit_depth_b200 import synthetic
    from implicit_depth_b200.bd_model import B200BDModel, default_options
    from implicit_depth_b200.parallel import GatherPlan
    from implicit_depth_b200.pipeline import FramePipeline
    from implicit_depth_b200.staging import FrameStaging

    m = B200BDModel(default_options(image_width=256, image_height=192, matching_num_depth_bins=16))
    synthetic.init_model_weights(m, seed=0)
    m = m.cuda().eval()
    m.use_cuda_graph = True
    st = FrameStaging(1, 7, 192, 256, P=8)
    hosts, direct = [], []
    for i in range(5):
        cur, src = synthetic.make_frame_batch(6200 + i, 1, 7, 192, 256)
        hosts.append(st.host_frame().fill(cur, src))
        o = m("test", {k: torch.from_numpy(v).cuda() for k, v in cur.items()},
              {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
        direct.append({k: v.cpu() for k, v in o.items()})
    plan = GatherPlan({k: v.cuda() for k, v in direct[0].items()}, 1)
    pipe = FramePipeline(m, "cuda", gather=plan, return_mask=True)
    n = 0
    for res, d in zip(pipe.run(iter(hosts)), direct):
        for k in d:
            assert torch.equal(res[k], d[k]), k
        n += 1
    assert n == 5 and pipe.d2h_bytes == plan.packed.nbytes(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_gather_equals_single_gpu_outputs():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(HERE, "gather_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("GATHER_OK ")]
    assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-4000:]
    sys.stderr.write("\n".join(l for l in r.stderr.splitlines() if l.startswith("[rank")) + "\n")
    res = json.loads(line[-1][len("GATHER_OK "):])
    from cases import record

    record("two_gpu_gather", kind="nccl_gather_equality", **{k: v for k, v in res.items() if v is not None})
    assert res["world"] == 2 and res["general"] and res["root"] and res["all"], res
