"""Input side of the step on the GPU: `b200_relative_poses` and `b200_intrinsics_pyramid` against
the reference-generated vectors and the oracle, and the staged forward (one buffer per batch, read in place by the
CUDA graph) against the ordinary dictionary call -- bit-identical, since both run the same kernels."""
import numpy as np
import pytest
import torch

from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from implicit_depth_b200.staging import FrameStaging, intrinsics_pyramid, relative_poses
from oracle import planesweep as O

from cases import GOLDEN

pytestmark = pytest.mark.gpu


def test_relative_poses_kernel_vs_reference_golden_and_oracle():
    g = np.load(f"{GOLDEN}/input_side.npz")
    cur, src = synthetic.make_frame_batch(4100, 2, 5, 192, 256)
    c = lambda x: torch.from_numpy(x).cuda()
    s2c, c2s = relative_poses(c(src["cam_T_world_b44"]), c(src["world_T_cam_b44"]), c(cur["cam_T_world_b44"]),
                              c(cur["world_T_cam_b44"]))
    assert np.abs(s2c.cpu().numpy() - g["src_cam_T_cur_cam"]).max() <= 5e-7   # entries are O(1): ~2 ulp
    assert np.abs(c2s.cpu().numpy() - g["cur_cam_T_src_cam"]).max() <= 5e-7
    # larger batch against the fp64 oracle; non-contiguous / fp64 inputs are accepted like .float()
    cur, src = synthetic.make_frame_batch(4101, 16, 7, 24, 32)
    r64 = O.relative_poses(*(x.astype(np.float64) for x in (src["cam_T_world_b44"], src["world_T_cam_b44"],
                                                            cur["cam_T_world_b44"], cur["world_T_cam_b44"])))
    s2c, c2s = relative_poses(c(src["cam_T_world_b44"]).double(), c(src["world_T_cam_b44"]),
                              c(cur["cam_T_world_b44"]), c(cur["world_T_cam_b44"]))
    assert np.abs(s2c.cpu().numpy() - r64[0]).max() <= 1e-6
    assert np.abs(c2s.cpu().numpy() - r64[1]).max() <= 1e-6


def test_intrinsics_pyramid_kernel_vs_reference_golden():
    g = np.load(f"{GOLDEN}/input_side.npz")
    K0 = torch.from_numpy(g["K_s"][0]).cuda()  # level 0 of the three reference cameras
    Ks, invKs = intrinsics_pyramid(K0, levels=5)
    assert len(Ks) == 5 and tuple(Ks[0].shape) == (3, 4, 4)
    for i in range(5):
        assert np.array_equal(Ks[i].cpu().numpy(), g["K_s"][i])  # bit-exact (power-of-two scaling)
        ref = g["invK_s"][i]
        # the reference inverts with fp32 LAPACK, the kernel with an fp64 adjugate: a few ulps of each entry's row
        assert np.abs(invKs[i].cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    # batched [B,K,4,4] input keeps its shape
    Kb = K0[:2, None].expand(2, 7, 4, 4)
    Ks2, inv2 = intrinsics_pyramid(Kb, levels=2)
    assert tuple(Ks2[1].shape) == (2, 7, 4, 4) and torch.equal(Ks2[1][1, 3], Ks[1][1])
    assert torch.equal(inv2[1][0, 6], invKs[1][0])


def _model(**kw):
    m = B200BDModel(default_options(image_width=256, image_height=192, matching_num_depth_bins=16, **kw))
    synthetic.init_model_weights(m, seed=0)
    return m.cuda().eval()


@pytest.mark.parametrize("graph", [False, True])
def test_staged_forward_is_bit_identical_to_dictionary_forward(graph):
    m = _model()
    m.use_cuda_graph = graph
    st = FrameStaging(1, 7, 192, 256, P=8)
    slots = [st.device_frame("cuda"), st.device_frame("cuda")]
    for i in range(4):
        cur, src = synthetic.make_frame_batch(6100 + i, 1, 7, 192, 256)
        want = m("test", {k: torch.from_numpy(v).cuda() for k, v in cur.items()},
                 {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
        host = st.host_frame().fill(cur, src)
        dev = slots[i & 1]
        FrameStaging.upload(host, dev)
        assert m._staged_images(dev.cur, dev.src, dev.cur["image_b3hw"]) is not None
        got = m("test", dev.cur, dev.src, return_mask=True)
        for k in want:
            assert torch.equal(got[k], want[k]), (i, k)
    if graph:
        # one graph per staging slot, sharing the launch plans; the dictionary path keeps its own
        staged = [v for v in m._graphs.values() if v[4]]
        assert len(staged) == 2 and len(m._state) == 1


def test_staged_temporal_forward_and_pipeline():
    """Temporal model (prior warp outside the graph) through staged frames, and `FramePipeline` fed with staged
    host frames: one H2D copy per batch, same results as direct calls."""
    from implicit_depth_b200.pipeline import FramePipeline

    m = _model(use_prior=True)
    m.use_cuda_graph = True
    st = FrameStaging(1, 7, 192, 256, P=1, temporal=True)
    hosts, direct = [], []
    for i in range(4):
        cur, src = synthetic.make_frame_batch(6200 + i, 1, 7, 192, 256, num_rendered=1, temporal=True)
        o = m("test", {k: torch.from_numpy(v).cuda() for k, v in cur.items()},
              {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
        direct.append({k: v.cpu().clone() for k, v in o.items()})
        hosts.append(st.host_frame().fill(cur, src))
    pipe = FramePipeline(m, "cuda", return_mask=True)
    got = [{k: v.clone() for k, v in res.items()} for res in pipe.run(iter(hosts))]
    assert len(got) == 4
    for g_, d_ in zip(got, direct):
        for k in d_:
            assert torch.equal(g_[k], d_[k]), k
    assert pipe.h2d_bytes == st.nbytes


def test_staged_entry_replaced_by_the_caller_is_honoured():
    """A caller may overwrite one entry of the staged dictionaries with its own tensor; the graph copies it in."""
    m = _model()
    m.use_cuda_graph = True
    st = FrameStaging(1, 7, 192, 256, P=8)
    cur, src = synthetic.make_frame_batch(6300, 1, 7, 192, 256)
    dev = st.device_frame("cuda")
    FrameStaging.upload(st.host_frame().fill(cur, src), dev)
    base = m("test", dev.cur, dev.src, return_mask=True)
    planes = torch.full_like(dev.cur["rendered_depth"], 2.5)
    cur2 = type(dev.cur)(dev.cur)
    cur2.frame = dev.cur.frame
    cur2["rendered_depth"] = planes
    got = m("test", cur2, dev.src, return_mask=True)
    ref_cur = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    ref_cur["rendered_depth"] = planes
    want = m("test", ref_cur, {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
    assert torch.equal(got["pred_0"], want["pred_0"]) and not torch.equal(got["pred_0"], base["pred_0"])


@pytest.mark.parametrize("graph", [False, True])
def test_encoder_ahead_pipeline_is_bit_identical(graph):
    """`FramePipeline(encoder_ahead=True)`: the image-prior encoder of batch i+1 runs under the forward of batch i;
    every batch still gets exactly the results of a direct call."""
    from implicit_depth_b200.pipeline import FramePipeline

    ref_model = _model()
    ref_model.use_cuda_graph = graph
    st = FrameStaging(1, 7, 192, 256, P=8)
    hosts, direct = [], []
    for i in range(5):
        cur, src = synthetic.make_frame_batch(6400 + i, 1, 7, 192, 256)
        o = ref_model("test", {k: torch.from_numpy(v).cuda() for k, v in cur.items()},
                      {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
        direct.append({k: v.cpu().clone() for k, v in o.items()})
        hosts.append(st.host_frame().fill(cur, src))
    m = _model()
    m.use_cuda_graph = graph
    pipe = FramePipeline(m, "cuda", encoder_ahead=True, return_mask=True)
    for rep in range(2):  # second pass: all graphs already captured
        got = [{k: v.clone() for k, v in res.items()} for res in pipe.run(iter(hosts))]
        assert len(got) == 5
        for g_, d_ in zip(got, direct):
            for k in d_:
                assert torch.equal(g_[k], d_[k]), (rep, k)
    # a direct call on a model in encoder-ahead mode runs the encoder inline
    cur, src = synthetic.make_frame_batch(6400, 1, 7, 192, 256)
    o = m("test", {k: torch.from_numpy(v).cuda() for k, v in cur.items()},
          {k: torch.from_numpy(v).cuda() for k, v in src.items()}, return_mask=True)
    assert torch.equal(o["pred_0"].cpu(), direct[0]["pred_0"])
