"""GPU parity of the conv networks and the full BDModel forward: B200 modules (tcgen05 conv path) vs the
reference goldens and the CPU oracle, 1e-3 relative (max|diff| / max|ref|)."""
import numpy as np
import pytest
import torch

from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from oracle import networks as ON
from oracle import planesweep as O

from cases import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3


def seeded(**kw):
    m = B200BDModel(default_options(**kw))
    checksum = synthetic.init_model_weights(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda().eval(), checksum, sd


def cuda(xs):
    return [torch.from_numpy(x).cuda() for x in xs]


def test_cv_encoder_and_decoders_vs_reference_golden():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16)
    golden_ok = abs(checksum - float(g["checksum_unet_pp"])) <= 1e-6 * checksum
    enc_c, cv_c = cuda(enc), torch.from_numpy(cv).cuda()
    cvf = m.cost_volume_net(cv_c, enc_c[1:])
    with torch.no_grad():
        ref_cvf = ON.cv_encoder(sd, "cost_volume_net", torch.from_numpy(cv), [torch.from_numpy(e) for e in enc[1:]])
    for i, f in enumerate(cvf):
        assert rel_err(f.cpu().numpy(), ref_cvf[i].numpy()) < TOL
        if golden_ok:
            assert rel_err(f.cpu().numpy(), g[f"cvenc_{i}"]) < TOL
    dec = m.depth_decoder(enc_c[:1] + cvf)
    with torch.no_grad():
        ref_dec = ON.bd_decoder_pp(sd, "depth_decoder", [torch.from_numpy(enc[0])] + ref_cvf)
    for i in range(4):
        got = dec[f"feature_s{i}_b1hw"].cpu().numpy()
        assert rel_err(got, ref_dec[f"feature_s{i}_b1hw"].numpy()) < TOL
        if golden_ok:
            assert rel_err(got, g[f"unetpp_s{i}"]) < TOL


def test_skip_decoder_vs_reference_golden():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16, depth_decoder_name="skip")
    enc_c, cv_c = cuda(enc), torch.from_numpy(cv).cuda()
    cvf = m.cost_volume_net(cv_c, enc_c[1:])
    dec = m.depth_decoder(enc_c[:1] + cvf)
    golden_ok = abs(checksum - float(g["checksum_skip"])) <= 1e-6 * checksum
    with torch.no_grad():
        ref_cvf = ON.cv_encoder(sd, "cost_volume_net", torch.from_numpy(cv), [torch.from_numpy(e) for e in enc[1:]])
        ref = ON.skip_decoder(sd, "depth_decoder", [torch.from_numpy(enc[0])] + ref_cvf)
    for i in range(4):
        got = dec[f"feature_s{i}_b1hw"].cpu().numpy()
        assert rel_err(got, ref[f"feature_s{i}_b1hw"].numpy()) < TOL
        if golden_ok:
            assert rel_err(got, g[f"skip_s{i}"]) < TOL


def test_matching_encoder_vs_reference_and_batch_invariance():
    g = np.load(f"{GOLDEN}/nets_96x128.npz")
    enc, cv, img = synthetic.make_net_inputs(3000)
    m, checksum, sd = seeded(image_width=128, image_height=96, matching_num_depth_bins=16)
    x = torch.from_numpy(img).cuda()
    got = m.matching_model(x)
    with torch.no_grad():
        ref = ON.matching_encoder(sd, "matching_model", torch.from_numpy(img))
    assert rel_err(got.cpu().numpy(), ref.numpy()) < TOL
    if abs(checksum - float(g["checksum_unet_pp"])) <= 1e-6 * checksum:
        assert rel_err(got.cpu().numpy(), g["matching"]) < TOL
    # the reference runs this encoder one image at a time because batched cuDNN differs in low bits
    # (depth_model.py:235-241); here a frame's features are bit-identical whatever the batch
    rng = np.random.default_rng(1)
    other = torch.from_numpy(rng.standard_normal((2, 3, 96, 128)).astype(np.float32)).cuda()
    both = m.matching_model(torch.cat([other[:1], x, other[1:]], 0))
    assert torch.equal(both[1:2], got)


@pytest.mark.parametrize("fv_type,K", [("mlp_feature_volume", 7), ("simple_cost_volume", 2)])
def test_full_forward_vs_reference_golden(fv_type, K):
    """BASELINE config 1 (256x192, random weights): the whole B200BDModel.forward against the reference
    BDModel run with the same state dict (golden) and the CPU oracle."""
    g = np.load(f"{GOLDEN}/model_256x192_{fv_type}.npz")
    m, checksum, sd = seeded(image_width=256, image_height=192, matching_num_depth_bins=16,
                             feature_volume_type=fv_type, num_source_views=K)
    cur, src = synthetic.make_frame_batch(4000 + K, 1, K, 192, 256)
    cur_c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    src_c = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    out = m("test", cur_c, src_c, unbatched_matching_encoder_forward=True, return_mask=True)
    assert set(out) == {"pred_0", "lowest_cost_bhw", "overall_mask_bhw"}
    assert tuple(out["pred_0"].shape) == (1, 8, 96, 128)
    planes = O.generate_depth_planes(0.25, 5.0, 16)
    to_idx = lambda z: np.abs(np.log(z)[..., None] - np.log(planes)).argmin(-1)
    if abs(checksum - float(g["checksum"])) <= 1e-6 * checksum:
        # note: the image-prior encoder runs through cuDNN on the GPU (TF32 disabled below) vs CPU in the golden
        assert rel_err(out["pred_0"].cpu().numpy(), g["pred_0"]) < TOL
        assert (to_idx(out["lowest_cost_bhw"].cpu().numpy()) != to_idx(g["lowest_cost_bhw"])).mean() < 5e-3
        if "overall_mask_bhw" in g:
            assert (out["overall_mask_bhw"].cpu().numpy() != g["overall_mask_bhw"]).mean() < 2e-3
    else:
        cpu = B200BDModel(m.run_opts)
        cpu.load_state_dict(sd)
        ref = ON.bd_forward(sd, cpu.encoder.eval(), {k: torch.from_numpy(v) for k, v in cur.items()},
                            {k: torch.from_numpy(v) for k, v in src.items()}, m.run_opts, feature_volume=fv_type)
        assert rel_err(out["pred_0"].cpu().numpy(), ref["pred_0"].numpy()) < TOL
    # CUDA-graph replay gives the same answer as eager launches
    m.use_cuda_graph = True
    out2 = m("test", cur_c, src_c, return_mask=True)
    out3 = m("test", cur_c, src_c, return_mask=True)
    assert torch.equal(out2["pred_0"], out3["pred_0"])
    assert rel_err(out2["pred_0"].cpu().numpy(), out["pred_0"].cpu().numpy()) < 1e-5


@pytest.fixture(autouse=True)
def _strict_fp32_library_convs():
    """The out-of-scope image encoder runs through cuDNN; keep it in strict fp32 so the comparison with the
    CPU golden measures OUR kernels, not cuDNN's TF32 default."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
