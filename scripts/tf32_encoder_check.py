"""Does the out-of-scope cuDNN image encoder in TF32 (PyTorch's default) keep the full forward inside the 1e-3 parity
budget?  Compares pred_0 against the CPU oracle with cudnn.allow_tf32 on and off (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from implicit_depth_b200 import synthetic
from implicit_depth_b200.bd_model import B200BDModel, default_options
from oracle import networks as ON  # checker

torch.set_grad_enabled(False)
for (W, H, D) in ((256, 192, 16), (512, 384, 64)):
    opts = default_options(image_width=W, image_height=H, matching_num_depth_bins=D)
    model = B200BDModel(opts)
    synthetic.init_model_weights(model, seed=0)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cpu_enc = B200BDModel(opts).encoder
    cpu_enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    cur, src = synthetic.make_frame_batch(4007, 1, 7, H, W)
    ref = ON.bd_forward(sd, cpu_enc.eval(), {k: torch.from_numpy(v) for k, v in cur.items()},
                        {k: torch.from_numpy(v) for k, v in src.items()}, opts, torch_volume=True)
    model = model.cuda().eval()
    c = {k: torch.from_numpy(v).cuda() for k, v in cur.items()}
    s = {k: torch.from_numpy(v).cuda() for k, v in src.items()}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        out = model("test", c, s, return_mask=True)
        e = (out["pred_0"].cpu() - ref["pred_0"]).abs().max().item() / ref["pred_0"].abs().max().item()
        print(f"{W}x{H} D={D} cudnn.allow_tf32={tf32}: pred_0 max-abs-err / max-abs-ref = {e:.3e}")
