"""GPU diagnostic: stage-by-stage difference between the strict-fp32 SIMT path and the tcgen05 paths."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import numpy as np, torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
from oracle.weights import synth_state_dict, synth_images, synth_boxes

cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
m = MultiPersonKeypointModel(cfg.model, cfg.training); m.load_state_dict(synth_state_dict(0)); m = m.cuda().eval()
x = torch.from_numpy(synth_images(2, 1, 224, seed=1234)); b = torch.from_numpy(synth_boxes(2, 3, seed=99, zero_slots=((0, 1),)))
m.set_precision("fp32_strict")
out0 = m.predict({"image": x, "bboxes": b}); torch.cuda.synchronize()
s0 = {k: v.float().clone() for k, v in m.stage_tensors().items()}
h0 = out0["heatmap"].clone()
n = int(s0["roi_count"].item())
for mode in ("fp32", "bf16"):
    m.set_precision(mode)
    out = m.predict({"image": x, "bboxes": b}, topk_override=s0["topk"]); torch.cuda.synchronize()
    s = m.stage_tensors()
    print("==", mode)
    for k in ("lateral0", "fpn0", "roi_feat", "head_in", "head_a", "head_b", "head_c"):
        a, r = s[k].float(), s0[k]
        if k.startswith(("roi", "head")):
            a, r = a[:n], r[:n]
        d = (a - r).abs()
        peak = r.abs().max().item()
        idx = np.unravel_index(int(d.argmax().item()), d.shape)
        rms = float(d.pow(2).mean().sqrt() / r.pow(2).mean().sqrt())
        # error by position class: border rows/cols vs interior
        H, W = d.shape[1], d.shape[2]
        border = torch.zeros(H, W, dtype=torch.bool, device=d.device); border[0] = border[-1] = True; border[:, 0] = border[:, -1] = True
        eb = d[:, border].max().item() / peak; ei = d[:, ~border].max().item() / peak
        nan = int(torch.isnan(a).sum().item())
        print(f"{k:9s} max_err/peak {d.max().item()/peak:.3e} rms_rel {rms:.3e} border {eb:.3e} interior {ei:.3e} argmax {idx} nan {nan}")
    d = (out["heatmap"] - h0).abs()
    print("heatmap   max abs", d.max().item(), "mean abs", d.mean().item(), "kp diff", (out["keypoints"] - out0["keypoints"]).abs().max().item())
