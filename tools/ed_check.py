#!/usr/bin/env python
"""Trunk check for the fused expand+depthwise kernel: fp32-mode trunk (fused) against the strict CUDA-core trunk.

    python tools/ed_check.py [B] [S]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel, _native
from dll.configs.config_loader import load_config
from oracle.weights import synth_state_dict, synth_images

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
m = MultiPersonKeypointModel(cfg.model, cfg.training)
m.load_state_dict(synth_state_dict(0), strict=True)
m = m.cuda().eval()
x = torch.from_numpy(synth_images(B, 1, S, seed=11)).cuda()
boxes = torch.tensor([0.5, 0.5, 1.0, 1.0]).repeat(B, 1, 1).cuda()
feats = {}
for mode in ("fp32_strict", "fp32"):
    m.set_precision(mode)
    m.predict({"image": x, "bboxes": boxes})
    torch.cuda.synchronize()
    st = m.stage_tensors()
    feats[mode] = [st[f"feat{i}"].clone() for i in range(4)] + [st["scores"].clone()]
for i in range(5):
    a, b = feats["fp32_strict"][i], feats["fp32"][i]
    print(f"{'feat%d' % i if i < 4 else 'scores'}: max|strict| {float(a.abs().max()):.4g}  max abs diff {float((a - b).abs().max()):.3e}  rel-range {float((a - b).abs().max() / a.abs().max()):.3e}", flush=True)
lib = _native.load()
import ctypes as C
lib.kpd_profile_enable(2)
for _ in range(5):
    m.predict({"image": x, "bboxes": boxes})
torch.cuda.synchronize()
for tag in ("body", "body.stem", "body.expand_dw", "body.pointwise", "body.depthwise", "body.se") + tuple(f"ed.f{i}" for i in range(2, 12)) + ("fpn_laterals",):
    t, n = C.c_double(0), C.c_int64(0)
    if lib.kpd_profile_read(tag.encode(), C.byref(t), C.byref(n)) == 0 and n.value:
        print(f"{tag}: {t.value / 5:.4f} ms/step over {n.value // 5} launches")
