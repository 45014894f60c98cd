#!/usr/bin/env python
"""Host-side cost of enqueueing one forward (no synchronisation inside the loop) against its GPU time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
from bench import make_inputs
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
torch.manual_seed(0)
m = MultiPersonKeypointModel(cfg.model, cfg.training, precision="bf16").cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x, b = make_inputs(B, 224, 1)
batch = {"image": x.cuda(), "bboxes": b.cuda()}
for _ in range(5):
    m.predict(batch)
torch.cuda.synchronize()
for n in (1, 20):
    t0 = time.perf_counter()
    for _ in range(n):
        m.predict(batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"B={B} n={n}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step")
