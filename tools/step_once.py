#!/usr/bin/env python
"""Profiling driver: W warm-up forwards + K forwards of the bench workload (C2: B=256, 224x224, P=1) and nothing else, so that
an ncu pass can address the launches of one step by index.  Prints the launches per forward.

    python tools/step_once.py [--precision bf16] [--batch 256] [--size 224] [--persons 1] [--warmup 3] [--steps 1]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel, _native
from dll.configs.config_loader import load_config
from bench import make_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16"); ap.add_argument("--batch", type=int, default=256); ap.add_argument("--size", type=int, default=224)
ap.add_argument("--persons", type=int, default=1); ap.add_argument("--warmup", type=int, default=3); ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
torch.manual_seed(0)
m = MultiPersonKeypointModel(cfg.model, cfg.training, precision=a.precision).cuda().eval()
x, b = make_inputs(a.batch, a.size, a.persons)
batch = {"image": x.cuda(), "bboxes": b.cuda()}
lib = _native.load()
for _ in range(a.warmup):
    m.predict(batch)
torch.cuda.synchronize()
lib.kpd_launch_count(1)
for _ in range(a.steps):
    m.predict(batch)
torch.cuda.synchronize()
print("launches_per_step", int(lib.kpd_launch_count(1)) // a.steps)
