#!/usr/bin/env python
"""Probe: capture kpd_forward (through model.predict on device-resident static buffers) in a CUDA graph at small batch sizes and
compare latency / results with the eager call."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import torch
from dll import MultiPersonKeypointModel
from dll.configs.config_loader import load_config
from bench import make_inputs

cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
torch.manual_seed(0)
m = MultiPersonKeypointModel(cfg.model, cfg.training, precision="bf16").cuda().eval()
res = {}
for B in (1, 8, 32, 256):
    x, b = make_inputs(B, 224, 1)
    batch = {"image": x.cuda(), "bboxes": b.cuda()}
    for _ in range(3):
        ref = m.predict(batch)
    torch.cuda.synchronize()
    def timeit(fn, n=200):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    eager = timeit(lambda: m.predict(batch))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        m.predict(batch)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g):
            out = m.predict(batch)
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        same = bool(torch.equal(out["keypoints"], ref["keypoints"]) and torch.equal(out["heatmap"], ref["heatmap"]))
        graph = timeit(g.replay)
        res[B] = {"eager_ms": eager, "graph_ms": graph, "bit_identical": same}
    except Exception as e:
        res[B] = {"eager_ms": eager, "error": repr(e)[:300]}
    print(B, res[B], flush=True)
print(json.dumps(res))
