#!/usr/bin/env python
"""ITransform throughput: kpd_itransform on the GPU against the reference's OpenCV / Pillow path on the host.

    python tools/bench_itransform.py [--batch 256] [--height 480] [--width 640] [--steps 20]

Prints one JSON line: images/s resident (uint8 batch already on the device), end to end (pinned H2D of the uint8 batch inside
the timed region), per-kernel-class times from the library's event tags, and the host pipeline's images/s on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "keypoint-detection_b200"), ROOT):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-images", type=int, default=64)
    a = ap.parse_args()
    from dll.data import ITransform
    B, H, W = a.batch, a.height, a.width
    def synth_photo(h, w, seed):                 # smooth illumination + hard-edged shapes + sensor noise
        rng = np.random.default_rng(seed)
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        img = np.stack([120 + 70 * np.sin(xx / (19.0 + 3 * c) + seed) * np.cos(yy / (27.0 - 2 * c)) for c in range(3)], axis=-1)
        for k in range(6):
            cy, cx, ry, rx = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(h / 12, h / 4), rng.uniform(w / 12, w / 4)
            img[(np.abs(yy - cy) < ry) & (np.abs(xx - cx) < rx)] += rng.uniform(-80, 80, 3)
        return np.clip(img + rng.normal(0, 5, img.shape), 0, 255).astype(np.uint8)

    uniq = [synth_photo(H, W, 500 + i) for i in range(16)]
    host = torch.from_numpy(np.stack([uniq[i % 16] for i in range(B)])).pin_memory()
    dev = host.cuda()
    t = ITransform(224, 2.0, (8, 8), True, device="cuda")
    for _ in range(a.warmup):
        t.batch(dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = t.batch(dev)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    e0.record()
    for _ in range(a.steps):
        out = t.batch(host.cuda(non_blocking=True))
    e1.record(); torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / a.steps
    # host reference path (cv2 + PIL), bounded sample
    host_t = ITransform(224, 2.0, (8, 8), True)
    n = min(a.cpu_images, B)
    t0 = time.perf_counter()
    for i in range(n):
        host_t(uniq[i % 16])
    cpu_s = (time.perf_counter() - t0) / n
    algo = B * (H * W * 3 + 224 * 224 * 4)
    print(json.dumps({"metric": "images/sec (ITransform grayscale pipeline)", "batch": B, "height": H, "width": W,
                      "value": B / ms * 1e3, "ms_per_batch": ms, "e2e": {"value": B / ms_e2e * 1e3, "h2d_bytes": int(host.numel())},
                      "algorithmic_GBps": algo / ms / 1e6,
                      "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "images/sec", "cores": os.cpu_count(), "kind": "reference libraries (cv2+PIL), one image at a time",
                                       "sample": f"{n} images"}}))


if __name__ == "__main__":
    main()
