#!/usr/bin/env python
"""Micro-benchmarks for BASELINE configs C4 (heat-map decode, 65,536 x 17 x 64 x 48) and C5 (NMS, 128 x 100).

    python tools/bench_micro.py [--instances 65536] [--out gpurun_out/micro.json]

Decode is HBM-bound: algorithmic bytes = N*17*64*48*elt read + 476 B/instance written; the achieved
GB/s is reported against MEASURED_PEAKS.json.  Inputs (13.7 GB fp32) exceed L2 by two orders of magnitude.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import numpy as np
import torch

from dll.models import heatmap_head as HH
from dll.models.person_head import PERSON_HEAD
from dll.configs import PersonDetectionConfig
from oracle import decode as OD
from oracle.weights import synth_blob_heatmaps, synth_nms_candidates


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=65536)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "micro.json"))
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    res = {"hbm_peak_gbs": peaks["hbm_gbs"]}
    N, K, H, W = args.instances, 17, 64, 48
    # Gaussian blobs generated on the device in chunks (same distribution as oracle.weights.synth_blob_heatmaps)
    g = torch.Generator(device="cuda").manual_seed(7)
    hm = torch.empty(N, K, H, W, device="cuda", dtype=torch.float32)
    ys = torch.arange(H, device="cuda").view(1, 1, H, 1).float(); xs = torch.arange(W, device="cuda").view(1, 1, 1, W).float()
    for s in range(0, N, 4096):
        n = min(4096, N - s)
        cx = torch.rand(n, K, 1, 1, device="cuda", generator=g) * (W - 1); cy = torch.rand(n, K, 1, 1, device="cuda", generator=g) * (H - 1)
        amp = 0.2 + 0.8 * torch.rand(n, K, 1, 1, device="cuda", generator=g)
        hm[s:s + n] = amp * torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / 8.0) + 0.01 * torch.rand(n, K, H, W, device="cuda", generator=g)
    for name, t in (("fp32", hm), ("bf16", hm.to(torch.bfloat16))):
        elt = t.element_size()
        algo_bytes = N * K * H * W * elt + N * 476
        ms_sub = ev_time(lambda: HH._decode(t, {"argmax", "kp_subpix", "score_subpix", "vis"}), 10)
        ms_all = ev_time(lambda: HH.decode_heatmaps_full(t), 10)
        res[f"decode_{name}"] = {
            "instances": N, "ms_argmax_subpixel_vis": ms_sub, "instances_per_s": N / ms_sub * 1e3,
            "achieved_gbs": algo_bytes / ms_sub / 1e6, "frac_of_hbm_peak": algo_bytes / ms_sub / 1e6 / peaks["hbm_gbs"],
            "ms_all_decoders_incl_softargmax": ms_all, "achieved_gbs_all": algo_bytes / ms_all / 1e6}
    # parity spot check on a slice + CPU baseline (reference algorithm, python loop) on 512 instances
    sl = hm[:512].cpu().numpy()
    t0 = time.perf_counter(); kp, sc, idx = OD.decode_heatmaps_subpixel(sl); cpu_s = time.perf_counter() - t0
    full = HH.decode_heatmaps_full(hm[:512])
    assert np.array_equal(full["argmax"].cpu().numpy(), idx)
    assert np.allclose(full["kp_subpix"].cpu().numpy(), kp, atol=2e-6)
    res["decode_cpu_port"] = {"instances": 512, "seconds": cpu_s, "instances_per_s": 512 / cpu_s, "cores": 1,
                              "note": "oracle.decode.decode_heatmaps_subpixel (per-joint Python loop like heatmap_head.py:298-370)"}
    del hm
    # NMS C5
    b, s = synth_nms_candidates(128, 100, seed=5)
    ph = PERSON_HEAD(PersonDetectionConfig())
    bt, st = torch.from_numpy(b).cuda(), torch.from_numpy(s).cuda()
    ms = ev_time(lambda: ph.batched_nms(bt, st, 0.3, 100), 50)
    t0 = time.perf_counter()
    for i in range(16):
        OD.non_max_suppression(b[i], s[i], 0.3, 100)
    cpu = (time.perf_counter() - t0) / 16
    res["nms"] = {"images": 128, "candidates": 100, "ms_per_batch": ms, "us_per_image": ms * 1e3 / 128,
                  "images_per_s": 128 / ms * 1e3, "cpu_port_ms_per_image": cpu * 1e3}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
