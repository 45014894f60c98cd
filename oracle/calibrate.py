#!/usr/bin/env python
"""Derive oracle/calibration.npz (test infrastructure): BN running statistics and three logit scales
that keep the *synthetic* network's activations O(1) and its sigmoids unsaturated.

    python -m oracle.calibrate

The raw PCG weights of `synth_state_dict(calibrated=False)` are pushed through the oracle forward on a
seeded calibration batch; every BatchNorm's running_mean/var is set to the batch statistics of its input
(what a trained network would carry), the channel-/spatial-attention logits are scaled to std 1.5 and the
heat-map logits to std 3.  The resulting table (~55 KB) is committed so that `synth_state_dict()` is a
pure function of (seed, table) on every machine.
"""
import os
import numpy as np
import torch

from . import model as M
from .weights import synth_state_dict, synth_images, synth_boxes


@torch.no_grad()
def main(out_path=None):
    sd = synth_state_dict(seed=0, in_channels=1, calibrated=False)
    nb = 32
    x = torch.from_numpy(synth_images(nb, 1, 224, seed=4321))
    boxes = torch.from_numpy(synth_boxes(nb, 1, seed=17))
    M._CALIBRATE = {}
    try:
        feats = M.body_forward(sd, x)
        fpn0 = M.fpn_forward(sd, feats, levels=(0, 1, 2, 3))[0]
        scores = M.channel_scores(sd, "channel_attention", fpn0)
        idx = M.select_top_k(scores, 64)
        sel = fpn0[torch.arange(nb)[:, None], idx]
        rois = torch.cat([M.extract_roi_features(sel[b:b + 1], boxes[b, p]) for b in range(nb) for p in range(1)], 0)
        M.heatmap_head_forward(sd, rois)
        table = {k: (v.numpy() if torch.is_tensor(v) else np.asarray([v], dtype=np.float64))
                 for k, v in M._CALIBRATE.items()}
    finally:
        M._CALIBRATE = None
    out_path = out_path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "calibration.npz")
    np.savez_compressed(out_path, **table)
    print("wrote", out_path, len(table), "entries")


if __name__ == "__main__":
    main()
