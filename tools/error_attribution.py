#!/usr/bin/env python
"""Where does the reduced-precision heat-map error on the calibrated (peaked) weights come from?

    python tools/error_attribution.py [--out gpurun_out/error_attribution.json]

The calibrated stress weights sharpen the heat-map logits (sigma ~2, bias -6), so that operand rounding anywhere upstream
is amplified by the logit gain before the sigmoid.  This script isolates each rounding source on the GPU, everything else
held at strict fp32 (CUDA-core kernels):
  * head convs one at a time with tf32 operands (round-to-nearest inputs, tf32 weights, fp32 accumulate), with bf16 operands
    and bf16 output storage, and with bf16 STORAGE only (strict conv on a bf16-rounded input) -- kpd_dense_conv leaf op;
  * all three head convs reduced (what the fp32 / bf16 modes run);
  * the composed tf32 FPN-0 alone (fp32-mode FPN-0 map spliced into an otherwise strict forward through kpd_roi_head), and the
    bf16-stored FPN-0 map alone.
Errors are max |delta| / map peak over all heat maps ("range-relative", what tests/test_gpu_parity.py bounds) and the
maximum point-wise relative error over values above 10 % of the map peak.  PyTorch is used for glue only (casts, the 64->17
1x1 + sigmoid of the strict reference chain).
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "keypoint-detection_b200"), ROOT]
import numpy as np
import torch

from dll import MultiPersonKeypointModel, _native as N
from dll.configs.config_loader import load_config
from oracle.weights import synth_state_dict, synth_images, synth_boxes


def tf32_rna(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def errors(a, b):
    a = a.reshape(-1, 56 * 56).double(); b = b.reshape(-1, 56 * 56).double()
    peak = b.abs().max(dim=1, keepdim=True).values
    live = peak[:, 0] > 0
    rr = float(((a - b).abs().max(dim=1).values[live] / peak[live, 0]).max())
    mask = (b > 0.1 * peak) & (peak > 0)
    pw = float(((a - b).abs()[mask] / b[mask]).max())
    return {"range_rel": rr, "pointwise_rel": pw}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "error_attribution.json"))
    args = ap.parse_args()
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    sd = synth_state_dict(0)                                   # calibrated stress weights
    m = MultiPersonKeypointModel(cfg.model, cfg.training, precision="fp32_strict")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    lib = N.load()
    dev = torch.device("cuda", 0)
    B, P = 8, 2
    x = torch.from_numpy(synth_images(B, 1, 224, seed=1234)).cuda()
    boxes = torch.from_numpy(synth_boxes(B, P, seed=99)).cuda()
    batch = {"image": x, "bboxes": boxes}

    # ---- strict reference: heat maps + the head's attention-scaled input (NHWC fp32)
    ref = m.predict(batch)
    st = m.stage_tensors()
    n = int(st["roi_count"].item())
    topk = st["topk"].clone()
    head_in = st["head_in"][:n].clone()
    slot = st["roi_slot"][:n].long()
    ref_hm = ref["heatmap"].reshape(B * P, 17, 56, 56)[slot].clone()
    plan, _ = m._plan()
    wf = sd["heatmap_head.final_layer.3.weight"].view(17, 64).cuda()
    bf = sd["heatmap_head.final_layer.3.bias"].cuda()
    st_ptr = N.stream_ptr(dev)

    def conv(idx, impl, inp):
        """head conv idx (1..3) with implementation impl (0 strict, 1 tf32, 2 bf16) on NHWC `inp` (dtype must match impl)."""
        cout = {1: 256, 2: 256, 3: 64}[idx]
        out = torch.empty(n, 56, 56, cout, device=dev, dtype=torch.bfloat16 if impl == 2 else torch.float32)
        N.check(lib.kpd_dense_conv(plan.handle, idx, impl, N.ptr(inp.contiguous()), N.ptr(out), n, 56, 56, st_ptr), "kpd_dense_conv")
        return out

    def chain(kinds):
        """kinds[k] in {'fp32', 'tf32', 'bf16', 'bf16_storage'} for conv k+1; returns heat maps [n,17,56,56]."""
        a = head_in
        for k, kind in enumerate(kinds):
            if kind == "fp32":
                a = conv(k + 1, 0, a.float())
            elif kind == "tf32":
                a = conv(k + 1, 1, tf32_rna(a.float()))
            elif kind == "bf16":
                a = conv(k + 1, 2, a.to(torch.bfloat16)).float()
            elif kind == "bf16_storage":
                a = conv(k + 1, 0, a.to(torch.bfloat16).float()).to(torch.bfloat16).float()
        logits = a.float().reshape(-1, 64) @ wf.t() + bf
        return torch.sigmoid(logits).view(n, 56, 56, 17).permute(0, 3, 1, 2).contiguous(), logits

    report = {"config": {"batch": B, "persons": P, "rois": n, "weights": "calibrated stress set (oracle.weights.synth_state_dict(0))"}}
    base, logits = chain(["fp32"] * 3)
    report["strict_chain_vs_model_strict"] = errors(base, ref_hm)
    report["logit_stats"] = {"max_abs": float(logits.abs().max()), "std": float(logits.std()),
                             "note": "d sigmoid / d logit <= 0.25: a logit error e moves a heat-map value by up to e / 4"}
    for kind in ("tf32", "bf16", "bf16_storage"):
        for k in range(3):
            kinds = ["fp32"] * 3; kinds[k] = kind
            report[f"head_conv{k + 1}_only/{kind}"] = errors(chain(kinds)[0], base)
        report[f"head_all/{kind}"] = errors(chain([kind] * 3)[0], base)

    # ---- FPN-0 alone: splice the reduced-precision FPN-0 map into an otherwise strict forward
    ws, L, _, S, n_rois, prec = m._last
    strict_fpn0 = st["fpn0"].clone()
    m.set_precision("fp32")
    m.predict(batch, topk_override=topk)
    fpn0_tf32 = m.stage_tensors()["fpn0"].float().clone()
    m.set_precision("bf16")
    m.predict(batch, topk_override=topk)
    fpn0_bf16 = m.stage_tensors()["fpn0"].float().clone()
    m.set_precision("fp32_strict")
    m.predict(batch)                                            # strict workspace again
    ws, L, _, S, n_rois, prec = m._last
    views = m.stage_tensors()

    def with_fpn0(f):
        views["fpn0"].copy_(f)
        hm = torch.empty(B, P, 17, 56, 56, device=dev); kp = torch.empty(B, P, 1, 17, 2, device=dev); vis = torch.empty(B, P, 1, 17, 3, device=dev)
        N.check(lib.kpd_roi_head(plan.handle, N.PREC_FP32_SIMT, B, S, N.ptr(boxes), P, N.ptr(topk), N.ptr(hm), N.ptr(kp), N.ptr(vis),
                                 N.ptr(ws), C.byref(L), st_ptr), "kpd_roi_head")
        return hm.reshape(B * P, 17, 56, 56)[slot]
    same = with_fpn0(strict_fpn0)
    report["fpn0_splice_sanity"] = errors(same, ref_hm)
    peak = float(strict_fpn0.abs().max())
    report["fpn0_only/tf32_composed"] = dict(errors(with_fpn0(fpn0_tf32), ref_hm), fpn0_range_rel=float((fpn0_tf32 - strict_fpn0).abs().max()) / peak)
    report["fpn0_only/tf32_composed_bf16_storage"] = dict(errors(with_fpn0(fpn0_bf16), ref_hm), fpn0_range_rel=float((fpn0_bf16 - strict_fpn0).abs().max()) / peak)

    # ---- the shipped modes end to end (teacher-forced top-k), for reference
    for mode in ("fp32", "bf16"):
        m.set_precision(mode)
        out = m.predict(batch, topk_override=topk)
        report[f"mode/{mode}"] = errors(out["heatmap"].reshape(B * P, 17, 56, 56)[slot], ref_hm)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)
    for k, v in report.items():
        print(k, v)


if __name__ == "__main__":
    main()
