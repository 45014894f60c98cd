#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

It imports `dll` from /root/reference (never from this repo), applies the single oracle-side
patch SURVEY.md 8c requires (`mobilenet_v3_small(weights=None)`; the stock call downloads
ImageNet weights), loads the deterministic synthetic state_dict from `oracle.weights`
with strict=True (which pins the 318-key / shape contract), runs the reference's own
`MultiPersonKeypointModel.forward`, decoders and NMS on seeded inputs, and writes small
.npz files.  Big stage tensors are stored as strided sub-samples + float64 moment checksums.
"""
import os
import sys

REF = os.environ.get("KPD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)          # for `oracle` only; the product package is NOT on the path

import numpy as np
import torch
import torchvision

import dll.models.backbone as _bb
from torchvision.models import mobilenet_v3_small as _tv_mbv3
_bb.mobilenet_v3_small = lambda weights=None, **kw: _tv_mbv3(weights=None, **kw)

from dll.configs.config_loader import load_config
from dll.models.keypoint_model import MultiPersonKeypointModel, select_top_k_channels
from dll.models import heatmap_head as ref_hh
from dll.models.person_head import PERSON_HEAD

from oracle.weights import (synth_state_dict, synth_images, synth_boxes, synth_blob_heatmaps,
                            synth_nms_candidates)
from oracle.golden_util import sample_big, checksum, GOLDEN_CASES

assert os.path.realpath(sys.modules["dll"].__file__).startswith(os.path.realpath(REF)), "wrong dll imported"
torch.set_num_threads(os.cpu_count())


def build_reference_model():
    cfg = load_config(os.path.join(REF, "configs", "default_config.yaml"))
    model = MultiPersonKeypointModel(cfg.model, cfg.training)
    sd = synth_state_dict(seed=0, in_channels=cfg.model.backbone.in_channels)
    ref_keys = list(model.state_dict().keys())
    assert ref_keys == list(sd.keys()), "state_dict key order/contents differ from the oracle spec"
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape) and v.dtype == sd[k].dtype, k
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, cfg, ref_keys


@torch.no_grad()
def run_case(model, name, batch, size, persons, zero_slots):
    x = torch.from_numpy(synth_images(batch, 1, size, seed=1234))
    boxes = torch.from_numpy(synth_boxes(batch, persons, seed=99, zero_slots=zero_slots))
    cap = {}

    def hook(tag):
        def fn(mod, inp, out):
            cap.setdefault(tag, []).append(out)
        return fn
    hs = [model.backbone.body.register_forward_hook(hook("body")),
          model.backbone.fpn.register_forward_hook(hook("fpn")),
          model.channel_attention.register_forward_hook(hook("scores")),
          model.heatmap_head.register_forward_hook(hook("head")),
          model.heatmap_head.register_forward_pre_hook(lambda m, i: cap.setdefault("roi", []).append(i[0]))]
    out = model({"image": x, "bboxes": boxes})
    for h in hs:
        h.remove()
    feats = cap["body"][0]
    fpn0 = cap["fpn"][0][0]
    scores = cap["scores"][0]
    topk = torch.topk(scores, 64, dim=1).indices
    d = {
        "feat0_s": sample_big(feats["feat0"]), "feat0_c": checksum(feats["feat0"]),
        "feat1_s": sample_big(feats["feat1"]), "feat1_c": checksum(feats["feat1"]),
        "feat2_s": sample_big(feats["feat2"]), "feat2_c": checksum(feats["feat2"]),
        "feat3_s": sample_big(feats["feat3"]), "feat3_c": checksum(feats["feat3"]),
        "fpn0_s": sample_big(fpn0), "fpn0_c": checksum(fpn0),
        "scores": scores.numpy(), "topk": topk.numpy().astype(np.int64),
        "keypoints": out["keypoints"].numpy(), "visibilities": out["visibilities"].numpy(),
        "heatmap_shape": np.asarray(out["heatmap"].shape), "heatmap_c": checksum(out["heatmap"]),
        "boxes": boxes.numpy(),
    }
    if "roi" in cap:
        roi = torch.cat(cap["roi"], 0)
        d["roi_s"] = sample_big(roi); d["roi_c"] = checksum(roi)
        d["cw"] = torch.cat([h[1][0] for h in cap["head"]], 0).numpy()
        d["sw"] = torch.cat([h[1][1] for h in cap["head"]], 0).numpy()
    if size == 224:
        d["heatmap"] = out["heatmap"].numpy()
    else:
        d["heatmap_s"] = sample_big(out["heatmap"].reshape(-1, 17, 56, 56))
    hm = out["heatmap"].reshape(-1, 17, 56, 56)
    d["heatmap_argmax"] = hm.reshape(hm.shape[0], 17, -1).argmax(dim=2).numpy().astype(np.int64)
    print(f"[{name}] heatmap range {float(hm.min()):.4f}..{float(hm.max()):.4f} "
          f"vis classes {out['visibilities'].sum(dim=(0, 1, 2, 3)).tolist()} "
          f"min top-65 score gap {float((-torch.sort(scores, dim=1, descending=True).values.diff(dim=1))[:, :64].min()):.3e}")
    np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **d)


@torch.no_grad()
def run_stock(model, calibrated_sd):
    """Reference-style random init ('stock'): flat heat maps and near-tied channel scores."""
    model.load_state_dict(synth_state_dict(seed=0, style="stock"), strict=True)
    x = torch.from_numpy(synth_images(2, 1, 224, seed=1234))
    boxes = torch.from_numpy(synth_boxes(2, 3, seed=99, zero_slots=((0, 1),)))
    cap = {}
    h = model.channel_attention.register_forward_hook(lambda m, i, o: cap.setdefault("scores", o))
    out = model({"image": x, "bboxes": boxes})
    h.remove()
    scores = cap["scores"]
    srt = torch.sort(scores, dim=1, descending=True).values
    print(f"[stock] heatmap range {float(out['heatmap'][out['heatmap'] > 0].min()):.4f}..{float(out['heatmap'].max()):.4f} "
          f"min top-65 score gap {float((srt[:, :64] - srt[:, 1:65]).min()):.3e}")
    np.savez_compressed(os.path.join(HERE, "forward_stock_s224_b2_p3.npz"), scores=scores.numpy(),
                        topk=torch.topk(scores, 64, dim=1).indices.numpy().astype(np.int64),
                        heatmap=out["heatmap"].numpy(), keypoints=out["keypoints"].numpy(),
                        visibilities=out["visibilities"].numpy(), boxes=boxes.numpy())
    model.load_state_dict(calibrated_sd, strict=True)


@torch.no_grad()
def run_degenerate(model):
    """bboxes=None / every list empty -> the eval-mode 'empty_output' shapes (keypoint_model.py:123-135)."""
    x = torch.from_numpy(synth_images(2, 1, 224, seed=1234))
    out = model({"image": x, "bboxes": None})
    np.savez_compressed(os.path.join(HERE, "forward_no_boxes.npz"),
                        keypoints_shape=np.asarray(out["keypoints"].shape),
                        visibilities_shape=np.asarray(out["visibilities"].shape),
                        heatmap_shape=np.asarray(out["heatmap"].shape),
                        total=np.asarray([float(out["keypoints"].abs().sum() + out["visibilities"].abs().sum()
                                                + out["heatmap"].abs().sum())]))


@torch.no_grad()
def run_decoders(model):
    hm = synth_blob_heatmaps(6, 17, 64, 48, seed=7)
    hm[0, 0] = 0.0                      # zero-mass window -> (0,0,0) branch of the sub-pixel decoder
    hm[1, 3, 0, 0] = 2.0                # corner maximum -> clipped 2x2 window
    hm[2, 5, 63, 47] = 2.0
    hm[3, 7, 10, 20] = hm[3, 7].max() + 0.5
    hm[3, 7, 40, 5] = hm[3, 7, 10, 20]  # exact tie -> first index wins
    t = torch.from_numpy(hm)
    k1, s1 = ref_hh.decode_heatmaps(t)
    k2, s2 = ref_hh.decode_heatmaps_subpixel(t)
    k3, s3 = ref_hh.decode_heatmaps_soft_argmax(t, temperature=0.05)
    k4, v4 = model.decode_heatmap(t)
    k5, s5 = ref_hh.decode_heatmaps(t[0])          # 3-D input form
    amax = t.reshape(6, 17, -1).max(dim=2).indices
    np.savez_compressed(os.path.join(HERE, "decoders.npz"), heatmaps=hm,
                        argmax=amax.numpy().astype(np.int64),
                        plain_kp=k1.numpy(), plain_sc=s1.numpy(),
                        subpix_kp=k2.numpy(), subpix_sc=s2.numpy(),
                        soft_kp=k3.numpy(), soft_sc=s3.numpy(), soft_temperature=np.asarray([0.05]),
                        model_kp=k4.numpy(), model_vis=v4.numpy(),
                        plain3d_kp=k5.numpy(), plain3d_sc=s5.numpy())


@torch.no_grad()
def run_nms(model):
    boxes, scores = synth_nms_candidates(8, 100, seed=5)
    ph = model.person_detector
    keeps, counts = np.full((8, 100), -1, dtype=np.int64), np.zeros(8, dtype=np.int64)
    for i in range(8):
        k = ph.non_max_suppression(torch.from_numpy(boxes[i]), torch.from_numpy(scores[i]), 0.3, 100).numpy()
        keeps[i, :len(k)] = k; counts[i] = len(k)
    k20 = ph.non_max_suppression(torch.from_numpy(boxes[0]), torch.from_numpy(scores[0]), 0.5, 20).numpy()
    iou = ph.box_iou(torch.from_numpy(boxes[0][:16]), torch.from_numpy(boxes[1][:24])).numpy()
    np.savez_compressed(os.path.join(HERE, "nms.npz"), boxes=boxes, scores=scores, keep=keeps, count=counts,
                        keep_thr05_max20=k20.astype(np.int64), iou_16x24=iou)


@torch.no_grad()
def run_roi_align():
    g = np.random.default_rng(42)
    feat = g.standard_normal((2, 8, 64, 64)).astype(np.float32)
    # x1,y1,x2,y2 in feature pixels: tiny (<1px -> clamped to 1), medium, full map, >56px (2x2 grid), 3x3 grid
    rois = np.asarray([[0, 10.2, 11.7, 10.5, 12.1], [0, 3.3, 4.1, 40.9, 35.2], [1, 0, 0, 64, 64],
                       [1, 1.5, 2.5, 63.5, 61.0], [0, 20.0, 0.0, 64.0, 30.0]], dtype=np.float32)
    out = torchvision.ops.roi_align(torch.from_numpy(feat), torch.from_numpy(rois), (56, 56))
    big = g.standard_normal((1, 4, 256, 256)).astype(np.float32)
    rois_big = np.asarray([[0, 0, 0, 256, 256], [0, 17.3, 30.1, 200.4, 140.9]], dtype=np.float32)   # 5x5, 4x2 grids
    out_big = torchvision.ops.roi_align(torch.from_numpy(big), torch.from_numpy(rois_big), (56, 56))
    np.savez_compressed(os.path.join(HERE, "roi_align.npz"), feat=feat, rois=rois, out=out.numpy(),
                        big_seed=np.asarray([42]), rois_big=rois_big, out_big=out_big.numpy())


@torch.no_grad()
def run_person_heads(model):
    g = np.random.default_rng(11)
    x = g.standard_normal((2, 128, 12, 12)).astype(np.float32)
    ph = model.person_detector
    box = ph.box_heads[0](torch.from_numpy(x)).numpy()
    cls = ph.cls_heads[0](torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(HERE, "person_heads.npz"), x=x, box=box, cls=cls)


def main():
    model, cfg, keys = build_reference_model()
    with open(os.path.join(HERE, "state_dict_keys.txt"), "w") as f:
        for k, v in model.state_dict().items():
            f.write(f"{k} {list(v.shape)} {str(v.dtype).replace('torch.', '')}\n")
    for name, kw in GOLDEN_CASES.items():
        run_case(model, name, **kw)
    run_stock(model, synth_state_dict(seed=0))
    run_degenerate(model)
    run_decoders(model)
    run_nms(model)
    run_roi_align()
    run_person_heads(model)
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write(f"generated by tests/golden/make_golden.py from {REF}\n"
                f"torch {torch.__version__} torchvision {torchvision.__version__} threads {torch.get_num_threads()}\n"
                f"patch: dll.models.backbone.mobilenet_v3_small -> weights=None (SURVEY.md 8c)\n")
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
