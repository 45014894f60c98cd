"""CPU restatement of the reference inference forward (test infrastructure; torch-CPU fp32).

Follows, stage by stage:
  body     dll/models/backbone.py:247-264 + tv:models/mobilenetv3.py:53-116,224-266 + tv:ops/misc.py:225-261
  FPN      dll/models/backbone.py:29-39
  top-k    dll/models/keypoint_model.py:18-44 (ChannelAttention), :653-661 (select_top_k_channels)
  ROI      dll/models/keypoint_model.py:212-228, :630-638 + tv:ops/roi_align.py (legacy aligned=False,
           adaptive sampling_ratio), restated from the published detectron/torchvision kernel
  head     dll/models/heatmap_head.py:81-151
  decode   dll/models/keypoint_model.py:250-313
  assemble dll/models/keypoint_model.py:121-206, :230-248, :640-651

The state_dict `sd` is a plain {name: torch.Tensor} with the reference's 318 keys.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

from .arch import MBV3_SMALL_BLOCKS, BODY_BN_EPS, HEAD_BN_EPS


# ----------------------------------------------------------------------------- helpers
# Calibration hook used ONLY by oracle/calibrate.py to derive the committed oracle/calibration.npz
# (BN running statistics / logit scales that keep the synthetic network's activations O(1)).
_CALIBRATE = None


def _calib_scale(sd, z, wkey, bkey, target_std, tag):
    """calibration only: rescale the producing layer so that std(z) == target_std."""
    s = float(target_std / z.std())
    sd[wkey] = sd[wkey] * s
    if bkey is not None:
        sd[bkey] = sd[bkey] * s
    _CALIBRATE["scale:" + tag] = s
    return s


def _bn(x, sd, prefix, eps):
    if _CALIBRATE is not None:
        # measured statistics, deliberately de-tuned per channel (var x U(0.5,1.5), mean +- 0.1 sigma) so the
        # folded BN is a non-trivial affine map; downstream layers are calibrated on the de-tuned output.
        from .weights import _rng
        g = _rng(0, "calib:" + prefix)
        v = x.var(dim=(0, 2, 3), unbiased=False)
        m = x.mean(dim=(0, 2, 3))
        sd[prefix + ".running_var"] = v * torch.from_numpy(g.uniform(0.5, 1.5, v.shape).astype(np.float32))
        sd[prefix + ".running_mean"] = m + v.sqrt() * torch.from_numpy((0.1 * g.standard_normal(v.shape)).astype(np.float32))
        _CALIBRATE["mean:" + prefix] = sd[prefix + ".running_mean"]
        _CALIBRATE["var:" + prefix] = sd[prefix + ".running_var"]
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, eps)


def _act(x, kind):
    if kind == "RE":
        return F.relu(x)
    if kind == "HS":
        return F.hardswish(x)
    return x


# ----------------------------------------------------------------------------- body (A1)
def body_forward(sd, x, return_all=False):
    """x [B,Cin,S,S] -> (feat0, feat1, feat2, feat3) = outputs of features.0/.3/.8/.12."""
    p = "backbone.body.features."
    y = F.conv2d(x, sd[p + "0.0.weight"], None, stride=2, padding=1)
    y = F.hardswish(_bn(y, sd, p + "0.1", BODY_BN_EPS))
    outs = {0: y}
    for (i, cin, k, exp, cout, use_se, act, stride) in MBV3_SMALL_BLOCKS:
        b = f"{p}{i}.block."
        j = 0
        inp = y
        if exp != cin:
            y = F.conv2d(y, sd[f"{b}{j}.0.weight"])
            y = _act(_bn(y, sd, f"{b}{j}.1", BODY_BN_EPS), act)
            j += 1
        y = F.conv2d(y, sd[f"{b}{j}.0.weight"], None, stride=stride, padding=(k - 1) // 2, groups=exp)
        y = _act(_bn(y, sd, f"{b}{j}.1", BODY_BN_EPS), act)
        j += 1
        if use_se:
            s = F.adaptive_avg_pool2d(y, 1)
            s = F.relu(F.conv2d(s, sd[f"{b}{j}.fc1.weight"], sd[f"{b}{j}.fc1.bias"]))
            s = F.hardsigmoid(F.conv2d(s, sd[f"{b}{j}.fc2.weight"], sd[f"{b}{j}.fc2.bias"]))
            y = s * y
            j += 1
        y = _bn(F.conv2d(y, sd[f"{b}{j}.0.weight"]), sd, f"{b}{j}.1", BODY_BN_EPS)
        if stride == 1 and cin == cout:
            y = y + inp
        outs[i] = y
    y = F.conv2d(y, sd[p + "12.0.weight"])
    y = F.hardswish(_bn(y, sd, p + "12.1", BODY_BN_EPS))
    outs[12] = y
    if return_all:
        return outs
    return outs[0], outs[3], outs[8], outs[12]


# ----------------------------------------------------------------------------- FPN (A2)
def fpn_forward(sd, feats, levels=(0, 1, 2, 3)):
    """4 feats -> list of FPN outputs for `levels` (the model only consumes level 0)."""
    lat = [F.conv2d(f, sd[f"backbone.fpn.lateral_convs.{i}.weight"]) for i, f in enumerate(feats)]
    for i in range(3, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = []
    for i in levels:
        y = F.conv2d(lat[i], sd[f"backbone.fpn.fpn_convs.{i}.0.weight"], None, padding=1)
        outs.append(F.relu(_bn(y, sd, f"backbone.fpn.fpn_convs.{i}.1", HEAD_BN_EPS)))
    return outs


# ----------------------------------------------------------------------------- scores + top-k (A3, A4)
def channel_scores(sd, prefix, x):
    """sigmoid(MLP(avgpool) + MLP(maxpool)); MLP = Linear-ReLU-Linear with bias."""
    b, c = x.shape[:2]
    avg = F.adaptive_avg_pool2d(x, 1).view(b, c)
    mx = F.adaptive_max_pool2d(x, 1).view(b, c)

    def mlp(v):
        v = F.relu(F.linear(v, sd[prefix + ".fc.0.weight"], sd[prefix + ".fc.0.bias"]))
        return F.linear(v, sd[prefix + ".fc.2.weight"], sd[prefix + ".fc.2.bias"])
    z = mlp(avg) + mlp(mx)
    if _CALIBRATE is not None:
        z = z * _calib_scale(sd, z, prefix + ".fc.2.weight", prefix + ".fc.2.bias", 1.5, prefix)
    return torch.sigmoid(z)


def select_top_k(scores, k=64):
    """indices sorted by descending score (torch.topk), keypoint_model.py:657."""
    return torch.topk(scores, min(k, scores.shape[1]), dim=1).indices


# ----------------------------------------------------------------------------- ROI align (A5)
def box_center_to_corners(box):
    cx, cy, w, h = [box[..., i] for i in range(4)]
    return torch.stack([(cx - w / 2).clamp(0, 1), (cy - h / 2).clamp(0, 1),
                        (cx + w / 2).clamp(0, 1), (cy + h / 2).clamp(0, 1)], dim=-1)


def roi_align_legacy(feat, rois, out_hw=(56, 56)):
    """feat [C,H,W] (one image), rois [K,4] (x1,y1,x2,y2 in feature pixels) -> [K,C,PH,PW].

    aligned=False, spatial_scale=1, sampling_ratio=-1 (adaptive: ceil(roi/pooled) samples per bin).
    float32 arithmetic in the kernel's order of operations.
    """
    f = feat.detach().cpu().numpy().astype(np.float32)
    C, H, W = f.shape
    PH, PW = out_hw
    r = np.asarray(rois, dtype=np.float32).reshape(-1, 4)
    out = np.zeros((r.shape[0], C, PH, PW), dtype=np.float32)
    f32 = np.float32
    for n in range(r.shape[0]):
        x1, y1, x2, y2 = [f32(v) for v in r[n]]
        rw = max(f32(x2 - x1), f32(1.0))
        rh = max(f32(y2 - y1), f32(1.0))
        bh = f32(rh / f32(PH))
        bw = f32(rw / f32(PW))
        gh = int(math.ceil(float(rh) / PH))
        gw = int(math.ceil(float(rw) / PW))
        count = f32(max(gh * gw, 1))

        def taps(start, bin_sz, grid, pooled, size):
            # -> low, high [pooled*grid], l, h weights (float32), valid mask
            p = np.repeat(np.arange(pooled, dtype=np.float32), grid)
            i = np.tile(np.arange(grid, dtype=np.float32), pooled)
            c = (f32(start) + p * f32(bin_sz) + (i + f32(0.5)) * f32(bin_sz) / f32(grid)).astype(np.float32)
            valid = ~((c < -1.0) | (c > size))
            c = np.maximum(c, f32(0.0))
            lo = c.astype(np.int32)
            edge = lo >= size - 1
            hi = np.where(edge, size - 1, lo + 1)
            lo = np.where(edge, size - 1, lo)
            c = np.where(edge, lo.astype(np.float32), c)
            l = (c - lo.astype(np.float32)).astype(np.float32)
            h = (f32(1.0) - l).astype(np.float32)
            return lo, hi, l, h, valid
        ylo, yhi, ly, hy, yv = taps(y1, bh, gh, PH, H)
        xlo, xhi, lx, hx, xv = taps(x1, bw, gw, PW, W)
        w1 = hy[:, None] * hx[None, :]
        w2 = hy[:, None] * lx[None, :]
        w3 = ly[:, None] * hx[None, :]
        w4 = ly[:, None] * lx[None, :]
        m = (yv[:, None] & xv[None, :]).astype(np.float32)
        v = (w1 * f[:, ylo[:, None], xlo[None, :]] + w2 * f[:, ylo[:, None], xhi[None, :]]
             + w3 * f[:, yhi[:, None], xlo[None, :]] + w4 * f[:, yhi[:, None], xhi[None, :]]) * m
        v = v.reshape(C, PH, gh, PW, gw)
        acc = np.zeros((C, PH, PW), dtype=np.float32)
        for iy in range(gh):          # kernel accumulation order: iy outer, ix inner
            for ix in range(gw):
                acc = acc + v[:, :, iy, :, ix]
        out[n] = acc / count
    return torch.from_numpy(out)


def extract_roi_features(features_b, box, out_hw=(56, 56)):
    """features_b [1,C,H,W], box [4] normalised (cx,cy,w,h) -> [1,C,56,56]  (keypoint_model.py:212-228)."""
    _, C, H, W = features_b.shape
    corners = box_center_to_corners(box) * torch.tensor([W, H, W, H], dtype=torch.float32)
    return roi_align_legacy(features_b[0], corners.view(1, 4).numpy(), out_hw)


# ----------------------------------------------------------------------------- heatmap head (A6)
def heatmap_head_forward(sd, x, prefix="heatmap_head"):
    """x [n,64,56,56] -> (heatmaps [n,17,56,56] in (0,1), (cw [n,64,1,1], sw [n,1,56,56]))."""
    n, c = x.shape[:2]
    cw = channel_scores(sd, prefix + ".channel_attention", x).view(n, c, 1, 1)
    x = x * cw
    sa = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True).values], dim=1)
    z = F.conv2d(sa, sd[prefix + ".spatial_attention.conv.weight"],
                 sd[prefix + ".spatial_attention.conv.bias"], padding=3)
    if _CALIBRATE is not None:
        z = z * _calib_scale(sd, z, prefix + ".spatial_attention.conv.weight",
                             prefix + ".spatial_attention.conv.bias", 1.5, prefix + ".spatial_attention")
    sw = torch.sigmoid(z)
    x = x * sw
    x = F.conv2d(x, sd[prefix + ".deconv_layers.0.weight"], sd[prefix + ".deconv_layers.0.bias"], padding=1)
    x = F.relu(_bn(x, sd, prefix + ".deconv_layers.1", HEAD_BN_EPS))
    x = F.conv2d(x, sd[prefix + ".deconv_layers.4.weight"], sd[prefix + ".deconv_layers.4.bias"], padding=1)
    x = F.relu(_bn(x, sd, prefix + ".deconv_layers.5", HEAD_BN_EPS))
    x = F.conv2d(x, sd[prefix + ".final_layer.0.weight"], sd[prefix + ".final_layer.0.bias"], padding=1)
    x = F.relu(_bn(x, sd, prefix + ".final_layer.1", HEAD_BN_EPS))
    if _CALIBRATE is not None:
        _calib_scale(sd, F.conv2d(x, sd[prefix + ".final_layer.3.weight"]),
                     prefix + ".final_layer.3.weight", None, 2.0, prefix + ".final_layer.3")
    x = F.conv2d(x, sd[prefix + ".final_layer.3.weight"], sd[prefix + ".final_layer.3.bias"])
    return torch.sigmoid(x), (cw, sw)


# ----------------------------------------------------------------------------- decode (A7)
def soft_argmax(heatmap):
    """keypoint_model.py:284-313: softmax over H*W of the (already sigmoided) map, expected (x,y)/(W-1,H-1)."""
    B, K, H, W = heatmap.shape
    sm = torch.softmax(heatmap.reshape(B, K, -1), dim=-1).view(B, K, H, W)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    x = (sm * xs).sum(dim=(2, 3)) / (W - 1)
    y = (sm * ys).sum(dim=(2, 3)) / (H - 1)
    return torch.stack([x, y], dim=-1)


def decode_heatmap(heatmap):
    """keypoint_model.py:250-282 -> (keypoints [B,K,2], visibilities one-hot [B,K,3])."""
    B, K, H, W = heatmap.shape
    kp = soft_argmax(heatmap)
    conf = torch.sigmoid(heatmap.reshape(B, K, -1).max(dim=2).values)
    cls = torch.where(conf < 0.3, 0, torch.where(conf < 0.7, 1, 2))
    vis = F.one_hot(cls, 3).to(torch.float32)
    return kp, vis


def convert_to_original_coords(kp, box):
    """keypoint_model.py:230-248; kp [1,K,2], box [4] (cx,cy,w,h)."""
    cx, cy, w, h = box
    x = (kp[..., 0] * w + (cx - w / 2)).clamp(0, 1)
    y = (kp[..., 1] * h + (cy - h / 2)).clamp(0, 1)
    return torch.stack([x, y], dim=-1)


# ----------------------------------------------------------------------------- person heads (A10, function level)
def person_heads(sd, fpn0, level=0):
    """1x1 box (36ch) and cls (9ch) heads applied to a 128-ch map (person_head.py:27-34)."""
    box = F.conv2d(fpn0, sd[f"person_detector.box_heads.{level}.weight"], sd[f"person_detector.box_heads.{level}.bias"])
    cls = F.conv2d(fpn0, sd[f"person_detector.cls_heads.{level}.weight"], sd[f"person_detector.cls_heads.{level}.bias"])
    return box, cls


# ----------------------------------------------------------------------------- whole forward (A1-A9)
@torch.no_grad()
def forward(sd, images, bboxes, topk_override=None, return_stages=False, num_keypoints=17):
    """Reference-structured forward: per-image, per-box Python loop with batch-1 head calls.

    images [B,Cin,S,S] float32 tensor; bboxes [B,P,4] tensor (normalised cx,cy,w,h; all-zero = padding).
    Returns the reference's output dict ('heatmap','keypoints','visibilities','boxes') (+ 'stages').
    """
    images = torch.as_tensor(images, dtype=torch.float32)
    bboxes = torch.as_tensor(bboxes, dtype=torch.float32)
    B = images.shape[0]
    feats = body_forward(sd, images)
    fpn0 = fpn_forward(sd, feats, levels=(0,))[0]
    scores = channel_scores(sd, "channel_attention", fpn0)
    idx = select_top_k(scores, 64) if topk_override is None else torch.as_tensor(topk_override).long()
    selected = fpn0[torch.arange(B)[:, None], idx]           # [B,64,H,W] in rank order
    boxes_list = [bboxes[i] for i in range(B)]
    stages = dict(feats=feats, fpn0=fpn0, scores=scores, topk=idx, roi=[], cw=[], sw=[], roi_index=[])

    if all(len(b) == 0 for b in boxes_list):
        return {"keypoints": torch.zeros(B, 1, num_keypoints, 2),
                "visibilities": torch.zeros(B, 1, num_keypoints),
                "heatmap": torch.zeros(B, 1, num_keypoints, 56, 56), "boxes": boxes_list}
    max_persons = max(len(b) for b in boxes_list)
    out_k, out_h, out_v = [], [], []
    for b, boxes in enumerate(boxes_list):
        ks, hs, vs = [], [], []
        for p, box in enumerate(boxes):
            if torch.all(box == 0):
                continue
            roi = extract_roi_features(selected[b:b + 1], box)
            hm, (cw, sw) = heatmap_head_forward(sd, roi)
            kp, vis = decode_heatmap(hm)
            kp = convert_to_original_coords(kp, box)
            ks.append(kp); hs.append(hm); vs.append(vis)
            if return_stages:
                stages["roi"].append(roi); stages["cw"].append(cw); stages["sw"].append(sw)
                stages["roi_index"].append((b, p))
        if not ks:
            ks.append(torch.zeros(1, num_keypoints, 2))
            hs.append(torch.zeros(1, num_keypoints, 56, 56))
            v = torch.zeros(1, num_keypoints, 3); v[:, :, 0] = 1.0
            vs.append(v)
        for lst in (ks, hs, vs):
            lst.extend(torch.zeros_like(lst[0]) for _ in range(max_persons - len(lst)))
        out_k.append(torch.stack(ks[:max_persons]))
        out_h.append(torch.stack(hs[:max_persons]))
        out_v.append(torch.stack(vs[:max_persons]))
    out = {"heatmap": torch.stack(out_h).squeeze(2), "keypoints": torch.stack(out_k),
           "visibilities": torch.stack(out_v), "boxes": boxes_list}
    if return_stages:
        out["stages"] = stages
    return out
