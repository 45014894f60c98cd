"""Architecture tables of the reference model, restated for the oracle (test infrastructure).

Sources:
  * dll/models/backbone.py:247-264   MobileNetV3Wrapper = torchvision mobilenet_v3_small
    features[0:13] (return nodes features.0/.3/.8/.12) + LightweightFPN([16,24,48,576], 128)
  * tv:models/mobilenetv3.py:53-116,224-266   InvertedResidual and the "small" config table
  * tv:ops/misc.py:225-261                    SqueezeExcitation (1x1 convs WITH bias)
  * dll/models/keypoint_model.py:49-71        top-level attribute names (state_dict prefixes)
  * dll/models/heatmap_head.py:20-68,115-151  HeatmapHead / ChannelAttention / SpatialAttention
  * dll/models/person_head.py:7-52            PERSON_HEAD heads + anchors buffer
"""
from collections import OrderedDict


def _make_divisible(v, divisor=8):
    # tv:models/_utils.py:_make_divisible
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# (index, cin, kernel, expanded, cout, use_se, activation, stride)   tv:models/mobilenetv3.py:246-258
MBV3_SMALL_BLOCKS = [
    (1, 16, 3, 16, 16, True, "RE", 2),
    (2, 16, 3, 72, 24, False, "RE", 2),
    (3, 24, 3, 88, 24, False, "RE", 1),
    (4, 24, 5, 96, 40, True, "HS", 2),
    (5, 40, 5, 240, 40, True, "HS", 1),
    (6, 40, 5, 240, 40, True, "HS", 1),
    (7, 40, 5, 120, 48, True, "HS", 1),
    (8, 48, 5, 144, 48, True, "HS", 1),
    (9, 48, 5, 288, 96, True, "HS", 2),
    (10, 96, 5, 576, 96, True, "HS", 1),
    (11, 96, 5, 576, 96, True, "HS", 1),
]
BODY_BN_EPS = 1e-3      # tv:models/mobilenetv3.py:155  partial(BatchNorm2d, eps=0.001, momentum=0.01)
HEAD_BN_EPS = 1e-5      # nn.BatchNorm2d default (backbone.py:24, heatmap_head.py:37,62)
FPN_IN_CHANNELS = [16, 24, 48, 576]   # backbone.py:255
FPN_OUT = 128


def se_squeeze(expanded):
    return _make_divisible(expanded // 4, 8)     # tv:models/mobilenetv3.py:79


def _bn(spec, prefix, c):
    spec[prefix + ".weight"] = ((c,), "bn_weight")
    spec[prefix + ".bias"] = ((c,), "bn_bias")
    spec[prefix + ".running_mean"] = ((c,), "bn_mean")
    spec[prefix + ".running_var"] = ((c,), "bn_var")
    spec[prefix + ".num_batches_tracked"] = ((), "nbt")


def state_dict_spec(in_channels=1, num_keypoints=17, anchor_sizes=(32, 64, 128), num_classes=1):
    """Ordered {key: (shape, kind)} of the reference model's 318-entry state_dict."""
    s = OrderedDict()
    p = "backbone.body.features."
    s[p + "0.0.weight"] = ((16, in_channels, 3, 3), "conv")
    _bn(s, p + "0.1", 16)
    for (i, cin, k, exp, cout, use_se, act, stride) in MBV3_SMALL_BLOCKS:
        b = f"{p}{i}.block."
        j = 0
        if exp != cin:
            s[f"{b}{j}.0.weight"] = ((exp, cin, 1, 1), "conv")
            _bn(s, f"{b}{j}.1", exp)
            j += 1
        s[f"{b}{j}.0.weight"] = ((exp, 1, k, k), "conv")
        _bn(s, f"{b}{j}.1", exp)
        j += 1
        if use_se:
            sq = se_squeeze(exp)
            s[f"{b}{j}.fc1.weight"] = ((sq, exp, 1, 1), "conv")
            s[f"{b}{j}.fc1.bias"] = ((sq,), "bias")
            s[f"{b}{j}.fc2.weight"] = ((exp, sq, 1, 1), "conv")
            s[f"{b}{j}.fc2.bias"] = ((exp,), "bias")
            j += 1
        s[f"{b}{j}.0.weight"] = ((cout, exp, 1, 1), "conv")
        _bn(s, f"{b}{j}.1", cout)
    s[p + "12.0.weight"] = ((576, 96, 1, 1), "conv")
    _bn(s, p + "12.1", 576)
    for i, c in enumerate(FPN_IN_CHANNELS):
        s[f"backbone.fpn.lateral_convs.{i}.weight"] = ((FPN_OUT, c, 1, 1), "conv")
    for i in range(4):
        s[f"backbone.fpn.fpn_convs.{i}.0.weight"] = ((FPN_OUT, FPN_OUT, 3, 3), "conv")
        _bn(s, f"backbone.fpn.fpn_convs.{i}.1", FPN_OUT)
    na = len(anchor_sizes) * 3
    s["person_detector.anchors"] = ((56 * 56 * na, 4), "anchors")
    for i in range(4):
        s[f"person_detector.box_heads.{i}.weight"] = ((4 * na, FPN_OUT, 1, 1), "conv")
        s[f"person_detector.box_heads.{i}.bias"] = ((4 * na,), "bias")
    for i in range(4):
        s[f"person_detector.cls_heads.{i}.weight"] = ((num_classes * na, FPN_OUT, 1, 1), "conv")
        s[f"person_detector.cls_heads.{i}.bias"] = ((num_classes * na,), "bias")
    h = "heatmap_head."
    s[h + "channel_attention.fc.0.weight"] = ((4, 64), "linear")
    s[h + "channel_attention.fc.0.bias"] = ((4,), "bias")
    s[h + "channel_attention.fc.2.weight"] = ((64, 4), "linear")
    s[h + "channel_attention.fc.2.bias"] = ((64,), "bias")
    s[h + "spatial_attention.conv.weight"] = ((1, 2, 7, 7), "conv")
    s[h + "spatial_attention.conv.bias"] = ((1,), "bias")
    s[h + "deconv_layers.0.weight"] = ((256, 64, 3, 3), "conv")
    s[h + "deconv_layers.0.bias"] = ((256,), "bias")
    _bn(s, h + "deconv_layers.1", 256)
    s[h + "deconv_layers.4.weight"] = ((256, 256, 3, 3), "conv")
    s[h + "deconv_layers.4.bias"] = ((256,), "bias")
    _bn(s, h + "deconv_layers.5", 256)
    s[h + "final_layer.0.weight"] = ((64, 256, 3, 3), "conv")
    s[h + "final_layer.0.bias"] = ((64,), "bias")
    _bn(s, h + "final_layer.1", 64)
    s[h + "final_layer.3.weight"] = ((num_keypoints, 64, 1, 1), "conv")
    s[h + "final_layer.3.bias"] = ((num_keypoints,), "bias")
    s["channel_attention.fc.0.weight"] = ((8, 128), "linear")
    s["channel_attention.fc.0.bias"] = ((8,), "bias")
    s["channel_attention.fc.2.weight"] = ((128, 8), "linear")
    s["channel_attention.fc.2.bias"] = ((128,), "bias")
    return s


def person_anchors(anchor_sizes=(32, 64, 128), grid=(56, 56), aspect_ratios=(0.5, 1.0, 2.0)):
    """dll/models/person_head.py:39-52 -- fixed 56x56 grid, (cx,cy) normalised, (w,h) in pixels."""
    import numpy as np
    h, w = grid
    out = []
    for i in range(h):
        for j in range(w):
            cx = (j + 0.5) / w
            cy = (i + 0.5) / h
            for size in anchor_sizes:
                for ar in aspect_ratios:
                    out.append([cx, cy, size * ar, size / ar])
    return np.asarray(out, dtype=np.float32)
