"""Parameter-tree description of the hot-path model: the 318 state_dict entries, in the reference's order.

The reference gets this tree from torchvision's mobilenet_v3_small (backbone.py:250-254) and its own nn.Modules
(keypoint_model.py:55-63).  Here it is generated from a compact table so that the drop-in modules are plain
parameter containers whose state_dict()/load_state_dict() are checkpoint-compatible, while all compute goes through
the native library.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

# name: (cin, kernel, expanded, cout, squeeze-excite, activation, stride)   tv:models/mobilenetv3.py:246-258
_BLOCKS = OrderedDict([
    (1, (16, 3, 16, 16, True, 'RE', 2)), (2, (16, 3, 72, 24, False, 'RE', 2)), (3, (24, 3, 88, 24, False, 'RE', 1)),
    (4, (24, 5, 96, 40, True, 'HS', 2)), (5, (40, 5, 240, 40, True, 'HS', 1)), (6, (40, 5, 240, 40, True, 'HS', 1)),
    (7, (40, 5, 120, 48, True, 'HS', 1)), (8, (48, 5, 144, 48, True, 'HS', 1)), (9, (48, 5, 288, 96, True, 'HS', 2)),
    (10, (96, 5, 576, 96, True, 'HS', 1)), (11, (96, 5, 576, 96, True, 'HS', 1)),
])
FPN_LEVEL_CHANNELS = (16, 24, 48, 576)
FPN_CHANNELS = 128


def _round_channels(v, divisor=8):
    out = max(divisor, int(v + divisor / 2) // divisor * divisor)
    return out + divisor if out < 0.9 * v else out


class Entry:
    __slots__ = ("shape", "kind", "buffer", "dtype")

    def __init__(self, shape, kind, buffer=False, dtype=torch.float32):
        self.shape, self.kind, self.buffer, self.dtype = tuple(shape), kind, buffer, dtype


def _conv(d, name, cout, cin, k, bias=False, kind="conv_fanout"):
    d[name + ".weight"] = Entry((cout, cin, k, k), kind)
    if bias:
        d[name + ".bias"] = Entry((cout,), "zeros" if kind == "conv_fanout" else "bias_uniform:%d" % (cin * k * k))


def _norm(d, name, ch):
    d[name + ".weight"] = Entry((ch,), "ones")
    d[name + ".bias"] = Entry((ch,), "zeros")
    d[name + ".running_mean"] = Entry((ch,), "zeros", buffer=True)
    d[name + ".running_var"] = Entry((ch,), "ones", buffer=True)
    d[name + ".num_batches_tracked"] = Entry((), "zeros", buffer=True, dtype=torch.int64)


def _linear(d, name, cout, cin):
    d[name + ".weight"] = Entry((cout, cin), "linear_default")
    d[name + ".bias"] = Entry((cout,), "bias_uniform:%d" % cin)


def body_entries(in_channels):
    d = OrderedDict()
    f = "features."
    _conv(d, f + "0.0", 16, in_channels, 3); _norm(d, f + "0.1", 16)
    for i, (cin, k, exp, cout, se, _act, _stride) in _BLOCKS.items():
        pos = 0
        if exp != cin:
            _conv(d, f"{f}{i}.block.{pos}.0", exp, cin, 1); _norm(d, f"{f}{i}.block.{pos}.1", exp); pos += 1
        _conv(d, f"{f}{i}.block.{pos}.0", exp, 1, k); _norm(d, f"{f}{i}.block.{pos}.1", exp); pos += 1
        if se:
            sq = _round_channels(exp // 4)
            _conv(d, f"{f}{i}.block.{pos}.fc1", sq, exp, 1, bias=True)
            _conv(d, f"{f}{i}.block.{pos}.fc2", exp, sq, 1, bias=True)
            pos += 1
        _conv(d, f"{f}{i}.block.{pos}.0", cout, exp, 1); _norm(d, f"{f}{i}.block.{pos}.1", cout)
    _conv(d, f + "12.0", 576, 96, 1); _norm(d, f + "12.1", 576)
    return d


def fpn_entries(out_channels=FPN_CHANNELS):
    d = OrderedDict()
    for i, c in enumerate(FPN_LEVEL_CHANNELS):
        _conv(d, f"lateral_convs.{i}", out_channels, c, 1, kind="conv_default")
    for i in range(len(FPN_LEVEL_CHANNELS)):
        _conv(d, f"fpn_convs.{i}.0", out_channels, out_channels, 3, kind="conv_default"); _norm(d, f"fpn_convs.{i}.1", out_channels)
    return d


def person_head_entries(in_channels=128, num_classes=1, num_anchor_sizes=3):
    d = OrderedDict()
    na = num_anchor_sizes * 3
    d["anchors"] = Entry((56 * 56 * na, 4), "anchors", buffer=True)
    for i in range(4):
        _conv(d, f"box_heads.{i}", 4 * na, in_channels, 1, bias=True, kind="conv_default")
    for i in range(4):
        _conv(d, f"cls_heads.{i}", num_classes * na, in_channels, 1, bias=True, kind="conv_default")
    return d


def channel_attention_entries(channels, reduction=16):
    d = OrderedDict()
    r = max(1, channels // reduction)
    _linear(d, "fc.0", r, channels); _linear(d, "fc.2", channels, r)
    return d


def heatmap_head_entries(num_keypoints=17, in_channels=64, hidden=64, deconv=(256, 256)):
    d = OrderedDict()
    for k, v in channel_attention_entries(in_channels).items():
        d["channel_attention." + k] = v
    _conv(d, "spatial_attention.conv", 1, 2, 7, bias=True)
    prev = in_channels
    for i, ch in enumerate(deconv):
        _conv(d, f"deconv_layers.{4 * i}", ch, prev, 3, bias=True); _norm(d, f"deconv_layers.{4 * i + 1}", ch)
        prev = ch
    _conv(d, "final_layer.0", hidden, prev, 3, bias=True); _norm(d, "final_layer.1", hidden)
    _conv(d, "final_layer.3", num_keypoints, hidden, 1, bias=True)
    return d


def model_entries(in_channels=1, num_keypoints=17):
    """Full 318-entry spec with the top-level prefixes of MultiPersonKeypointModel."""
    d = OrderedDict()
    for k, v in body_entries(in_channels).items():
        d["backbone.body." + k] = v
    for k, v in fpn_entries().items():
        d["backbone.fpn." + k] = v
    for k, v in person_head_entries().items():
        d["person_detector." + k] = v
    for k, v in heatmap_head_entries(num_keypoints).items():
        d["heatmap_head." + k] = v
    for k, v in channel_attention_entries(FPN_CHANNELS).items():
        d["channel_attention." + k] = v
    return d


def _anchors(sizes=(32, 64, 128), ratios=(0.5, 1.0, 2.0), grid=56):
    # person_head.py:39-52: (cx, cy) normalised cell centres, (w, h) = size*ratio, size/ratio in pixels
    ys, xs = torch.meshgrid(torch.arange(grid, dtype=torch.float64), torch.arange(grid, dtype=torch.float64), indexing="ij")
    centres = torch.stack([(xs + 0.5) / grid, (ys + 0.5) / grid], dim=-1).reshape(-1, 1, 2)
    wh = torch.tensor([[s * r, s / r] for s in sizes for r in ratios], dtype=torch.float64).reshape(1, -1, 2)
    return torch.cat([centres.expand(-1, wh.shape[1], -1), wh.expand(centres.shape[0], -1, -1)], dim=-1).reshape(-1, 4).float()


def init_tensor(entry, anchor_sizes=(32, 64, 128)):
    t = torch.empty(entry.shape, dtype=entry.dtype)
    kind = entry.kind
    if kind == "zeros":
        t.zero_()
    elif kind == "ones":
        t.fill_(1)
    elif kind == "conv_fanout":          # torchvision / HeatmapHead.init_weights: kaiming_normal_(mode='fan_out')
        nn.init.kaiming_normal_(t, mode="fan_out", nonlinearity="relu")
    elif kind in ("conv_default", "linear_default"):
        nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    elif kind.startswith("bias_uniform:"):
        bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
        nn.init.uniform_(t, -bound, bound)
    elif kind == "anchors":
        t.copy_(_anchors(anchor_sizes))
    else:
        raise KeyError(kind)
    return t


class ParamTree(nn.Module):
    """Generic parameter container: builds nested sub-modules from dotted names (state_dict-compatible)."""

    def __init__(self, entries=None):
        super().__init__()
        for name, e in (entries or {}).items():
            self._insert(name.split("."), e)

    def _insert(self, parts, entry):
        if len(parts) == 1:
            t = init_tensor(entry)
            if entry.buffer:
                self.register_buffer(parts[0], t)
            else:
                self.register_parameter(parts[0], nn.Parameter(t, requires_grad=t.is_floating_point()))
            return
        child = self._modules.get(parts[0])
        if child is None:
            child = ParamTree()
            self.add_module(parts[0], child)
        child._insert(parts[1:], entry)

    def forward(self, *a, **k):
        raise NotImplementedError("parameter container: compute runs in the fused CUDA kernels of libkpd_b200")
