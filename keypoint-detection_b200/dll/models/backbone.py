"""Backbone drop-in: MobileNetV3-Small body + LightweightFPN (reference: dll/models/backbone.py:7-39,247-264).

The modules are parameter containers with the reference's state_dict names; `MobileNetV3Wrapper.forward` runs the
fused CUDA trunk (kpd_backbone_fpn0 + kpd_fpn_levels123) and returns the four FPN maps as NCHW views.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import _native as N
from ..configs.model_config import BackboneConfig
from ._runtime import PlanCache, Workspace, as_cuda_f32, view
from ._spec import ParamTree, body_entries, fpn_entries


class LightweightFPN(ParamTree):
    """backbone.py:7-39: four bias-free 1x1 laterals, nearest top-down adds, four conv3x3+BN+ReLU."""

    def __init__(self, in_channels_list, out_channels):
        if not isinstance(in_channels_list, list):
            raise ValueError("in_channels_list must be a list of input channel sizes")
        if list(in_channels_list) != [16, 24, 48, 576] or out_channels != 128:
            raise ValueError("the B200 hot path implements LightweightFPN([16, 24, 48, 576], 128)")
        super().__init__(fpn_entries(out_channels))


class MobileNetV3Wrapper(nn.Module):
    def __init__(self, config: BackboneConfig, out_channels=128, attention_module=None):
        super().__init__()
        if attention_module is not None:
            raise NotImplementedError("attention_module is never used by MultiPersonKeypointModel (backbone.py:257)")
        self.in_channels = config.in_channels
        self.body = ParamTree(body_entries(config.in_channels))      # width_mult is ignored, like upstream
        self.fpn = LightweightFPN([16, 24, 48, 576], out_channels)
        self.attention_module = None
        self._plans = PlanCache(config.in_channels, 17)
        self._workspace = Workspace()

    def forward(self, x):
        """x [B,Cin,S,S] -> list of 4 FPN maps [B,128,S/2..S/32] (NCHW-shaped views of NHWC storage)."""
        if self.training:
            raise RuntimeError("inference-only drop-in: call .eval()")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.NativeLibraryError("move the module to a CUDA device (no CPU fallback)")
        state = {"backbone." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        plan = self._plans.get(state, dev)
        with torch.cuda.device(dev):
            x = as_cuda_f32(x, dev)
            B, S = x.size(0), x.size(2)
            L = plan.layout(B, S, 0, N.PREC_TF32)
            ws = self._workspace.get(L.total_bytes, dev)
            lib, st = N.load(), N.stream_ptr(dev)
            N.check(lib.kpd_backbone_fpn0(plan.handle, N.PREC_TF32, N.ptr(x), B, S, N.ptr(ws), C.byref(L), st), "kpd_backbone_fpn0")
            hs = [S // 2, S // 8, S // 16, S // 32]
            outs = [view(ws, L.fpn0, (B, hs[0], hs[0], 128), torch.float32).clone()]
            extra = [torch.empty(B, h, h, 128, device=dev, dtype=torch.float32) for h in hs[1:]]
            table = (C.c_void_p * 4)(None, *[C.c_void_p(t.data_ptr()) for t in extra])
            N.check(lib.kpd_fpn_levels123(plan.handle, B, S, N.ptr(ws), C.byref(L), table, st), "kpd_fpn_levels123")
        return [t.permute(0, 3, 1, 2) for t in outs + extra]


# ------------------------------------------------------------------------------------------------------------------------
# the hand-written `BACKBONE` (reference backbone.py:41-245).  Exported by the reference package, never
# instantiated by its model.  Same module tree (so state_dicts interchange), forward on the generic fp32 layer kernels.
from . import _layers as L          # noqa: E402

_ACT_NAME = {nn.ReLU: "relu", nn.Hardswish: "hswish", None: None}


class SEModule(nn.Module):
    """backbone.py:41-58: squeeze (global mean) -> Linear (no bias) -> ReLU -> Linear (no bias) -> Hardsigmoid -> scale."""

    def __init__(self, in_channels, reduction=4):
        super().__init__()
        mid = max(1, in_channels // reduction)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(in_channels, mid, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(mid, in_channels, bias=False), nn.Hardsigmoid(inplace=True))

    def factors(self, x):
        """x NHWC -> per-image channel factors [n, c] (the projection conv applies them to its input)."""
        pooled = L.adaptive_avgpool(x, 1, 1).view(x.shape[0], x.shape[3])
        hidden = L.linear(pooled, self.fc[0].weight.detach().float().contiguous(), None, "relu")
        return L.linear(hidden, self.fc[2].weight.detach().float().contiguous(), None, "hsigmoid")

    def forward(self, x):
        xh = L.to_nhwc(x)
        f = self.factors(xh)
        return L.to_nchw_view(_scale_channels(xh, f))


def _scale_channels(xh, f):
    """x[n,h,w,c] * f[n,c] as a depthwise 1x1 conv with unit weights and a per-image input scale."""
    c = xh.shape[3]
    ones = torch.ones(1, 1, 1, c, device=xh.device)
    return L.conv2d(xh, ones, None, 1, groups=c, in_scale=f)


class ConvBNActivation(nn.Module):
    """backbone.py:60-78: bias-free conv (padding (k-1)//2) + BatchNorm + optional activation, under `.layer`."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, groups=1, activation_layer=nn.Hardswish):
        super().__init__()
        in_channels, out_channels = int(in_channels), int(out_channels)
        mods = [nn.Conv2d(in_channels, out_channels, kernel_size, stride, (kernel_size - 1) // 2, groups=groups, bias=False),
                nn.BatchNorm2d(out_channels)]
        if activation_layer:
            mods.append(activation_layer(inplace=True))
        self.layer = nn.Sequential(*mods)
        self._act = _ACT_NAME[activation_layer]
        self._fold = L.FoldCache(self, lambda: L.fold_conv_bn(self.layer[0], self.layer[1]))

    def run(self, xh, in_scale=None, residual=None):
        conv = self.layer[0]
        w, b = self._fold.get()
        return L.conv2d(xh, w, b, conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.groups, self._act, in_scale, residual)

    def forward(self, x):
        return L.to_nchw_view(self.run(L.to_nhwc(x)))


class InvertedResidual(nn.Module):
    """backbone.py:80-106: [expand 1x1] -> depthwise k x k -> [SE] -> project 1x1 (+ input when stride 1 and cin == cout)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, expand_ratio, se_ratio=None, activation_layer=nn.Hardswish):
        super().__init__()
        self.use_res_connect = stride == 1 and in_channels == out_channels
        exp = int(in_channels * expand_ratio)
        mods = []
        if expand_ratio != 1:
            mods.append(ConvBNActivation(in_channels, exp, kernel_size=1, activation_layer=activation_layer))
        mods.append(ConvBNActivation(exp, exp, kernel_size=kernel_size, stride=stride, groups=exp, activation_layer=activation_layer))
        if se_ratio:
            mods.append(SEModule(exp, reduction=exp // max(1, int(in_channels * se_ratio))))
        mods.append(ConvBNActivation(exp, out_channels, kernel_size=1, activation_layer=None))
        self.block = nn.Sequential(*mods)

    def run(self, xh):
        y, scale = xh, None
        for m in list(self.block)[:-1]:
            if isinstance(m, SEModule):
                scale = m.factors(y)                # applied by the projection conv to its input: no extra pass over the map
            else:
                y = m.run(y)
        return self.block[-1].run(y, in_scale=scale, residual=xh if self.use_res_connect else None)

    def forward(self, x):
        return L.to_nchw_view(self.run(L.to_nhwc(x)))


class LimbAttentionModule(nn.Module):
    """backbone.py:108-125: conv3x3 (c -> c/2, bias) + BN + ReLU -> conv1x1 (-> 1, bias) -> sigmoid -> x * map."""

    def __init__(self, in_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, in_channels // 2, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(in_channels // 2)
        self.conv2 = nn.Conv2d(in_channels // 2, 1, kernel_size=1)
        self.sigmoid = nn.Sigmoid()
        self._fold = L.FoldCache(self, lambda: (L.fold_conv_bn(self.conv1, self.bn1), L.fold_conv_bn(self.conv2)))

    def run(self, xh):
        (w1, b1), (w2, b2) = self._fold.get()
        amap = L.conv2d(L.conv2d(xh, w1, b1, 3, 1, 1, act="relu"), w2, b2, 1, act="sigmoid")
        return L.mul_map(xh, amap)

    def forward(self, x):
        return L.to_nchw_view(self.run(L.to_nhwc(x)))


class _FPN(nn.Module):
    """LightweightFPN (backbone.py:7-39) for arbitrary input widths; same parameter names as the fused-path class above."""

    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        if not isinstance(in_channels_list, list):
            raise ValueError("in_channels_list must be a list of input channel sizes")
        self.lateral_convs = nn.ModuleList([nn.Conv2d(c, out_channels, kernel_size=1, bias=False) for c in in_channels_list])
        self.fpn_convs = nn.ModuleList([nn.Sequential(nn.Conv2d(out_channels, out_channels, 3, 1, 1, bias=False),
                                                      nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True)) for _ in in_channels_list])
        self._fold = L.FoldCache(self, lambda: ([L.fold_conv_bn(c) for c in self.lateral_convs],
                                                [L.fold_conv_bn(s[0], s[1]) for s in self.fpn_convs]))

    def run(self, feats):
        if len(feats) != len(self.lateral_convs):
            raise ValueError(f"Expected {len(self.lateral_convs)} features, got {len(feats)}")
        lat_w, out_w = self._fold.get()
        for f, conv in zip(feats, self.lateral_convs):
            if f.shape[3] != conv.in_channels:
                raise RuntimeError(f"lateral conv expects {conv.in_channels} input channels, got a feature map with {f.shape[3]} "
                                   "(the reference fails the same way for width_mult != 1.0)")
        lats = [L.conv2d(f, w, None, 1) for f, (w, _) in zip(feats, lat_w)]
        for i in range(len(lats) - 1, 0, -1):
            L.upsample_add_(lats[i - 1], lats[i])
        return [L.conv2d(t, w, b, 3, 1, 1, act="relu") for t, (w, b) in zip(lats, out_w)]

    def forward(self, features):
        return [L.to_nchw_view(t) for t in self.run([L.to_nhwc(f) for f in features])]


def _make_divisible(v, divisor=8):
    return max(divisor, ((v + divisor // 2) // divisor) * divisor)


class BACKBONE(nn.Module):
    """backbone.py:127-245: MobileNetV3-Small-like trunk (bias-free Linear SE, width multiplier), features after layers
    1, 6, 10, 12 -> FPN([16, 40, 96, 576] -> out_channels) -> limb attention on level 0.  Returns the four FPN maps (NCHW)."""

    # expand ratio, out channels, kernel, stride, SE ratio, activation   (backbone.py:160-173)
    SETTING = [(1, 16, 3, 2, 0.25, nn.ReLU), (4.5, 24, 3, 2, None, nn.ReLU), (3.67, 24, 3, 1, None, nn.ReLU),
               (4, 40, 5, 2, 0.25, nn.Hardswish), (6, 40, 5, 1, 0.25, nn.Hardswish), (6, 40, 5, 1, 0.25, nn.Hardswish),
               (3, 48, 5, 1, 0.25, nn.Hardswish), (3, 48, 5, 1, 0.25, nn.Hardswish), (6, 96, 5, 2, 0.25, nn.Hardswish),
               (6, 96, 5, 1, 0.25, nn.Hardswish), (6, 96, 5, 1, 0.25, nn.Hardswish)]

    def __init__(self, config: BackboneConfig):
        super().__init__()
        self.config = config
        self.fpn = _FPN([16, 40, 96, 576], config.out_channels)
        width = _make_divisible(16 * config.width_mult)
        layers = [ConvBNActivation(config.in_channels, width, kernel_size=3, stride=2, activation_layer=nn.Hardswish)]
        self.selected_layers = [1, 6, 10, 12]
        self.backbone_outputs = []
        for ratio, cout, k, stride, se, act in self.SETTING:
            cout = _make_divisible(cout * config.width_mult)
            layers.append(InvertedResidual(width, cout, k, stride, ratio, se, act))
            width = cout
        self.lastconv_output_channel = _make_divisible(576 * config.width_mult)
        layers.append(ConvBNActivation(width, self.lastconv_output_channel, kernel_size=1, activation_layer=nn.Hardswish))
        self.features = nn.Sequential(*layers)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self._initialize_weights()
        self.limb_attention = LimbAttentionModule(config.out_channels)

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight); nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        if self.training:
            raise RuntimeError("inference-only drop-in: call .eval()")
        xh = L.to_nhwc(x)
        taps = []
        for i, layer in enumerate(self.features):
            xh = layer.run(xh)
            if i in self.selected_layers:
                taps.append(xh)
        self.backbone_outputs = [L.to_nchw_view(t) for t in taps]
        maps = self.fpn.run(taps)
        maps[0] = self.limb_attention.run(maps[0])
        return [L.to_nchw_view(t) for t in maps]
