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
