"""Heat-map head and decoders: drop-in for dll/models/heatmap_head.py (HeatmapHead :20-113, ChannelAttention :115-136,
SpatialAttention :138-151, decode_heatmaps :265-296, decode_heatmaps_subpixel :298-370, decode_heatmaps_soft_argmax
:372-413).  Modules are parameter containers; compute runs in libkpd_b200.so."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import _native as N
from ._runtime import PlanCache, Workspace, as_cuda_f32
from ._spec import ParamTree, channel_attention_entries, heatmap_head_entries


@dataclass
class HeatmapHeadConfig:
    in_channels: int = 64
    hidden_channels: int = 64
    num_keypoints: int = 17
    heatmap_size: Tuple[int, int] = (56, 56)
    dropout_rate: float = 0.1
    use_attention: bool = True
    num_deconv_layers: int = 2
    deconv_kernel_sizes: Tuple[int, ...] = (4, 4)
    deconv_channels: Tuple[int, ...] = (256, 256)


class ChannelAttention(ParamTree):
    def __init__(self, in_channels: int, reduction_ratio: int = 16):
        super().__init__(channel_attention_entries(in_channels, reduction_ratio))


class SpatialAttention(ParamTree):
    def __init__(self, kernel_size: int = 7):
        from ._spec import Entry
        super().__init__({"conv.weight": Entry((1, 2, kernel_size, kernel_size), "conv_fanout"),
                          "conv.bias": Entry((1,), "zeros")})


class HeatmapHead(ParamTree):
    """CBAM attention -> 3 x (conv3x3 + BN + ReLU) -> conv1x1 -> sigmoid.  forward(x) -> (heatmaps, (cw, sw))."""

    def __init__(self, config):
        if (config.in_channels, config.hidden_channels, config.num_keypoints, tuple(config.deconv_channels),
                config.num_deconv_layers, bool(config.use_attention)) != (64, 64, 17, (256, 256), 2, True):
            raise ValueError("the B200 hot path implements the default HeatmapHead (64 -> 256 -> 256 -> 64 -> 17, attention on)")
        super().__init__(heatmap_head_entries(config.num_keypoints, config.in_channels, config.hidden_channels,
                                              tuple(config.deconv_channels)))
        self.config = config
        self.precision = "fp32"
        self._plans = PlanCache(1, config.num_keypoints)
        self._workspace = Workspace()

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, Optional[Tuple[torch.Tensor, torch.Tensor]]]:
        """x [n,64,56,56] (ROI-aligned features) -> heatmaps [n,17,56,56] in (0,1), (channel [n,64,1,1], spatial [n,1,56,56])."""
        if self.training:
            raise RuntimeError("inference-only drop-in: call .eval()")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.NativeLibraryError("move the module to a CUDA device (no CPU fallback)")
        if x.dim() != 4 or tuple(x.shape[1:]) != (64, 56, 56):
            raise ValueError(f"expected [n,64,56,56], got {tuple(x.shape)}")
        state = {"heatmap_head." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        plan = self._plans.get(state, dev)
        prec = N.PRECISIONS[self.precision]
        with torch.cuda.device(dev):
            x = as_cuda_f32(x, dev)
            n = x.size(0)
            L = plan.layout(1, 32, n, prec)
            ws = self._workspace.get(L.total_bytes, dev)
            hm = torch.empty(n, 17, 56, 56, device=dev, dtype=torch.float32)
            cw = torch.empty(n, 64, 1, 1, device=dev, dtype=torch.float32)
            sw = torch.empty(n, 1, 56, 56, device=dev, dtype=torch.float32)
            N.check(N.load().kpd_heatmap_head(plan.handle, prec, N.ptr(x), n, N.ptr(hm), N.ptr(cw), N.ptr(sw),
                                              N.ptr(ws), C.byref(L), N.stream_ptr(dev)), "kpd_heatmap_head")
        return hm, (cw, sw)


# ---------------------------------------------------------------------------------------------- decoders
def _decode(heatmaps, want, temperature=1.0):
    """Run kpd_decode on [B,K,H,W] (fp32 or bf16 CUDA tensor); `want` selects outputs."""
    dev = N.require_cuda(heatmaps.device if heatmaps.is_cuda else None)
    h = heatmaps
    if h.device != dev:
        h = h.to(dev, non_blocking=True)
    if h.dtype not in (torch.float32, torch.bfloat16):
        h = h.float()
    h = h.contiguous()
    B, K, H, W = h.shape
    out = {}
    with torch.cuda.device(dev):
        def buf(name, shape, dtype=torch.float32):
            if name in want:
                out[name] = torch.empty(shape, device=dev, dtype=dtype)
            return N.ptr(out.get(name))
        args = [buf("argmax", (B, K), torch.int32), buf("maxval", (B, K)), buf("kp_plain", (B, K, 2)),
                buf("kp_subpix", (B, K, 2)), buf("score_subpix", (B, K)), buf("kp_soft", (B, K, 2)), buf("vis", (B, K, 3))]
        N.check(N.load().kpd_decode(N.ptr(h), 1 if h.dtype == torch.bfloat16 else 0, B, K, H, W, float(temperature),
                                    *args, N.stream_ptr(dev)), "kpd_decode")
    return out


def _ensure_batch(heatmaps):
    return (heatmaps.unsqueeze(0), True) if heatmaps.dim() == 3 else (heatmaps, False)


def decode_heatmaps(heatmaps: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """argmax decode -> (keypoints (B,K,2) = (x/(W-1), y/(H-1)), scores (B,K)); 3-D input returns un-batched results."""
    h, was_3d = _ensure_batch(heatmaps)
    o = _decode(h, {"kp_plain", "maxval"})
    kp, sc = o["kp_plain"], o["maxval"]
    return (kp.squeeze(0), sc.squeeze(0)) if was_3d else (kp, sc)


def decode_heatmaps_subpixel(heatmaps: torch.Tensor, window_size: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """argmax + 3x3 border-clipped centre of mass.  As upstream (heatmap_head.py:366-368), a 3-D input is NOT squeezed back."""
    if window_size != 3:
        raise NotImplementedError("the CUDA decoder implements the default window_size=3")
    h, _ = _ensure_batch(heatmaps)
    o = _decode(h, {"kp_subpix", "score_subpix"})
    return o["kp_subpix"], o["score_subpix"]


def decode_heatmaps_soft_argmax(heatmaps: torch.Tensor, temperature: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    h, was_3d = _ensure_batch(heatmaps)
    o = _decode(h, {"kp_soft", "maxval"}, temperature)
    kp, sc = o["kp_soft"], o["maxval"]
    return (kp.squeeze(0), sc.squeeze(0)) if was_3d else (kp, sc)


def decode_heatmaps_full(heatmaps: torch.Tensor):
    """Extension: every decoder output in one pass (argmax int32, max, plain/sub-pixel/soft keypoints, visibility)."""
    h, _ = _ensure_batch(heatmaps)
    return _decode(h, {"argmax", "maxval", "kp_plain", "kp_subpix", "score_subpix", "kp_soft", "vis"})


def decode_model_heatmap(heatmap: torch.Tensor):
    """MultiPersonKeypointModel.decode_heatmap (keypoint_model.py:250-282)."""
    o = _decode(heatmap, {"kp_soft", "vis"})
    return o["kp_soft"], o["vis"]
