"""Person head drop-in (reference: dll/models/person_head.py:7-166): 1x1 box/cls heads, cxcywh IoU, greedy NMS."""
import ctypes as C

import torch
import torch.nn as nn

from .. import _native as N
from ..configs.model_config import PersonDetectionConfig
from ._runtime import PlanCache, as_cuda_f32
from ._spec import ParamTree, person_head_entries, _anchors


class PERSON_HEAD(ParamTree):
    def __init__(self, config: PersonDetectionConfig):
        if config.in_channels != 128 or config.num_classes != 1 or len(config.anchor_sizes) != 3:
            raise ValueError("the B200 hot path implements PERSON_HEAD(in_channels=128, num_classes=1, 3 anchor sizes)")
        super().__init__(person_head_entries(config.in_channels, config.num_classes, len(config.anchor_sizes)))
        self.num_classes = config.num_classes
        self.conf_threshold = config.conf_threshold
        self.nms_iou_threshold = config.nms_iou_threshold
        self.anchor_sizes = config.anchor_sizes
        self.grid_size = (56, 56)
        self.aspect_ratios = [0.5, 1.0, 2.0]
        with torch.no_grad():
            self.anchors.copy_(_anchors(tuple(config.anchor_sizes)))
        self._plans = PlanCache(1, 17)

    # ------------------------------------------------------------------ IoU / NMS (person_head.py:54-139)
    def box_iou(self, boxes1, boxes2):
        """(N,4),(M,4) cx,cy,w,h -> IoU (N,M)."""
        dev = N.require_cuda(boxes1.device if boxes1.is_cuda else None)
        with torch.cuda.device(dev):
            b1, b2 = as_cuda_f32(boxes1, dev), as_cuda_f32(boxes2, dev)
            out = torch.empty(b1.size(0), b2.size(0), device=dev, dtype=torch.float32)
            N.check(N.load().kpd_box_iou(N.ptr(b1), b1.size(0), N.ptr(b2), b2.size(0), N.ptr(out), N.stream_ptr(dev)), "kpd_box_iou")
        return out

    def non_max_suppression(self, boxes, scores, iou_threshold=0.2, max_output_size=None):
        """Greedy NMS on one image: returns int64 indices of the kept boxes in descending-score order."""
        keep, count = self.batched_nms(boxes.unsqueeze(0), scores.unsqueeze(0), iou_threshold, max_output_size)
        return keep[0, :int(count[0])].to(torch.int64).cpu()

    def batched_nms(self, boxes, scores, iou_threshold=0.2, max_output_size=None):
        """Extension: boxes [B,M,4], scores [B,M] -> (keep int32 [B,M] padded with -1, count int32 [B]) on device."""
        dev = N.require_cuda(boxes.device if boxes.is_cuda else None)
        with torch.cuda.device(dev):
            b, s = as_cuda_f32(boxes, dev), as_cuda_f32(scores, dev)
            B, M = s.shape
            keep = torch.empty(B, M, device=dev, dtype=torch.int32)
            count = torch.empty(B, device=dev, dtype=torch.int32)
            N.check(N.load().kpd_nms(N.ptr(b), N.ptr(s), B, M, float(iou_threshold), int(max_output_size or 0),
                                     N.ptr(keep), N.ptr(count), N.stream_ptr(dev)), "kpd_nms")
        return keep, count

    # ------------------------------------------------------------------ heads (person_head.py:141-166)
    def heads(self, feature, level=0):
        """box (36ch) and cls (9ch) 1x1 heads of `level` on a [B,128,h,w] map -> ([B,36,h,w], [B,9,h,w])."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.NativeLibraryError("move the module to a CUDA device (no CPU fallback)")
        state = {"person_detector." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        plan = self._plans.get(state, dev)
        with torch.cuda.device(dev):
            f = as_cuda_f32(feature, dev)
            B, Cc, h, w = f.shape
            if Cc != 128:
                raise ValueError(f"expected 128 input channels, got {Cc}")
            nhwc = f.permute(0, 2, 3, 1).contiguous()
            box = torch.empty(B, h, w, 36, device=dev, dtype=torch.float32)
            cls = torch.empty(B, h, w, 9, device=dev, dtype=torch.float32)
            N.check(N.load().kpd_person_heads(plan.handle, level, N.ptr(nhwc), B * h * w, N.ptr(box), N.ptr(cls),
                                              N.stream_ptr(dev)), "kpd_person_heads")
        return box.permute(0, 3, 1, 2), cls.permute(0, 3, 1, 2)

    def forward(self, features, targets=None):
        """Eval-mode reference behaviour: raw box-head output of the LAST feature level (person_head.py:158-166)."""
        if self.training and targets is not None and 'bboxes' in targets:
            return features[-1], targets['bboxes']
        level = min(len(features), 4) - 1
        return self.heads(features[level], level)[0]
