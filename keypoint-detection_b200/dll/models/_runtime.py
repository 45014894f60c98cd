"""Host-side plumbing shared by the drop-in modules: plan cache, workspace cache, stage views.

PyTorch is used here only for device memory (torch.empty), streams and tensor views; every compute step is a call
into libkpd_b200.so on the current CUDA stream.
"""
import ctypes as C

import torch

from .. import _native as N
from ._spec import model_entries, init_tensor


def _fingerprint(tensors):
    return tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors)


class PlanCache:
    """Builds the native plan from a (possibly partial) state and rebuilds it when any tensor changes.

    `prefix_map` maps full-model key prefixes to the owning module's local state_dict, e.g. for a standalone
    HeatmapHead: {"heatmap_head.": head.state_dict()}.  Missing entries are filled with their default
    initialisation (they are not touched by the stages that module runs).
    """

    def __init__(self, in_channels=1, num_keypoints=17):
        self.in_channels, self.num_keypoints = in_channels, num_keypoints
        self._plan, self._fp, self._dev = None, None, None
        self._fill = {}
        self._state, self._tensors, self._versions, self._ptrs = None, None, None, None

    def get(self, state, device):
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        # fast path (every forward): same state dict object as last time -> compare every tensor's version counter (in-place
        # edits) and address (`param.data = ...` swaps, sub-module moves; whole-model moves and load_state_dict drop the
        # cached dict in the owning module).  ~60 us instead of ~1 ms for the full fingerprint.
        if state is self._state and self._plan is not None and self._dev == dev_index:
            ts = self._tensors
            if [t._version for t in ts] == self._versions and [t.data_ptr() for t in ts] == self._ptrs:
                return self._plan
        spec = model_entries(self.in_channels, self.num_keypoints)
        tensors = []
        for name, entry in spec.items():
            t = state.get(name)
            if t is None:
                if name not in self._fill:
                    self._fill[name] = init_tensor(entry).zero_() if entry.kind != "ones" else init_tensor(entry)
                t = self._fill[name]
            elif tuple(t.shape) != entry.shape:
                raise ValueError(f"state tensor {name}: expected shape {entry.shape}, got {tuple(t.shape)}")
            tensors.append(t)
        fp = _fingerprint(tensors)
        if self._plan is None or fp != self._fp or self._dev != dev_index:
            if self._plan is not None:
                self._plan.close()
            self._plan = N.Plan(tensors, self.in_channels, self.num_keypoints, dev_index)
            self._fp, self._dev = fp, dev_index
        self._state, self._tensors = state, tensors
        self._versions = [t._version for t in tensors]
        self._ptrs = [t.data_ptr() for t in tensors]
        return self._plan

    def invalidate(self):
        if self._plan is not None:
            self._plan.close()
        self._plan, self._fp = None, None
        self._state = None


class Workspace:
    """Grow-only device scratch buffer (1024-byte aligned view of a torch uint8 tensor)."""

    def __init__(self):
        self._raw, self._view = None, None

    def get(self, nbytes, device):
        if self._view is None or self._view.numel() < nbytes or self._view.device != device:
            self._raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            off = (-self._raw.data_ptr()) % 1024
            self._view = self._raw[off:off + nbytes]
        return self._view


def view(ws, offset, shape, dtype):
    """Typed view of a stage tensor inside the workspace."""
    n = 1
    for s in shape:
        n *= s
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    return ws[offset:offset + nbytes].view(dtype).view(*shape)


def stage_views(ws, L, batch, size, n_rois, precision):
    """Dict of named stage tensors (views, NHWC) for parity tests / debugging."""
    elt = torch.bfloat16 if precision == N.PREC_BF16 else torch.float32
    h = [size // 2, size // 8, size // 16, size // 32]
    c = [16, 24, 48, 576]
    out = {f"feat{i}": view(ws, L.feat[i], (batch, h[i], h[i], c[i]), torch.float32) for i in range(4)}
    out["lateral0"] = view(ws, L.lateral0, (batch, h[0], h[0], 128), torch.float32)
    out["fpn0"] = view(ws, L.fpn0, (batch, h[0], h[0], 128), torch.float32)
    out["pool_avg"] = view(ws, L.pool_avg, (batch, 128), torch.float32)
    out["pool_max"] = view(ws, L.pool_max, (batch, 128), torch.float32)
    out["scores"] = view(ws, L.scores, (batch, 128), torch.float32)
    out["topk"] = view(ws, L.topk, (batch, 64), torch.int32)
    out["roi_count"] = view(ws, L.roi_count, (1,), torch.int32)
    if n_rois > 0:
        out["roi_box"] = view(ws, L.roi_box, (n_rois, 4), torch.float32)
        out["roi_batch"] = view(ws, L.roi_batch, (n_rois,), torch.int32)
        out["roi_slot"] = view(ws, L.roi_slot, (n_rois,), torch.int32)
        out["img_count"] = view(ws, L.img_count, (batch,), torch.int32)
        out["roi_feat"] = view(ws, L.roi_feat, (n_rois, 56, 56, 64), torch.float32)
        out["head_cw"] = view(ws, L.head_cw, (n_rois, 64), torch.float32)
        out["head_sw"] = view(ws, L.head_sw, (n_rois, 56, 56), torch.float32)
        out["head_in"] = view(ws, L.head_in, (n_rois, 56, 56, 64), elt)
        out["head_a"] = view(ws, L.head_a, (n_rois, 56, 56, 256), elt)
        out["head_b"] = view(ws, L.head_b, (n_rois, 56, 56, 256), elt)
        out["head_c"] = view(ws, L.head_c, (n_rois, 56, 56, 64), elt)
    return out


def as_cuda_f32(t, device):
    """Inputs may arrive on the host (the reference's DataLoader path): one H2D copy, no compute."""
    if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
    return t
