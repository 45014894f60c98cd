"""Multi-person keypoint model: drop-in for the reference's dll/models/keypoint_model.py inference path.

Same constructor, attribute tree (=> same 318 state_dict keys), batch-dict contract and output layout as the
reference's MultiPersonKeypointModel (keypoint_model.py:46-210); the whole forward + heat-map decode runs as
hand-written sm_100a kernels behind the C-ABI of libkpd_b200.so (one call per forward, no per-box Python loop,
no host synchronisation).  Training (loss, backward) is outside the hot path and not provided.
"""
import os
from typing import Tuple

import torch
import torch.nn as nn

from .. import _native as N
from ..configs.model_config import ModelConfig
from ..configs.training_config import TrainingConfig
from ._runtime import PlanCache, Workspace, as_cuda_f32, stage_views
from ._spec import ParamTree, channel_attention_entries
from .backbone import MobileNetV3Wrapper
from .heatmap_head import HeatmapHead, decode_model_heatmap
from .person_head import PERSON_HEAD


class ChannelAttention(ParamTree):
    """Parameter container for keypoint_model.py:18-44 (Linear(C, C/16) + ReLU + Linear(C/16, C), both pools)."""

    def __init__(self, in_channels, reduction_ratio=16):
        super().__init__(channel_attention_entries(in_channels, reduction_ratio))
        self.in_channels = in_channels
        self._plans = PlanCache(1, 17)

    def _run(self, x, want_selected):
        """kpd_channel_attention on x [B,128,H,W] -> (scores [B,128], topk int32 [B,64], selected [B,64,H,W] | None)."""
        import ctypes as C  # noqa: F401  (pointer helpers live in _native)
        if self.in_channels != 128:
            raise NotImplementedError("the CUDA channel attention is built for the model's 128 FPN channels")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise N.NativeLibraryError("move the module to a CUDA device (no CPU fallback)")
        if x.dim() != 4 or x.size(1) != 128:
            raise ValueError(f"expected [B,128,H,W], got {tuple(x.shape)}")
        state = {"channel_attention." + k: v for k, v in self.state_dict(keep_vars=True).items()}
        plan = self._plans.get(state, dev)
        with torch.cuda.device(dev):
            x = as_cuda_f32(x, dev)
            B, _, H, W = x.shape
            scores = torch.empty(B, 128, device=dev, dtype=torch.float32)
            topk = torch.empty(B, 64, device=dev, dtype=torch.int32)
            sel = torch.empty(B, 64, H, W, device=dev, dtype=torch.float32) if want_selected else None
            scratch = torch.empty(B * 512, device=dev, dtype=torch.float32)
            N.check(N.load().kpd_channel_attention(plan.handle, N.ptr(x), B, H * W, N.ptr(scores), N.ptr(topk), N.ptr(sel),
                                                   N.ptr(scratch), N.stream_ptr(dev)), "kpd_channel_attention")
        return scores, topk, sel

    def forward(self, x):
        """keypoint_model.py:33-44: x [B,C,H,W] -> channel scores [B,C] = sigmoid(MLP(avgpool) + MLP(maxpool))."""
        return self._run(x, False)[0]


class _InferenceOnlyLoss(nn.Module):
    """Placeholder for `loss_fn` (keypoint_model.py:65-69): parameter-free, so state_dict is unaffected."""

    def forward(self, *a, **k):
        raise NotImplementedError("training loss is outside the B200 inference hot path")


class MultiPersonKeypointModel(nn.Module):
    """Multi-person keypoint detection model (inference): backbone -> channel top-k -> ROI-align -> heat-map head -> decode."""

    def __init__(self, config: ModelConfig, training_config: TrainingConfig = None, precision: str = "fp32"):
        super().__init__()
        self.config = config
        self.training_config = training_config
        self.backbone = MobileNetV3Wrapper(config.backbone)
        self.person_detector = PERSON_HEAD(config.person_head)
        self.heatmap_head = HeatmapHead(config.heatmap_head)
        self.channel_attention = ChannelAttention(in_channels=config.backbone.out_channels, reduction_ratio=16)
        self.loss_fn = _InferenceOnlyLoss()
        self.num_keypoints = config.num_keypoints
        if config.backbone.out_channels != 128 or config.num_keypoints != 17 or config.heatmap_head.in_channels != 64:
            raise ValueError("the B200 hot path is built for out_channels=128, num_keypoints=17, heatmap_head.in_channels=64")
        self.set_precision(precision)
        self._plans = PlanCache(config.backbone.in_channels, config.num_keypoints)
        self._workspace = Workspace()
        self._chunk_streams, self._chunk_workspaces = [], []     # host-input pipelining (forward)
        self._tensors = None
        self._last = None
        self._pipe = None                                        # submit(): copy stream + two device input slots

    # ------------------------------------------------------------------ plumbing
    def set_precision(self, precision: str):
        """'fp32' (tcgen05 tf32, fp32 storage), 'bf16' (bf16 head, tf32 trunk) or 'fp32_strict' (CUDA-core debug)."""
        if precision not in N.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(N.PRECISIONS)}")
        self.precision = precision
        return self

    def _apply(self, fn, *a, **k):
        self._tensors = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._tensors = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _device(self):
        return next(self.parameters()).device

    def _plan(self):
        dev = self._device()
        if dev.type != "cuda":
            raise N.NativeLibraryError("model parameters are on %s: move the model to a CUDA device (no CPU fallback)" % dev)
        if self._tensors is None:
            self._tensors = dict(self.state_dict(keep_vars=True))
        return self._plans.get(self._tensors, dev), dev

    # ------------------------------------------------------------------ forward
    def forward(self, batch, topk_override=None):
        """Same contract as keypoint_model.py:73-210 (inference).  batch: dict with 'image' [B,C,S,S] and 'bboxes'
        ([Tensor[B,P,4]] | [Tensor[P,4]] | Tensor[B,P,4]; normalised cx,cy,w,h; all-zero rows = padding)."""
        x = batch['image'] if isinstance(batch, dict) else batch
        if not isinstance(x, torch.Tensor):
            raise TypeError("Input must be a tensor or a dict with 'image' key containing a tensor")
        if self.training:
            raise RuntimeError("inference-only drop-in: call model.eval() (or model.predict(batch))")
        if x.dim() != 4 or x.size(1) != self.config.backbone.in_channels or x.size(2) != x.size(3) or x.size(2) % 32:
            raise ValueError(f"image must be [B,{self.config.backbone.in_channels},S,S] with S % 32 == 0, got {tuple(x.shape)}")
        plan, dev = self._plan()
        batch_size = x.size(0)

        if isinstance(batch, dict) and 'bboxes' in batch:
            pb = batch['bboxes']
            if isinstance(pb, list) and len(pb) > 0:
                t = pb[0]
                if t.dim() == 3:
                    boxes = t
                elif t.dim() == 2:
                    boxes = t.unsqueeze(0)
                else:
                    raise ValueError(f"Invalid bboxes tensor format: {t.shape}")
            elif isinstance(pb, torch.Tensor):
                if pb.dim() == 3 and pb.size(-1) == 4:
                    boxes = pb
                else:
                    raise ValueError(f"Invalid bboxes format: {pb.shape}")
            else:
                boxes = torch.zeros(batch_size, 0, 4, device=dev)
        else:
            # Upstream routes this case to PERSON_HEAD.forward, which returns raw head outputs and then fails with a channel
            # mismatch (keypoint_model.py:114-119): there is no reference behaviour to reproduce.  NEW behaviour,
            # defined in csrc/kpd_detect.cu: boxes from the level-0 person heads on FPN-0 (anchor decode, top-100, confidence
            # threshold, NMS), padded to `max_detections` slots with all-zero boxes.
            return self._forward_detect(x, plan, dev)
        if boxes.size(-1) != 4:
            raise ValueError(f"Invalid bboxes format: {boxes.shape}")
        person_bboxes = list(boxes.unbind(0))          # per-image box tensors, as the reference returns them (one op, not B slices)
        K = self.num_keypoints
        if boxes.size(0) == 0 or boxes.size(1) == 0:
            # keypoint_model.py:121-135 -- eval-mode empty result (note the reference's differing ranks)
            return {'keypoints': torch.zeros(batch_size, 1, K, 2, device=dev),
                    'visibilities': torch.zeros(batch_size, 1, K, device=dev),
                    'heatmap': torch.zeros(batch_size, 1, K, 56, 56, device=dev),
                    'boxes': person_bboxes if person_bboxes else [torch.zeros(0, 4, device=dev) for _ in range(batch_size)]}
        if boxes.size(0) != batch_size:
            raise ValueError(f"bboxes batch {boxes.size(0)} != image batch {batch_size}")

        with torch.cuda.device(dev):
            P, S = boxes.size(1), x.size(2)
            prec = N.PRECISIONS[self.precision]
            heat = torch.empty(batch_size, P, K, 56, 56, device=dev, dtype=torch.float32)
            kpts = torch.empty(batch_size, P, 1, K, 2, device=dev, dtype=torch.float32)
            vis = torch.empty(batch_size, P, 1, K, 3, device=dev, dtype=torch.float32)
            tk = None
            if topk_override is not None:
                tk = topk_override.to(device=dev, dtype=torch.int32).contiguous()
            # Host-resident input (the DataLoader path): images are independent, so the batch is cut into a few chunks, each
            # copied and processed on its own stream and workspace -- the H2D copy of chunk i+1 overlaps the kernels of chunk i.
            # (Results are bit-identical to the one-shot path: nothing in the forward depends on an image's batch-mates.)
            nchunks = 1
            if x.device.type == "cpu" and tk is None:
                want = int(os.environ.get("KPD_H2D_CHUNKS", "0")) or (2 if batch_size >= 128 else 1)
                nchunks = max(1, min(want, batch_size // 32))
            if nchunks == 1:
                xd = as_cuda_f32(x, dev)
                bx = as_cuda_f32(boxes, dev)
                L = plan.layout(batch_size, S, batch_size * P, prec)
                ws = self._workspace.get(L.total_bytes, dev)
                N.check(N.load().kpd_forward(plan.handle, prec, N.ptr(xd), batch_size, S, N.ptr(bx), P, N.ptr(tk),
                                             N.ptr(heat), N.ptr(kpts), N.ptr(vis), N.ptr(ws), L.total_bytes,
                                             N.stream_ptr(dev)), "kpd_forward")
                self._last = (ws, L, batch_size, S, batch_size * P, prec)
            else:
                cur = torch.cuda.current_stream(dev)
                while len(self._chunk_streams) < nchunks:
                    self._chunk_streams.append(torch.cuda.Stream(device=dev))
                    self._chunk_workspaces.append(Workspace())
                step = (batch_size + nchunks - 1) // nchunks
                for i in range(nchunks):
                    a, b = i * step, min(batch_size, (i + 1) * step)
                    if a >= b:
                        break
                    st = self._chunk_streams[i]
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        xi = as_cuda_f32(x[a:b], dev)
                        bi = as_cuda_f32(boxes[a:b], dev)
                        L = plan.layout(b - a, S, (b - a) * P, prec)
                        ws = self._chunk_workspaces[i].get(L.total_bytes, dev)
                        N.check(N.load().kpd_forward(plan.handle, prec, N.ptr(xi), b - a, S, N.ptr(bi), P, None,
                                                     N.ptr(heat[a:b]), N.ptr(kpts[a:b]), N.ptr(vis[a:b]), N.ptr(ws), L.total_bytes,
                                                     N.stream_ptr(dev)), "kpd_forward")
                        self._last = (ws, L, b - a, S, (b - a) * P, prec)     # stage views of the last chunk only
                for i in range(nchunks):
                    cur.wait_stream(self._chunk_streams[i])
        return {'heatmap': heat, 'keypoints': kpts, 'visibilities': vis, 'boxes': person_bboxes}

    max_detections = 4       # person slots of the detector path (fixed, so that no host synchronisation is needed)
    detect_top_m = 100       # candidates per image entering NMS ("<= 100 candidates per image")

    def _forward_detect(self, x, plan, dev):
        """Bare-tensor forward (NEW behaviour, see kpd_detect.cu): trunk + FPN-0 -> person heads -> kpd_detect_persons ->
        the same ROI / head / decode stages as the box path.  Output dict as forward(); 'boxes' holds the detected boxes
        [P,4] per image (all-zero rows = empty slots)."""
        lib = N.load()
        B, S, K, P = x.size(0), x.size(2), self.num_keypoints, self.max_detections
        prec = N.PRECISIONS[self.precision]
        det = self.person_detector
        with torch.cuda.device(dev):
            xd = as_cuda_f32(x, dev)
            L = plan.layout(B, S, B * P, prec)
            ws = self._workspace.get(L.total_bytes, dev)
            st = N.stream_ptr(dev)
            N.check(lib.kpd_backbone_fpn0(plan.handle, prec, N.ptr(xd), B, S, N.ptr(ws), L, st), "kpd_backbone_fpn0")
            N.check(lib.kpd_channel_topk(plan.handle, B, N.ptr(ws), L, st), "kpd_channel_topk")
            h0 = S // 2
            pixels = B * h0 * h0
            fpn0 = stage_views(ws, L, B, S, 0, prec)["fpn0"]
            box = torch.empty(pixels, 36, device=dev, dtype=torch.float32)
            cls = torch.empty(pixels, 9, device=dev, dtype=torch.float32)
            N.check(lib.kpd_person_heads(plan.handle, 0, N.ptr(fpn0), pixels, N.ptr(box), N.ptr(cls), st), "kpd_person_heads")
            boxes = torch.empty(B, P, 4, device=dev, dtype=torch.float32)
            count = torch.empty(B, device=dev, dtype=torch.int32)
            scratch = torch.empty(int(lib.kpd_detect_scratch_bytes(B, self.detect_top_m)), device=dev, dtype=torch.uint8)
            import ctypes as C
            sizes = (C.c_float * 3)(*[float(v) for v in det.anchor_sizes])
            N.check(lib.kpd_detect_persons(N.ptr(box), N.ptr(cls), B, h0, h0, S, sizes, float(det.conf_threshold),
                                           float(det.nms_iou_threshold), self.detect_top_m, P, N.ptr(boxes), N.ptr(count),
                                           N.ptr(scratch), st), "kpd_detect_persons")
            heat = torch.empty(B, P, K, 56, 56, device=dev, dtype=torch.float32)
            kpts = torch.empty(B, P, 1, K, 2, device=dev, dtype=torch.float32)
            vis = torch.empty(B, P, 1, K, 3, device=dev, dtype=torch.float32)
            N.check(lib.kpd_roi_head(plan.handle, prec, B, S, N.ptr(boxes), P, None, N.ptr(heat), N.ptr(kpts), N.ptr(vis),
                                     N.ptr(ws), L, st), "kpd_roi_head")
            self._last = (ws, L, B, S, B * P, prec)
            self._last_detect = {"box": box.view(B, h0 * h0, 36), "cls": cls.view(B, h0 * h0, 9), "count": count}
        return {'heatmap': heat, 'keypoints': kpts, 'visibilities': vis, 'boxes': list(boxes.unbind(0))}

    @torch.no_grad()
    def predict(self, batch, **kw):
        """eval() + no_grad forward (BASELINE north star: the reference itself has no predict method)."""
        if self.training:          # (Module.eval() walks the whole module tree: only when needed)
            self.eval()
        return self.forward(batch, **kw)

    # ------------------------------------------------------------------ pipelined host-input path
    def submit(self, batch):
        """Asynchronous predict for HOST-resident batches (pinned memory recommended): enqueue the H2D copy of this batch on
        a copy stream and its forward on the current stream, and return a PendingResult immediately.  The copy of batch i+1
        overlaps the kernels of batch i (two input slots), which is what a DataLoader-fed inference loop wants:

            pending = model.submit(next(it))
            for nxt in it:
                upcoming = model.submit(nxt)        # its H2D runs under the current batch's kernels
                out = pending.result()              # dict as returned by predict(); device tensors, forward finished
                pending = upcoming

        Same results as predict() bit for bit (the same kpd_forward call on the same device tensors)."""
        if self.training:
            self.eval()
        x = batch['image'] if isinstance(batch, dict) else batch
        boxes = batch.get('bboxes') if isinstance(batch, dict) else None
        if not isinstance(x, torch.Tensor) or not isinstance(boxes, torch.Tensor) or boxes.dim() != 3 or x.device.type != "cpu":
            # anything but the plain (host image tensor, host [B,P,4] boxes) form: no pipelining, run it now
            return PendingResult(self.predict(batch), None)
        plan, dev = self._plan()
        with torch.cuda.device(dev):
            pipe = self._pipe
            if pipe is None:
                pipe = self._pipe = {"copy": torch.cuda.Stream(device=dev), "slot": 0, "x": [None, None], "b": [None, None],
                                     "free": [None, None]}
            slot = pipe["slot"]; pipe["slot"] = slot ^ 1
            cur = torch.cuda.current_stream(dev)
            xs, bs = pipe["x"][slot], pipe["b"][slot]
            if xs is None or xs.shape != x.shape:
                xs = pipe["x"][slot] = torch.empty(x.shape, dtype=torch.float32, device=dev)
            if bs is None or bs.shape != boxes.shape:
                bs = pipe["b"][slot] = torch.empty(boxes.shape, dtype=torch.float32, device=dev)
            with torch.cuda.stream(pipe["copy"]):
                if pipe["free"][slot] is not None:
                    pipe["copy"].wait_event(pipe["free"][slot])          # the forward that last read this slot has finished
                xs.copy_(x, non_blocking=True)
                bs.copy_(boxes, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(pipe["copy"])
            cur.wait_event(copied)
            out = self.forward({'image': xs, 'bboxes': bs})
            done = torch.cuda.Event()
            done.record(cur)
            pipe["free"][slot] = done
        return PendingResult(out, done)

    # ------------------------------------------------------------------ CUDA-graph replay for fixed shapes
    def capture(self, batch):
        """Capture the forward of a fixed-shape device batch ({'image': [B,C,S,S], 'bboxes': [B,P,4]}, both on the GPU) in a CUDA
        graph and return a `CapturedForward`: `captured(image, bboxes)` copies the new inputs into the graph's static buffers and
        replays the 48 launches with one `cudaGraphLaunch`.  Small batches are bound by the host's launch rate (0.42 ms of enqueue
        per forward against 0.65 ms of kernels at B = 1): replay removes that part (B = 8: 0.85 -> 0.75 ms).  Results are bit-identical to
        predict().  The plan is frozen into the graph: re-capture after changing weights or precision."""
        if self.training:
            self.eval()
        x, boxes = batch['image'], batch['bboxes']
        if not (isinstance(x, torch.Tensor) and isinstance(boxes, torch.Tensor) and x.is_cuda and boxes.is_cuda and boxes.dim() == 3):
            raise ValueError("capture() needs device tensors: {'image': [B,C,S,S], 'bboxes': [B,P,4]}")
        plan, dev = self._plan()
        with torch.cuda.device(dev):
            sx, sb = x.detach().float().contiguous().clone(), boxes.detach().float().contiguous().clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                      # warm-up off the capture: plans, workspaces, kernel attributes
                self.forward({'image': sx, 'bboxes': sb})
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward({'image': sx, 'bboxes': sb})
        return CapturedForward(graph, sx, sb, out)

    def stage_tensors(self):
        """Views (NHWC) of the intermediate tensors of the last forward, for stage-wise parity tests."""
        if self._last is None:
            raise RuntimeError("no forward has run yet")
        return stage_views(*self._last)

    # ------------------------------------------------------------------ reference helper methods
    def decode_heatmap(self, heatmap, threshold=0.1):
        """keypoint_model.py:250-282: soft-argmax keypoints [B,K,2] + one-hot 3-class visibilities [B,K,3]."""
        return decode_model_heatmap(heatmap)

    def convert_to_original_coords(self, keypoints, box):
        """keypoint_model.py:230-248 (host glue; the fused forward does this inside the decode kernel)."""
        shape = keypoints.shape
        k = keypoints.reshape(-1, 2)
        cx, cy, w, h = box
        xs = torch.clamp(k[:, 0] * w + (cx - w / 2), 0, 1)
        ys = torch.clamp(k[:, 1] * h + (cy - h / 2), 0, 1)
        return torch.stack([xs, ys], dim=-1).view(shape)

    def extract_roi_features(self, features: torch.Tensor, box: torch.Tensor, output_size: Tuple[int, int] = (56, 56)):
        """keypoint_model.py:212-228: features [1,C,H,W], box [cx,cy,w,h] normalised -> [1,C,56,56] (legacy roi_align).

        Stand-alone form of the ROI-align kernel that forward() runs batched: channels are processed in slabs of 128
        (the kernel's NHWC tile) with two 64-channel gathers per slab."""
        if tuple(output_size) != (56, 56):
            raise NotImplementedError("the CUDA ROI-align kernel is built for the reference's default 56x56 output")
        dev = N.require_cuda(features.device if features.is_cuda else None)
        with torch.cuda.device(dev):
            f = as_cuda_f32(features, dev)
            _, Cc, H, W = f.shape
            bx = as_cuda_f32(box.reshape(1, 4), dev)
            rb = torch.zeros(1, dtype=torch.int32, device=dev)
            part = torch.empty(28 * 2 * 64, dtype=torch.float32, device=dev)
            outs = []
            for c0 in range(0, Cc, 128):
                slab = torch.zeros(1, H, W, 128, device=dev, dtype=torch.float32)
                cs = min(128, Cc - c0)
                slab[..., :cs] = f[:, c0:c0 + cs].permute(0, 2, 3, 1)
                for h0 in range(0, cs, 64):
                    idx = (torch.arange(64, device=dev, dtype=torch.int32) + h0).clamp(max=127).view(1, 64).contiguous()
                    o = torch.empty(1, 56, 56, 64, device=dev, dtype=torch.float32)
                    N.check(N.load().kpd_roi_align(N.ptr(slab), N.ptr(idx), N.ptr(bx), N.ptr(rb), 1, H, W, N.ptr(o), N.ptr(part),
                                                   N.stream_ptr(dev)), "kpd_roi_align")
                    outs.append(o[..., :min(64, cs - h0)].permute(0, 3, 1, 2))
        return torch.cat(outs, dim=1)


class PendingResult:
    """Handle returned by MultiPersonKeypointModel.submit(): result() waits for the forward and returns its output dict."""

    def __init__(self, out, event):
        self._out, self._event = out, event

    def result(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._out

    def outputs(self):
        """The (possibly still in-flight) output dict: enqueue further work on the same stream without waiting."""
        return self._out


class CapturedForward:
    """A forward frozen into a CUDA graph (`MultiPersonKeypointModel.capture`).  Calling it with new inputs of the captured shapes
    returns the same dict object each time: its tensors are the graph's static output buffers (clone what must outlive the next call)."""

    def __init__(self, graph, image, boxes, out):
        self.graph, self.image, self.boxes, self.out = graph, image, boxes, out

    def __call__(self, image=None, bboxes=None):
        if image is not None:
            self.image.copy_(image, non_blocking=True)
        if bboxes is not None:
            self.boxes.copy_(bboxes, non_blocking=True)
        self.graph.replay()
        return self.out


def select_top_k_channels(features, channel_attention, k=64):
    """keypoint_model.py:653-661: the k best-scoring channels of features [B,C,H,W] in descending score order
    -> [B,k,H,W].  (The fused forward never materialises this tensor: the indices go straight to ROI-align.)"""
    if k != 64 or features.size(1) != 128:
        raise NotImplementedError("the CUDA selection kernel is built for the model's top-64 of 128 channels")
    return channel_attention._run(features, True)[2]


def parse_ground_truth(file_path):
    """keypoint_model.py:587-628: first line of a YOLO-pose label file `class cx cy w h (x y v)*K` ->
    {'bboxes': [Tensor[1,4]], 'keypoints': Tensor[1,1,K,2], 'visibilities': Tensor[1,1,K]}."""
    with open(file_path, 'r') as f:
        data = f.readline().strip().split()
    bbox = torch.tensor([float(v) for v in data[1:5]])
    num_keypoints = (len(data) - 5) // 3
    kp = [[float(data[5 + 3 * i]), float(data[6 + 3 * i])] for i in range(num_keypoints)]
    vis = [int(data[7 + 3 * i]) for i in range(num_keypoints)]
    return {'bboxes': [bbox.unsqueeze(0)],
            'keypoints': torch.tensor(kp).reshape(1, 1, num_keypoints, 2),
            'visibilities': torch.tensor(vis).reshape(1, 1, num_keypoints)}


def box_center_to_corners(box):
    """keypoint_model.py:630-638."""
    cx, cy, w, h = box.unbind()
    return torch.stack([torch.clamp(cx - w / 2, 0, 1), torch.clamp(cy - h / 2, 0, 1),
                        torch.clamp(cx + w / 2, 0, 1), torch.clamp(cy + h / 2, 0, 1)])


def pad_to_length(tensor_list, length):
    """keypoint_model.py:640-651."""
    if len(tensor_list) == 0:
        return []
    if len(tensor_list) >= length:
        return tensor_list[:length]
    return tensor_list + [torch.zeros_like(tensor_list[0]) for _ in range(length - len(tensor_list))]
