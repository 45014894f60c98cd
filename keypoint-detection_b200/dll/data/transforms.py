"""ITransform: the image preparation `scripts/predict.py` applies before the model (reference: dll/data/transforms.py:9-114).

Two executions of the same pipeline:
  * `ITransform(...)` -- the reference's own host path (OpenCV + Pillow, one image at a time), kept for callers without a GPU
    tensor at hand;
  * `ITransform(..., device="cuda")` -- SURVEY.md 8f N2: the whole pipeline as batched CUDA kernels (`kpd_itransform`,
    csrc/kpd_itransform.cu), bit-exact with the libraries for every byte stage, output already on the device in the
    [B,C,S,S] float layout the model consumes.  `transform.batch(u8_tensor)` takes [B,H,W,C] uint8 (host or device).

  grayscale (default): gray -> CLAHE -> {Gaussian 5x5 sigma 1.5 -> median 5 -> Canny(100,200) -> close/open-like
                       dilate,erode,dilate,erode with a 3x3 box -> Gaussian 3x3 -> rescale to 0..255} -> 0.7*CLAHE + 0.3*edges
                       -> resize (PIL bilinear, antialiased, as torchvision's Resize on a PIL image) -> /255 -> (x-0.5)/0.5
  rgb:                 per-channel CLAHE -> Gaussian 3x3 sigma 0.5 -> resize -> /255 -> ImageNet mean/std
"""
import numpy as np
import torch
from PIL import Image

_IMAGENET_MEAN = (0.485, 0.456, 0.406)
_IMAGENET_STD = (0.229, 0.224, 0.225)


def _cv2():
    try:
        import cv2
    except ImportError as e:          # pragma: no cover - cv2 is a hard dependency of the reference's transform too
        raise ImportError("ITransform needs OpenCV (cv2), like the reference's dll/data/transforms.py") from e
    return cv2


class _Prefetched:
    """Ticket of `ITransform.prefetch`: the staged device view of a pinned host batch and the events of its chunked copy."""

    def __init__(self, device_view, landed, per, reuse, host):
        self.device_view, self.landed, self.per, self.reuse, self.host = device_view, landed, per, reuse, host


class ITransform:
    def __init__(self, img_size=224, clip_limit=1.5, tile_size=(8, 8), grayscale=True, device=None):
        self.img_size, self.clip_limit, self.tile_size, self.grayscale = img_size, clip_limit, tuple(tile_size), grayscale
        self.device = None if device is None else torch.device(device)
        if self.device is not None and self.device.type != "cuda":
            raise ValueError(f"ITransform(device=...) selects the CUDA pipeline; got {self.device} (omit it for the host path)")
        self._scratch, self._staging, self.h2d_chunks = {}, {}, 4

    # ---- CUDA pipeline (kpd_itransform)
    def _check_images(self, images):
        x = torch.as_tensor(images)
        if x.dtype != torch.uint8:
            raise TypeError(f"ITransform.batch expects uint8 images, got {x.dtype}")
        if x.dim() == 3:
            x = x.unsqueeze(-1)
        if x.dim() != 4 or x.shape[-1] not in (1, 3, 4):
            raise ValueError(f"ITransform.batch expects [B,H,W] or [B,H,W,C] with C in 1, 3, 4; got {tuple(x.shape)}")
        return x

    def prefetch(self, images):
        """Start the host-to-device copy of a PINNED uint8 batch on the side stream (in `h2d_chunks` pieces) and return a ticket
        for `batch()`.  Two staging slots per device: the copy of batch i + 1 can cross PCIe while batch i is still being
        transformed and consumed, which is how an inference loop over camera frames hides the 0.9 MB per frame."""
        from .. import _native
        if self.device is None:
            raise RuntimeError("ITransform.prefetch needs ITransform(..., device='cuda')")
        dev = _native.require_cuda(self.device)
        x = self._check_images(images)
        if x.is_cuda or not x.is_pinned() or not x.is_contiguous():
            raise ValueError("ITransform.prefetch expects a contiguous pinned host tensor")
        B = x.shape[0]
        chunks = max(1, min(self.h2d_chunks, B // 2))
        per = (B + chunks - 1) // chunks
        with torch.cuda.device(dev):
            st = self._staging.setdefault(dev, {"side": torch.cuda.Stream(dev), "slots": [None, None], "next": 0})
            k = st["next"]; st["next"] = k ^ 1
            slot = st["slots"][k]
            if slot is None or slot[0].numel() < x.numel():
                slot = st["slots"][k] = (torch.empty(x.numel(), dtype=torch.uint8, device=dev), torch.cuda.Event())
            buf, reuse = slot
            side = st["side"]
            xd = buf[:x.numel()].view(x.shape)
            side.wait_event(reuse)                                  # the kernels of the batch that last used this slot are done
            landed = []
            with torch.cuda.stream(side):
                for lo in range(0, B, per):
                    xd[lo:lo + per].copy_(x[lo:lo + per], non_blocking=True)
                    e = torch.cuda.Event(); e.record(side); landed.append(e)
        return _Prefetched(xd, landed, per, reuse, x)

    def batch(self, images, return_stages=False):
        """[B,H,W] / [B,H,W,C] uint8 tensor or array (C = 1, 3 or 4; host or device), or a `prefetch()` ticket -> float32
        [B,C',S,S] on the device.  A pinned host batch is prefetched in chunks so that chunk i + 1 crosses PCIe under chunk
        i's kernels.  With return_stages=True also a dict of the intermediate uint8 planes (gray, clahe, edges, soft, hpass)."""
        import ctypes as C
        from .. import _native
        if self.device is None:
            raise RuntimeError("ITransform.batch needs ITransform(..., device='cuda')")
        dev = _native.require_cuda(self.device)
        ticket = images if isinstance(images, _Prefetched) else None
        if ticket is None:
            x = self._check_images(images)
            if not x.is_cuda and x.is_pinned() and x.is_contiguous() and x.shape[0] >= 2 * self.h2d_chunks and not return_stages:
                ticket = self.prefetch(x)
        if ticket is not None:
            x = ticket.device_view
        B, H, W, Cn = x.shape
        S, gray = int(self.img_size), 1 if self.grayscale else 0
        tx, ty = int(self.tile_size[0]), int(self.tile_size[1])
        lib = _native.load()
        per = ticket.per if ticket is not None else B
        geom = (min(per, B), H, W, gray, tx, ty, S)
        with torch.cuda.device(dev):
            need = lib.kpd_itransform_scratch_bytes(*geom)
            if need == 0:
                raise ValueError(f"ITransform.batch: bad geometry {geom}")
            ws = self._scratch.get(dev)
            if ws is None or ws.numel() < need:
                ws = self._scratch[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty((B, 1 if gray else 3, S, S), dtype=torch.float32, device=dev)
            cur = torch.cuda.current_stream(dev)
            if ticket is not None:
                for i, lo in enumerate(range(0, B, per)):
                    n = min(per, B - lo)
                    cur.wait_event(ticket.landed[i])
                    _native.check(lib.kpd_itransform(_native.ptr(x[lo:lo + n]), n, H, W, Cn, gray, float(self.clip_limit), tx, ty, S,
                                                     _native.ptr(out[lo:lo + n]), _native.ptr(ws), _native.stream_ptr(dev)), "kpd_itransform")
                ticket.reuse.record(cur)
                return out
            x = x.to(dev, non_blocking=True).contiguous()
            _native.check(lib.kpd_itransform(_native.ptr(x), B, H, W, Cn, gray, float(self.clip_limit), tx, ty, S,
                                             _native.ptr(out), _native.ptr(ws), _native.stream_ptr(dev)), "kpd_itransform")
            if not return_stages:
                return out
            off = (C.c_size_t * 6)()
            _native.check(lib.kpd_itransform_stage_offsets(*geom, off), "kpd_itransform_stage_offsets")
            P = 1 if gray else 3
            def plane(i, shape):
                n = 1
                for d in shape:
                    n *= d
                return ws[off[i]:off[i] + n].view(shape).clone()
            stages = {"plane": plane(0, (B, P, H, W)), "lut": plane(1, (B, P, ty, tx, 256)), "clahe": plane(2, (B, P, H, W)),
                      "hpass": plane(5, (B, P, H, S))}
            if gray:
                stages["edges"] = (plane(3, (B, P, H, W)) == 2).to(torch.uint8) * 255
                stages["soft"] = plane(4, (B, P, H, W))
            return out, stages

    # ---- stage functions (same names as the reference's methods: callers use them directly for visualisation)
    def to_grayscale_clahe(self, img):
        cv2 = _cv2()
        a = np.array(img)
        gray = a if a.ndim == 2 else cv2.cvtColor(a, cv2.COLOR_RGB2GRAY)
        eq = cv2.createCLAHE(clipLimit=self.clip_limit, tileGridSize=self.tile_size).apply(gray)
        smooth = cv2.medianBlur(cv2.GaussianBlur(eq, (5, 5), 1.5), 5)
        edges = cv2.Canny(smooth, 100, 200)
        box = np.ones((3, 3), np.uint8)
        for op in (cv2.dilate, cv2.erode, cv2.dilate, cv2.erode):
            edges = op(edges, box, iterations=1)
        edges = cv2.GaussianBlur(edges, (3, 3), 0)
        peak = edges.max()
        if peak > 0:
            edges = (edges.astype(float) / peak * 255).astype(np.uint8)
        mixed = cv2.addWeighted(eq, 0.7, edges, 0.3, 0)
        return Image.fromarray(mixed.astype(np.uint8))

    def to_rgb_clahe(self, img):
        cv2 = _cv2()
        a = np.array(img)
        if a.ndim == 2:
            a = cv2.cvtColor(a, cv2.COLOR_GRAY2RGB)
        elif a.shape[2] == 4:
            a = cv2.cvtColor(a, cv2.COLOR_RGBA2RGB)
        clahe = cv2.createCLAHE(clipLimit=self.clip_limit, tileGridSize=self.tile_size)
        out = cv2.merge([clahe.apply(c) for c in cv2.split(a)])
        return Image.fromarray(cv2.GaussianBlur(out, (3, 3), 0.5).astype(np.uint8))

    def __call__(self, img):
        """PIL image (or HxW / HxWx3 uint8 array) -> float32 tensor [1,S,S] (grayscale) or [3,S,S] (rgb)."""
        if self.device is not None:
            a = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.array(img))
            return self.batch(a.unsqueeze(0))[0]
        if not isinstance(img, Image.Image):
            img = Image.fromarray(np.asarray(img))
        staged = self.to_grayscale_clahe(img) if self.grayscale else self.to_rgb_clahe(img)
        staged = staged.resize((self.img_size, self.img_size), Image.BILINEAR)
        x = torch.from_numpy(np.asarray(staged, dtype=np.uint8).copy()).float().div_(255.0)
        if self.grayscale:
            return ((x - 0.5) / 0.5).unsqueeze(0)
        x = x.permute(2, 0, 1)
        mean = torch.tensor(_IMAGENET_MEAN).view(3, 1, 1); std = torch.tensor(_IMAGENET_STD).view(3, 1, 1)
        return (x - mean) / std
