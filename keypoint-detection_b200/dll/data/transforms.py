"""ITransform: the image preparation `scripts/predict.py` applies before the model (reference: dll/data/transforms.py:9-114).

Caller-side host glue, NOT part of the GPU hot path (SURVEY.md 8f N2 lists a GPU version as later work): it runs once per
image on the CPU with OpenCV, exactly like the reference, and produces the [C,S,S] float tensor the model consumes.

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


class ITransform:
    def __init__(self, img_size=224, clip_limit=1.5, tile_size=(8, 8), grayscale=True):
        self.img_size, self.clip_limit, self.tile_size, self.grayscale = img_size, clip_limit, tuple(tile_size), grayscale

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
