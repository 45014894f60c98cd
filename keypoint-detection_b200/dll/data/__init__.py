"""Host-side input preparation used by scripts/predict.py (reference: dll/data/__init__.py).  Only the image transform of
the inference caller is provided; datasets / loaders / augmentation belong to training and are out of scope."""
from .transforms import ITransform

__all__ = ['ITransform']
