"""Regenerates tests/golden/itransform.npz from the REFERENCE's own `ITransform` (dll/data/transforms.py:9-114), imported from
/root/reference (build container only; the fixture travels, the reference does not):

    python tests/golden/make_itransform_golden.py

Inputs are `oracle.itransform.synth_photo` images (regenerated from their seeds by the tests, not stored).  Stored per case:
the reference's `to_grayscale_clahe` / `to_rgb_clahe` uint8 result and the final float32 tensor of `ITransform.__call__`.
Library versions are recorded in the file (cv2 / Pillow / torchvision are un-pinned dependencies of the reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = [  # (name, h, w, seed, channels, grayscale, img_size, clip_limit, tiles)
    ("g_240x320", 240, 320, 1, 3, True, 224, 1.5, (8, 8)),
    ("g_227x301", 227, 301, 2, 3, True, 224, 2.0, (8, 8)),        # neither side divides the tile grid (padded CLAHE)
    ("g_200x333", 200, 333, 3, 3, True, 224, 2.0, (8, 8)),        # only the height divides (full extra tile row, clahe.cpp)
    ("g_gray_160x160", 160, 160, 4, 1, True, 96, 1.5, (5, 5)),    # single-channel input, upscaling-free square, 5x5 tiles
    ("g_rgba_128x192", 128, 192, 5, 4, True, 224, 1.5, (8, 8)),   # RGBA input, upsampling resize
    ("c_240x320", 240, 320, 6, 3, False, 224, 1.5, (8, 8)),       # rgb pipeline
    ("c_131x97", 131, 97, 7, 3, False, 64, 2.0, (4, 4)),
]


def main():
    sys.path.insert(0, ROOT)                       # for `oracle`; only when run as a script: the tests import this file for its tables
    sys.path.insert(0, "/root/reference")          # the reference package (build container only)
    import cv2
    import PIL
    import torchvision
    from PIL import Image
    from dll.data.transforms import ITransform as RefITransform          # the reference's class
    from oracle.itransform import synth_photo
    out = {"versions": np.array([f"cv2 {cv2.__version__}", f"PIL {PIL.__version__}", f"torchvision {torchvision.__version__}"])}
    for name, h, w, seed, ch, gray, size, clip, tiles in CASES:
        img = synth_photo(h, w, seed, ch)
        pil = Image.fromarray(img if ch != 1 else img)
        t = RefITransform(img_size=size, clip_limit=clip, tile_size=tiles, grayscale=gray)
        staged = np.asarray(t.to_grayscale_clahe(pil) if gray else t.to_rgb_clahe(pil))
        final = t(pil).numpy()
        out[name + ".staged"] = staged
        out[name + ".final"] = final.astype(np.float32)
        print(name, staged.shape, final.shape, float(final.mean()))
    np.savez_compressed(os.path.join(HERE, "itransform.npz"), **out)


if __name__ == "__main__":
    main()
