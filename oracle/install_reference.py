#!/usr/bin/env python
"""Make the UNMODIFIED reference importable on the GPU box: copy /root/reference/{dll,configs} -> baseline/_ref/.

    python oracle/install_reference.py        (build container only: /root/reference does not exist on the GPU box)

`baseline/_ref/` is git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so
the copy travels with the snapshot and `bench.py --impl reference` / the cpu_baseline leg can time the reference's own
`MultiPersonKeypointModel.forward` on the box's host cores (cpu_baseline.kind = "reference").  The reference is an
un-packaged pure-Python tree (its setup.py lists un-pinned torch / torchvision / opencv; `pip install --target` would only
copy the same files), so a file copy IS the install.  Nothing is patched on disk: the one patch the reference needs
offline -- torchvision's `mobilenet_v3_small(weights=None)` instead of the ImageNet download -- is applied in memory by
`load_reference_model()` below, as in tests/golden/make_golden.py.

Test / bench infrastructure: the product package never imports this module.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = os.environ.get("KPD_REFERENCE", "/root/reference")
REF_DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    """Copy the reference package; returns the install dir or None when /root/reference is absent (GPU box)."""
    if not os.path.isdir(os.path.join(REF_SRC, "dll")):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "dll")) else None
    if force and os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST, exist_ok=True)
    for sub in ("dll", "configs"):
        dst = os.path.join(REF_DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF_SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(REF_DST, "INSTALLED_FROM.txt"), "w") as f:
        f.write(f"file copy of {REF_SRC}/dll and /configs by oracle/install_reference.py (unmodified)\n")
    return REF_DST


def available():
    return os.path.isdir(os.path.join(REF_DST, "dll", "models"))


def load_reference_model(state_dict=None):
    """Import `dll` from baseline/_ref in THIS process (which must not have imported the product's `dll`), build the
    reference MultiPersonKeypointModel on the CPU with the in-memory weights=None patch and optionally load weights."""
    if not available():
        raise FileNotFoundError("baseline/_ref/dll missing: run oracle/install_reference.py in the build container")
    loaded = sys.modules.get("dll")
    if loaded is not None and not os.path.realpath(loaded.__file__).startswith(os.path.realpath(REF_DST)):
        raise RuntimeError("another `dll` package is already imported in this process")
    sys.path.insert(0, REF_DST)
    import torch
    import dll.models.backbone as bb
    from torchvision.models import mobilenet_v3_small as tv_small
    bb.mobilenet_v3_small = lambda weights=None, **kw: tv_small(weights=None, **kw)
    from dll.configs.config_loader import load_config
    from dll.models.keypoint_model import MultiPersonKeypointModel
    cfg = load_config(os.path.join(REF_DST, "configs", "default_config.yaml"))
    torch.manual_seed(0)
    model = MultiPersonKeypointModel(cfg.model, cfg.training)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    return model.eval()


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
