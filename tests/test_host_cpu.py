"""CPU-side tests (no GPU): C-ABI surface, host logic of the `dll` drop-in, loud failure without CUDA."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

LIB = os.path.join(ROOT, "keypoint-detection_b200", "lib", "libkpd_b200.so")
HEADER = os.path.join(ROOT, "include", "kpd_b200.h")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    if not os.path.exists(LIB):
        import importlib.util
        spec = importlib.util.spec_from_file_location("kpd_build", os.path.join(ROOT, "keypoint-detection_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        mod.build()
    return LIB


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kpd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kpd_b200.h but not exported"
    lib.kpd_abi_version.restype = ctypes.c_int
    assert lib.kpd_abi_version() == 1


def test_ctypes_binding_covers_the_header(built_library):
    from dll import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == _declared_symbols()
    _native.load()
    # struct layout agrees with the C side: ask the library for a layout without touching a GPU
    # (plan pointer is only dereferenced on the forward paths)


def test_library_contains_blackwell_tensor_core_code(built_library):
    """cuobjdump: sm_100a SASS with tcgen05 MMA (UTCHMMA), TMA (UTMALDG) and TMEM loads (LDTM)."""
    out = subprocess.run(["cuobjdump", "-sass", built_library], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout or "SM100" in out.stdout.upper()
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out.stdout, mnemonic


def test_config_loader_matches_reference_defaults():
    from dll.configs.config_loader import load_config
    from dll.configs import ModelConfig, HeatmapHeadConfig, BackboneConfig
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    assert cfg.model.backbone.in_channels == 1 and cfg.model.backbone.input_size == 224
    assert cfg.model.backbone.width_mult == 1.5                      # accepted, ignored (as upstream)
    assert cfg.model.heatmap_head.heatmap_size == (56, 56)
    assert cfg.model.person_head.nms_iou_threshold == 0.3
    assert cfg.model.num_keypoints == 17
    cfg.model.validate()
    assert ModelConfig().heatmap_head.heatmap_size == (224, 224)     # dataclass default differs from loader default
    assert BackboneConfig().in_channels == 3
    with pytest.raises(AssertionError):
        BackboneConfig(input_size=100).validate()
    with pytest.raises(FileNotFoundError):
        load_config("/nonexistent.yaml")
    d = HeatmapHeadConfig().to_dict()
    assert HeatmapHeadConfig.from_dict(d) == HeatmapHeadConfig()


def _model():
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    return MultiPersonKeypointModel(cfg.model, cfg.training)


def test_state_dict_is_checkpoint_compatible():
    from oracle.weights import synth_state_dict
    m = _model()
    lines = open(os.path.join(GOLDEN, "state_dict_keys.txt")).read().strip().splitlines()
    sd = m.state_dict()
    assert len(sd) == 318 == len(lines)
    for (k, v), line in zip(sd.items(), lines):
        name, rest = line.split(" ", 1)
        assert k == name and rest.startswith(str(list(v.shape))) and rest.endswith(str(v.dtype).replace("torch.", ""))
    assert sum(p.numel() for p in m.parameters()) == 2516212
    res = m.load_state_dict(synth_state_dict(0), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ck = {"epoch": 3, "model_state_dict": synth_state_dict(0), "optimizer_state_dict": {}, "history": {}}
    filt = {k: v for k, v in ck["model_state_dict"].items() if k in m.state_dict()}      # scripts/predict.py:48-54
    m.load_state_dict(filt, strict=False)
    assert {n for n, _ in m.named_children()} == {"backbone", "person_detector", "heatmap_head", "channel_attention", "loss_fn"}
    np.testing.assert_allclose(m.person_detector.anchors[:3].numpy(),
                               [[0.5 / 56, 0.5 / 56, 16, 64], [0.5 / 56, 0.5 / 56, 32, 32], [0.5 / 56, 0.5 / 56, 64, 16]], rtol=1e-6)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dll._native import NativeLibraryError
    from dll.models.heatmap_head import decode_heatmaps
    m = _model().eval()
    with pytest.raises(NativeLibraryError):
        m({"image": torch.zeros(1, 1, 224, 224), "bboxes": torch.zeros(1, 1, 4)})
    with pytest.raises(NativeLibraryError):
        decode_heatmaps(torch.zeros(1, 17, 64, 48))
    with pytest.raises(NativeLibraryError):
        m.person_detector.non_max_suppression(torch.zeros(4, 4), torch.zeros(4))


def test_forward_argument_errors_match_reference():
    m = _model().eval()
    with pytest.raises(TypeError):
        m({"image": "nope"})
    m.train()
    with pytest.raises(RuntimeError):
        m({"image": torch.zeros(1, 1, 224, 224), "bboxes": torch.zeros(1, 1, 4)})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "keypoint-detection_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(base, f)
