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


def test_fpn0_composed_form_is_the_same_function():
    """The algebra behind kpd_fpn0.cu, checked in fp64 on the CPU: conv3x3(L0 f0 + up4(lat1)) with zero padding equals
    (A) a 3x3 conv of f0 with weights composed through L0 plus (B) for every sub-position (py,px) of a 4x4 block the
    <= 4 coarse neighbours of lat1 times the sums of the taps that land in them -- 36 matrices in total, enumerated in
    the order kpd_plan.cu packs them ((py, px, ny, nx) nested)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    B, H0, Hc, C = 2, 16, 4, 128
    f0 = torch.randn(B, 16, H0, H0, dtype=torch.float64, generator=g)
    lat1 = torch.randn(B, C, Hc, Hc, dtype=torch.float64, generator=g)
    W0 = torch.randn(C, 16, dtype=torch.float64, generator=g)
    Wc = torch.randn(C, C, 3, 3, dtype=torch.float64, generator=g)
    lateral0 = torch.einsum("oc,bchw->bohw", W0, f0) + F.interpolate(lat1, scale_factor=4, mode="nearest")
    want = F.conv2d(lateral0, Wc, padding=1)

    def nbr(p, d):
        q = p + d
        return -1 if q < 0 else (1 if q > 3 else 0)

    got = F.conv2d(f0, torch.einsum("octs,ci->oits", Wc, W0), padding=1)           # (A): K = 9 x 16
    lat1p = F.pad(lat1, (1, 1, 1, 1))                                              # out-of-map coarse neighbours read zeros
    n_mats = 0
    for py in range(4):
        for px in range(4):
            for ny in (-1, 0, 1):
                for nx in (-1, 0, 1):
                    taps = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if nbr(py, dy) == ny and nbr(px, dx) == nx]
                    if not taps:
                        continue
                    n_mats += 1
                    A = sum(Wc[:, :, dy + 1, dx + 1] for dy, dx in taps)
                    src = lat1p[:, :, 1 + ny:1 + ny + Hc, 1 + nx:1 + nx + Hc]
                    got[:, :, py::4, px::4] += torch.einsum("oc,bchw->bohw", A, src)
    assert n_mats == 36
    assert float((got - want).abs().max()) < 1e-10
