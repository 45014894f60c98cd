"""the `BACKBONE` and `KEYPOINT_HEAD` classes the reference exports but never instantiates, on the generic fp32
layer kernels.  CPU part: module trees carry the reference's state_dict keys.  GPU part (`-m gpu`): outputs against vectors
produced by the reference's own classes (tests/golden/extra_modules.npz, made by make_extra_modules_golden.py)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

TOL = 2e-4        # fp32 FMA with a different summation order than the reference's oneDNN / cuDNN kernels; relative to the map's range


def _gold():
    return np.load(os.path.join(GOLDEN, "extra_modules.npz"))


def _mk():
    spec = importlib.util.spec_from_file_location("mk_extra", os.path.join(GOLDEN, "make_extra_modules_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _build(device="cpu"):
    from dll.configs import BackboneConfig, KeypointHeadConfig
    from dll.models import BACKBONE, KEYPOINT_HEAD
    mk = _mk()
    bb = mk.load_synth(BACKBONE(BackboneConfig(width_mult=1.0, in_channels=1)), 5).to(device)
    kh = mk.load_synth(KEYPOINT_HEAD(KeypointHeadConfig()), 6).to(device)
    return bb, kh, mk


def test_state_dict_keys_match_the_reference():
    bb, kh, _ = _build()
    g = _gold()
    assert list(bb.state_dict().keys()) == list(g["backbone.keys"])          # strict load of a reference checkpoint and back
    assert list(kh.state_dict().keys()) == list(g["keypoint_head.keys"])
    from dll.models import __all__ as exported
    assert {"BACKBONE", "KEYPOINT_HEAD", "MobileNetV3Wrapper", "HeatmapHead", "PERSON_HEAD", "MultiPersonKeypointModel"} <= set(exported)


def test_no_cpu_fallback():
    bb, kh, _ = _build()
    from dll._native import NativeLibraryError
    with pytest.raises(NativeLibraryError):
        bb(torch.zeros(1, 1, 64, 64))
    with pytest.raises(NativeLibraryError):
        kh(torch.zeros(1, 128, 8, 8))


def _err(got, ref):
    ref = torch.from_numpy(ref)
    return float((got.detach().cpu() - ref).abs().max() / max(1.0, float(ref.abs().max())))


@pytest.mark.gpu
def test_backbone_matches_reference_vectors():
    bb, _, mk = _build("cuda")
    g = _gold()
    for name, b, s, seed in mk.BACKBONE_INPUTS:
        x = torch.randn(b, 1, s, s, generator=torch.Generator().manual_seed(seed)).cuda()
        maps = bb(x)
        torch.cuda.synchronize()
        assert len(maps) == 4
        for i, m in enumerate(maps):
            assert tuple(m.shape) == g[f"{name}.fpn{i}"].shape
            assert _err(m, g[f"{name}.fpn{i}"]) < TOL, (name, "fpn", i)
        for i, m in enumerate(bb.backbone_outputs):
            assert _err(m, g[f"{name}.feat{i}"]) < TOL, (name, "feat", i)
    from dll.configs import BackboneConfig
    from dll.models import BACKBONE
    wide = BACKBONE(BackboneConfig(width_mult=1.5, in_channels=1)).cuda().eval()      # the reference's default width: its FPN rejects it
    with pytest.raises(RuntimeError):
        wide(torch.zeros(1, 1, 64, 64, device="cuda"))


@pytest.mark.gpu
def test_keypoint_head_matches_reference_vectors():
    _, kh, mk = _build("cuda")
    g = _gold()
    for name, b, s, seed in mk.HEAD_INPUTS:
        x = torch.randn(b, 128, s, s, generator=torch.Generator().manual_seed(seed)).cuda()
        k, v = kh(x)
        torch.cuda.synchronize()
        assert tuple(k.shape) == (b, 17, 2) and tuple(v.shape) == (b, 17, 3)
        assert _err(k, g[f"{name}.keypoints"]) < TOL and _err(v, g[f"{name}.visibility"]) < TOL, name
    # batch independence / determinism of the generic kernels
    x = torch.randn(4, 128, 28, 28, generator=torch.Generator().manual_seed(3)).cuda()
    k4, v4 = kh(x)
    k1, v1 = kh(x[2:3])
    assert torch.equal(k4[2:3], k1) and torch.equal(v4[2:3], v1)


@pytest.mark.gpu
def test_submodules_standalone():
    """The building blocks are public in the reference module too: SEModule, ConvBNActivation, InvertedResidual against torch."""
    import torch.nn.functional as F
    from dll.models.backbone import SEModule, InvertedResidual
    torch.manual_seed(0)
    se = SEModule(48, reduction=4).cuda().eval()
    for p in se.parameters():
        torch.nn.init.normal_(p, 0, 0.3)
    x = torch.randn(2, 48, 9, 7, device="cuda")
    y = x.mean((2, 3))
    ref = x * F.hardsigmoid(F.relu(y @ se.fc[0].weight.t()) @ se.fc[2].weight.t()).view(2, 48, 1, 1)
    assert float((se(x) - ref).detach().abs().max()) < 1e-5
    blk = InvertedResidual(24, 24, 5, 1, 3, 0.25, torch.nn.Hardswish).cuda().eval()
    for m in blk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.7, 1.3); m.bias.data.normal_(0, 0.1)
    x = torch.randn(2, 24, 11, 11, device="cuda")
    e, d, s, pj = blk.block
    t = F.hardswish(F.batch_norm(F.conv2d(x, e.layer[0].weight), e.layer[1].running_mean, e.layer[1].running_var, e.layer[1].weight, e.layer[1].bias, False, 0, e.layer[1].eps))
    t = F.hardswish(F.batch_norm(F.conv2d(t, d.layer[0].weight, None, 1, 2, 1, t.shape[1]), d.layer[1].running_mean, d.layer[1].running_var, d.layer[1].weight, d.layer[1].bias, False, 0, d.layer[1].eps))
    t = t * F.hardsigmoid(F.relu(t.mean((2, 3)) @ s.fc[0].weight.t()) @ s.fc[2].weight.t()).view(2, -1, 1, 1)
    ref = F.batch_norm(F.conv2d(t, pj.layer[0].weight), pj.layer[1].running_mean, pj.layer[1].running_var, pj.layer[1].weight, pj.layer[1].bias, False, 0, pj.layer[1].eps) + x
    assert float((blk(x) - ref).detach().abs().max()) < 1e-4
