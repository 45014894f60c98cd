"""scripts/predict.py (reference: scripts/predict.py:30-226): argument / IO logic on the CPU, and an end-to-end run of the
CLI on a B200 with a reference-format training checkpoint, a PNG and a YOLO-pose label file."""
import importlib.util
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT


def _cli():
    spec = importlib.util.spec_from_file_location("kpd_predict_cli", os.path.join(ROOT, "scripts", "predict.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LABEL = "0 0.5 0.45 0.6 0.8 " + " ".join(f"{0.3 + 0.02 * i:.3f} {0.2 + 0.03 * i:.3f} {i % 3}" for i in range(17)) + "\n"


def test_gt_box_parsing(tmp_path):
    cli = _cli()
    p = tmp_path / "a.txt"
    p.write_text(LABEL + "0 0.2 0.3 0.1 0.2\n" + "\n" + "garbage line\n")
    b = cli.read_gt_boxes(p)
    assert b.shape == (2, 4) and b.dtype == torch.float32
    np.testing.assert_allclose(b.numpy(), [[0.5, 0.45, 0.6, 0.8], [0.2, 0.3, 0.1, 0.2]], rtol=1e-6)
    assert cli.read_gt_boxes(tmp_path / "missing.txt") is None and cli.read_gt_boxes(None) is None
    (tmp_path / "empty.txt").write_text("\n")
    assert cli.read_gt_boxes(tmp_path / "empty.txt") is None
    # the model-side helper of the same label format (keypoint_model.py:587-628)
    from dll.models.keypoint_model import parse_ground_truth
    g = parse_ground_truth(p)
    assert g['bboxes'][0].shape == (1, 4) and g['keypoints'].shape == (1, 1, 17, 2) and g['visibilities'].shape == (1, 1, 17)
    assert g['visibilities'][0, 0].tolist() == [i % 3 for i in range(17)]
    np.testing.assert_allclose(g['keypoints'][0, 0, 3].numpy(), [0.36, 0.29], rtol=1e-6)


def test_checkpoint_forms():
    cli = _cli()
    sd = {"a.weight": torch.ones(2), "b.bias": torch.zeros(1)}
    ckpt = {"epoch": 3, "model_state_dict": dict(sd, **{"loss_fn.extra": torch.ones(1)}), "optimizer_state_dict": {}, "history": {"loss": [1.0]}}
    got = cli.extract_state_dict(ckpt, sd.keys())
    assert sorted(got) == sorted(sd)                        # unknown keys dropped (predict.py:50-53)
    assert cli.extract_state_dict(sd, sd.keys()) is sd      # bare state_dict passes through (predict.py:55-57)


def test_printout_and_input_listing(tmp_path):
    cli = _cli()
    k = np.arange(34, dtype=np.float32).reshape(17, 2) / 100
    lines = cli.format_keypoints(k)
    assert lines[1] == "Predicted Keypoints:" and lines[3] == "Keypoints shape: (17, 2)"
    assert lines[4] == " 1. nose            (0.000, 0.010)" and lines[-1] == "17. right_ankle     (0.320, 0.330)"
    assert len(lines) == 4 + 17
    assert cli.format_keypoints(np.stack([k, k + 1]))[4] == lines[4]       # [person, joint, xy] -> first person
    assert len(cli.KEYPOINT_NAMES) == 17
    for n in ("b.png", "a.jpg", "c.txt", "d.JPEG"):
        (tmp_path / n).write_bytes(b"")
    (tmp_path / "labels").mkdir()
    pairs = list(cli.iter_inputs(tmp_path, tmp_path / "labels"))
    assert [p.name for p, _ in pairs] == ["a.jpg", "b.png", "d.JPEG"]
    assert pairs[0][1] == tmp_path / "labels" / "a.txt"
    single = list(cli.iter_inputs(tmp_path / "a.jpg", None))
    assert single == [(tmp_path / "a.jpg", None)]


def test_transform_shapes():
    from dll.data import ITransform
    rng = np.random.default_rng(0)
    img = rng.integers(0, 255, (120, 160, 3), dtype=np.uint8)
    x = ITransform(224, 2.0, (8, 8))(img)
    assert x.shape == (1, 224, 224) and x.dtype == torch.float32 and -1.0 <= float(x.min()) and float(x.max()) <= 1.0
    assert ITransform(64, grayscale=False)(img).shape == (3, 64, 64)


@pytest.mark.gpu
def test_cli_end_to_end_with_reference_format_checkpoint(tmp_path, capsys):
    from PIL import Image
    from oracle.weights import synth_state_dict
    cli = _cli()
    sd = synth_state_dict(seed=0)
    ckpt = tmp_path / "checkpoint_epoch_3.pth"
    torch.save({"epoch": 3, "model_state_dict": dict(sd, **{"loss_fn.some_buffer": torch.zeros(1)}),
                "optimizer_state_dict": {"state": {}}, "history": {"train_loss": [0.5, 0.4]}}, ckpt)
    bare = tmp_path / "bare.pth"
    torch.save(sd, bare)
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:240, 0:320]
    img = ((np.sin(xx / 15.0) + np.cos(yy / 21.0)) * 60 + 128 + rng.normal(0, 12, (240, 320))).clip(0, 255).astype(np.uint8)
    (tmp_path / "images").mkdir(); (tmp_path / "labels").mkdir()
    Image.fromarray(np.stack([img, img[::-1], img[:, ::-1]], -1)).save(tmp_path / "images" / "p1.png")
    (tmp_path / "labels" / "p1.txt").write_text(LABEL + "0 0.3 0.3 0.4 0.5\n")
    cfg = os.path.join(ROOT, "configs", "default_config.yaml")
    results = {}
    for name, path in (("ckpt", ckpt), ("bare", bare)):
        out_dir = tmp_path / f"out_{name}"
        rc = cli.main(["--config", cfg, "--model", str(path), "--input", str(tmp_path / "images"), "--gt", str(tmp_path / "labels"),
                       "--output", str(out_dir), "--precision", "fp32_strict"])
        assert rc == 0
        text = capsys.readouterr().out
        assert "Predicted Keypoints:" in text and " 1. nose" in text and "17. right_ankle" in text
        results[name] = json.load(open(out_dir / "p1_keypoints.json"))
    assert results["ckpt"]["keypoints"] == results["bare"]["keypoints"]
    kp = np.asarray(results["ckpt"]["keypoints"])
    assert kp.shape == (2, 17, 2) and np.all((kp >= 0) & (kp <= 1))
    # the same numbers as the model API on the same transformed image (and as the CPU oracle)
    from dll.data import ITransform
    from oracle import model as OM
    x = ITransform(224, 2.0, (8, 8))(Image.open(tmp_path / "images" / "p1.png").convert("RGB")).unsqueeze(0)
    boxes = cli.read_gt_boxes(tmp_path / "labels" / "p1.txt").unsqueeze(0)
    ref = OM.forward(sd, x, boxes)
    np.testing.assert_allclose(kp, ref["keypoints"].numpy().reshape(2, 17, 2), atol=2e-5)


@pytest.mark.gpu
def test_channel_attention_and_select_top_k_channels():
    """ChannelAttention.forward + select_top_k_channels (keypoint_model.py:33-44, :653-661) as stand-alone calls."""
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    from dll.models.keypoint_model import select_top_k_channels
    from oracle import model as OM
    from oracle.weights import synth_state_dict
    sd = synth_state_dict(seed=0)
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    m = MultiPersonKeypointModel(cfg.model, cfg.training)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(5)
    f = torch.randn(3, 128, 28, 20, generator=g).clamp_min(0)
    scores = m.channel_attention(f.cuda()).cpu()
    want = OM.channel_scores(sd, "channel_attention", f)
    np.testing.assert_allclose(scores.numpy(), want.numpy(), atol=2e-6)
    sel = select_top_k_channels(f.cuda(), m.channel_attention, k=64).cpu()
    idx = torch.topk(want, 64, dim=1).indices
    gaps = torch.sort(want, dim=1, descending=True).values
    clear = (gaps[:, :64] - gaps[:, 1:65]).min(dim=1).values > 1e-5
    assert sel.shape == (3, 64, 28, 20)
    for b in range(3):
        if clear[b]:
            assert torch.equal(sel[b], f[b, idx[b]])
