"""Pins the CPU oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py, run in the build container against /root/reference)."""
import os
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import model as OM, decode as OD
from oracle.arch import state_dict_spec
from oracle.weights import synth_state_dict, synth_images, synth_boxes, synth_nms_candidates
from oracle.golden_util import sample_big, checksum, GOLDEN_CASES


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def sd():
    return synth_state_dict(seed=0)


def test_state_dict_contract_matches_reference_listing():
    """A0: 318 keys, reference order, shapes and dtypes (tests/golden/state_dict_keys.txt)."""
    lines = open(os.path.join(GOLDEN, "state_dict_keys.txt")).read().strip().splitlines()
    spec = state_dict_spec(in_channels=1)
    assert len(lines) == 318 == len(spec)
    for line, (k, (shape, kind)) in zip(lines, spec.items()):
        name, rest = line.split(" ", 1)
        assert name == k
        assert rest.startswith(str(list(shape)))
        assert rest.endswith("int64" if kind == "nbt" else "float32")


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_forward_stages_match_reference(sd, case):
    kw = GOLDEN_CASES[case]
    g = _load(f"forward_{case}.npz")
    x = synth_images(kw["batch"], 1, kw["size"], seed=1234)
    boxes = synth_boxes(kw["batch"], kw["persons"], seed=99, zero_slots=kw["zero_slots"])
    np.testing.assert_array_equal(boxes, g["boxes"])
    out = OM.forward(sd, torch.from_numpy(x), torch.from_numpy(boxes), return_stages=True)
    st = out["stages"]
    for i in range(4):
        np.testing.assert_allclose(sample_big(st["feats"][i]), g[f"feat{i}_s"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(checksum(st["feats"][i]), g[f"feat{i}_c"], rtol=1e-4)
    np.testing.assert_allclose(sample_big(st["fpn0"]), g["fpn0_s"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(checksum(st["fpn0"]), g["fpn0_c"], rtol=1e-4)
    np.testing.assert_allclose(st["scores"].numpy(), g["scores"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(st["topk"].numpy(), g["topk"])
    if st["roi"]:
        roi = torch.cat(st["roi"], 0)
        np.testing.assert_allclose(sample_big(roi), g["roi_s"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(checksum(roi), g["roi_c"], rtol=1e-4)
        np.testing.assert_allclose(torch.cat(st["cw"], 0).numpy(), g["cw"], atol=2e-6)
        np.testing.assert_allclose(torch.cat(st["sw"], 0).numpy(), g["sw"], atol=2e-5)
    assert list(out["heatmap"].shape) == list(g["heatmap_shape"])
    if "heatmap" in g:
        np.testing.assert_allclose(out["heatmap"].numpy(), g["heatmap"], rtol=1e-3, atol=1e-6)
    else:
        np.testing.assert_allclose(sample_big(out["heatmap"].reshape(-1, 17, 56, 56)), g["heatmap_s"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(checksum(out["heatmap"]), g["heatmap_c"], rtol=1e-4)
    np.testing.assert_allclose(out["keypoints"].numpy(), g["keypoints"], atol=2e-6)
    np.testing.assert_array_equal(out["visibilities"].numpy(), g["visibilities"])
    hm = out["heatmap"].reshape(-1, 17, 56 * 56)
    np.testing.assert_array_equal(hm.argmax(dim=2).numpy(), g["heatmap_argmax"])


def test_forward_stock_random_init_matches_reference():
    g = _load("forward_stock_s224_b2_p3.npz")
    sd = synth_state_dict(seed=0, style="stock")
    x = synth_images(2, 1, 224, seed=1234)
    out = OM.forward(sd, torch.from_numpy(x), torch.from_numpy(g["boxes"]), topk_override=g["topk"], return_stages=True)
    np.testing.assert_allclose(out["stages"]["scores"].numpy(), g["scores"], atol=2e-7)
    np.testing.assert_allclose(out["heatmap"].numpy(), g["heatmap"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(out["keypoints"].numpy(), g["keypoints"], atol=2e-6)
    np.testing.assert_array_equal(out["visibilities"].numpy(), g["visibilities"])


def test_forward_without_boxes_returns_reference_empty_shapes(sd):
    g = _load("forward_no_boxes.npz")
    x = synth_images(2, 1, 224, seed=1234)
    out = OM.forward(sd, torch.from_numpy(x), torch.zeros(2, 0, 4))
    assert list(out["keypoints"].shape) == list(g["keypoints_shape"])
    assert list(out["visibilities"].shape) == list(g["visibilities_shape"])
    assert list(out["heatmap"].shape) == list(g["heatmap_shape"])
    assert float(g["total"][0]) == 0.0 and float(out["heatmap"].abs().sum()) == 0.0


def test_decoders_match_reference():
    g = _load("decoders.npz")
    hm = g["heatmaps"]
    kp, sc, idx = OD.decode_heatmaps(hm)
    np.testing.assert_array_equal(idx, g["argmax"])
    np.testing.assert_array_equal(kp, g["plain_kp"])
    np.testing.assert_array_equal(sc, g["plain_sc"])
    kp, sc, idx = OD.decode_heatmaps_subpixel(hm)
    np.testing.assert_array_equal(idx, g["argmax"])
    np.testing.assert_allclose(kp, g["subpix_kp"], atol=1e-6)
    np.testing.assert_array_equal(sc, g["subpix_sc"])
    assert np.all(kp[0, 0] == 0) and sc[0, 0] == 0          # zero-mass branch
    kp, sc = OD.decode_heatmaps_soft_argmax(hm, temperature=float(g["soft_temperature"][0]))
    np.testing.assert_allclose(kp, g["soft_kp"], atol=1e-5)   # fp32 sum over 3072 terms, order differs
    np.testing.assert_array_equal(sc, g["soft_sc"])
    kp, vis, conf = OD.model_decode(hm)
    np.testing.assert_allclose(kp, g["model_kp"], atol=2e-6)
    np.testing.assert_array_equal(vis, g["model_vis"])
    kp3, sc3, _ = OD.decode_heatmaps(hm[0])
    np.testing.assert_array_equal(kp3[0], g["plain3d_kp"])
    # torch-side restatement of the model decoder agrees too
    kpt, vist = OM.decode_heatmap(torch.from_numpy(hm))
    np.testing.assert_allclose(kpt.numpy(), g["model_kp"], atol=2e-6)
    np.testing.assert_array_equal(vist.numpy(), g["model_vis"])


def test_nms_and_iou_match_reference():
    g = _load("nms.npz")
    boxes, scores = synth_nms_candidates(8, 100, seed=5)
    np.testing.assert_array_equal(boxes, g["boxes"])
    for i in range(8):
        keep = OD.non_max_suppression(boxes[i], scores[i], 0.3, 100)
        assert len(keep) == g["count"][i]
        np.testing.assert_array_equal(keep, g["keep"][i, :len(keep)])
    np.testing.assert_array_equal(OD.non_max_suppression(boxes[0], scores[0], 0.5, 20), g["keep_thr05_max20"])
    np.testing.assert_allclose(OD.box_iou(boxes[0][:16], boxes[1][:24]), g["iou_16x24"], atol=1e-6)


def test_roi_align_restatement_matches_torchvision():
    g = _load("roi_align.npz")
    feat = torch.from_numpy(g["feat"])
    for n, roi in enumerate(g["rois"]):
        out = OM.roi_align_legacy(feat[int(roi[0])], roi[1:].reshape(1, 4))
        np.testing.assert_allclose(out[0].numpy(), g["out"][n], rtol=1e-5, atol=2e-6)
    big = np.random.default_rng(42)
    big.standard_normal((2, 8, 64, 64))      # replay the generator stream up to the 256x256 map
    fb = torch.from_numpy(big.standard_normal((1, 4, 256, 256)).astype(np.float32))
    for n, roi in enumerate(g["rois_big"]):
        out = OM.roi_align_legacy(fb[0], roi[1:].reshape(1, 4))
        np.testing.assert_allclose(out[0].numpy(), g["out_big"][n], rtol=1e-5, atol=2e-6)


def test_person_heads_match_reference(sd):
    g = _load("person_heads.npz")
    box, cls = OM.person_heads(sd, torch.from_numpy(g["x"]))
    np.testing.assert_allclose(box.numpy(), g["box"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cls.numpy(), g["cls"], rtol=1e-5, atol=1e-6)


def test_detect_restatement_properties():
    """oracle.detect states NEW semantics (no reference exists for the bare-tensor path): check its defining properties."""
    from oracle.detect import detect_persons
    H = W = 8; S = 16
    rng = np.random.default_rng(5)
    box = (0.1 * rng.standard_normal((2, H * W, 36))).astype(np.float32)
    cls = (-4 + 0.1 * rng.standard_normal((2, H * W, 9))).astype(np.float32)      # everything below the confidence threshold ...
    cls[0, 3 * W + 4, 4] = 5.0; cls[0, 3 * W + 4, 5] = 4.0                          # ... except two overlapping anchors of one pixel
    cls[0, 6 * W + 1, 1] = 3.0                                                      # and one far away
    boxes, count = detect_persons(box, cls, H, W, S, (2, 4, 8), 0.3, 0.3, 100, 4)
    assert count.tolist() == [2, 0] and boxes.shape == (2, 4, 4)
    assert np.all(boxes[1] == 0) and np.all(boxes[0, 2:] == 0)
    # best candidate first: pixel (3,4), anchor 4 = size 4, ratio 1 -> centre near (4.5/8, 3.5/8), side near 4/16
    assert abs(boxes[0, 0, 0] - 4.5 / 8) < 0.05 and abs(boxes[0, 0, 1] - 3.5 / 8) < 0.05 and abs(boxes[0, 0, 2] - 0.25) < 0.05
    assert abs(boxes[0, 1, 0] - 1.5 / 8) < 0.08 and abs(boxes[0, 1, 1] - 6.5 / 8) < 0.08      # the overlapping second anchor was suppressed
