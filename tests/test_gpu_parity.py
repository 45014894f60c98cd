"""GPU parity tests: the CUDA path (through the `dll` drop-in -> ctypes -> C-ABI) against the CPU oracle and the
golden vectors produced by the reference.  Run with `pytest -m gpu` on a B200.

Tolerances (BASELINE.json north_star):
  * heat maps, reference-style random-init weights ("stock", the north star's literal condition): point-wise relative
    error <= 1e-3 in "fp32" mode (tcgen05 kind::tf32, operands rounded to nearest) and <= 2e-2 in bf16 mode;
    strict-fp32 (CUDA-core) mode <= 1e-4.
  * heat maps, calibrated stress weights (BN de-tuned, logits sharpened ~40x so maps span 0..1): strict mode <= 1e-4
    of the map's range; the reduced-precision modes are held to the error their operand rounding implies after the
    logit gain (<= 1e-2 tf32, <= 1e-1 bf16 of the range: the bf16 figure moves between 2e-2 and 7e-2 with last-bit changes
    upstream, because the ~40x logit gain turns bf16 operand rounding into near-random heat-map noise) -- measured values are written to parity_report.json.
  * channel top-k is teacher-forced from the oracle for the reduced-precision modes and compared
    free-running where score gaps allow.
  * integer outputs (top-k order, argmax, NMS keep lists, visibility class): bit-exact, except where the oracle's
    top-2 values / threshold distance are within the tolerance.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

from oracle import model as OM, decode as OD
from oracle.weights import (synth_state_dict, synth_images, synth_boxes, synth_blob_heatmaps,
                            synth_nms_candidates)
from oracle.golden_util import GOLDEN_CASES

REPORT = {}


def _report(key, value):
    REPORT[key] = value
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def sd():
    return synth_state_dict(seed=0)


@pytest.fixture(scope="module")
def model(sd):
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    m = MultiPersonKeypointModel(cfg.model, cfg.training)
    missing = m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


_ORACLE_CACHE = {}


def _oracle(sd, case):
    if case not in _ORACLE_CACHE:
        kw = GOLDEN_CASES[case]
        x = synth_images(kw["batch"], 1, kw["size"], seed=1234)
        boxes = synth_boxes(kw["batch"], kw["persons"], seed=99, zero_slots=kw["zero_slots"])
        out = OM.forward(sd, torch.from_numpy(x), torch.from_numpy(boxes), return_stages=True)
        _ORACLE_CACHE[case] = (x, boxes, out)
    return _ORACLE_CACHE[case]


def nchw(t):
    return t.permute(0, 3, 1, 2).float().cpu().numpy()


def rel_range_err(a, b):
    """max |a-b| / max|b| (error relative to the tensor's dynamic range)."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def heatmap_errors(a, b):
    """(range-relative error per map, max point-wise relative error over values above 10 % of the map peak)."""
    a = a.reshape(-1, 56 * 56); b = b.reshape(-1, 56 * 56)
    peak = np.abs(b).max(axis=1, keepdims=True)
    live = peak[:, 0] > 0
    rr = (np.abs(a - b).max(axis=1)[live] / peak[live, 0]).max() if live.any() else 0.0
    mask = (b > 0.1 * peak) & (peak > 0)
    pw = (np.abs(a - b)[mask] / b[mask]).max() if mask.any() else 0.0
    return float(rr), float(pw)


def check_discrete(got_idx, oracle_maps, tol, what):
    """argmax must equal the oracle's except where the oracle's top-2 values differ by < tol (relative)."""
    flat = oracle_maps.reshape(-1, oracle_maps.shape[-2] * oracle_maps.shape[-1])
    part = np.partition(flat, -2, axis=1)
    top1, top2 = part[:, -1], part[:, -2]
    exempt = (top1 - top2) <= tol * np.maximum(np.abs(top1), 1e-30)
    want = flat.argmax(axis=1)
    got = np.asarray(got_idx).reshape(-1)
    bad = (got != want) & ~exempt
    assert not bad.any(), f"{what}: {bad.sum()} argmax mismatches outside the near-tie exemption"
    return int(exempt.sum()), int((got != want).sum())


MODES = [("fp32_strict", 1e-4, False), ("fp32", 1e-2, True), ("bf16", 1e-1, True)]          # calibrated stress weights
STOCK_MODES = [("fp32_strict", 1e-4), ("fp32", 1e-3), ("bf16", 2e-2)]                          # north-star tolerances


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
@pytest.mark.parametrize("mode,tol,force_topk", MODES)
def test_forward_stagewise(model, sd, case, mode, tol, force_topk):
    x, boxes, ref = _oracle(sd, case)
    st = ref["stages"]
    model.set_precision(mode)
    override = st["topk"].to(torch.int32) if force_topk else None
    out = model.predict({"image": torch.from_numpy(x), "bboxes": torch.from_numpy(boxes)}, topk_override=override)
    torch.cuda.synchronize()
    g = model.stage_tensors()
    rep = {}
    REPORT[f"forward/{case}/{mode}"] = rep
    # --- trunk (always fp32 body; FPN-0 conv follows the mode)
    for i in range(4):
        e = rel_range_err(nchw(g[f"feat{i}"]), st["feats"][i].numpy()); rep[f"feat{i}"] = e
        assert e < (2e-5 if mode == "fp32_strict" else 1e-4), f"feat{i} range-relative error {e}"   # same exact-fp32 trunk in every mode; the looser bound is historical head-room
    trunk_tol = 1e-4 if mode == "fp32_strict" else 1e-3
    e = rel_range_err(nchw(g["fpn0"]), st["fpn0"].numpy()); rep["fpn0"] = e
    assert e < trunk_tol, f"fpn0 error {e}"
    ds = float(np.abs(g["scores"].cpu().numpy() - st["scores"].numpy()).max()); rep["scores_abs"] = ds
    assert ds < (2e-5 if mode == "fp32_strict" else 1e-3)
    # --- free-running top-k: exact unless a score gap among the top-65 is within 10x the observed score error
    srt = np.sort(st["scores"].numpy(), axis=1)[:, ::-1]
    gaps = (srt[:, :64] - srt[:, 1:65]).min(axis=1)
    same = (g["topk"].cpu().numpy() == st["topk"].numpy()).all(axis=1)
    exempt = gaps < 10 * max(ds, 1e-7)
    rep["topk_images_equal"] = int(same.sum()); rep["topk_images_exempt"] = int(exempt.sum())
    assert (same | exempt).all(), f"top-k order differs on images with clear score gaps: gaps={gaps} ds={ds}"
    # --- ROI stage, head, decode
    n = len(st["roi"])
    assert int(g["roi_count"].item()) == n
    if n:
        roi = torch.cat(st["roi"], 0).numpy()
        teacher = force_topk or bool(same.all())
        if teacher:
            e = rel_range_err(nchw(g["roi_feat"][:n]), roi); rep["roi_feat"] = e
            assert e < trunk_tol
            e = float(np.abs(g["head_cw"][:n].cpu().numpy() - torch.cat(st["cw"], 0).numpy().reshape(n, 64)).max()); rep["cw_abs"] = e
            assert e < max(tol, 1e-4)
            e = float(np.abs(g["head_sw"][:n].cpu().numpy() - torch.cat(st["sw"], 0).numpy().reshape(n, 56, 56)).max()); rep["sw_abs"] = e
            assert e < max(tol, 1e-4)
            hm = out["heatmap"].cpu().numpy()
            rr, pw = heatmap_errors(hm, ref["heatmap"].numpy()); rep["heatmap_range_rel"] = rr; rep["heatmap_pointwise_rel"] = pw
            _report(f"forward/{case}/{mode}", rep)
            assert rr < tol, f"heat-map range-relative error {rr} >= {tol}"
            assert pw < 5 * tol, f"heat-map point-wise relative error {pw}"
            assert list(out["heatmap"].shape) == list(ref["heatmap"].shape)
            kerr = float(np.abs(out["keypoints"].cpu().numpy() - ref["keypoints"].numpy()).max()); rep["keypoints_abs"] = kerr
            assert kerr < max(1e-5, tol * 1e-1)
            # visibility classes: exact unless sigmoid(max) is within tol of a threshold
            v_got = out["visibilities"].cpu().numpy(); v_ref = ref["visibilities"].numpy()
            mx = ref["heatmap"].numpy().reshape(*ref["heatmap"].shape[:3], -1).max(-1)
            conf = 1 / (1 + np.exp(-mx))
            near = (np.abs(conf - 0.3) < tol) | (np.abs(conf - 0.7) < tol)
            bad = (v_got != v_ref).any(-1)[:, :, 0] & ~near
            assert not bad.any()
            # argmax of the produced heat maps vs the oracle's (decode_heatmaps extension path)
            from dll.models.heatmap_head import decode_heatmaps_full
            full = decode_heatmaps_full(out["heatmap"].reshape(-1, 17, 56, 56))
            ex, diff = check_discrete(full["argmax"].cpu().numpy(), ref["heatmap"].numpy().reshape(-1, 56, 56), 2 * tol, "heat-map argmax")
            rep["argmax_exempt"] = ex; rep["argmax_diff"] = diff
    else:
        assert float(out["heatmap"].abs().sum()) == 0.0
    np.testing.assert_array_equal(out["visibilities"].cpu().numpy().shape, ref["visibilities"].numpy().shape)
    _report(f"forward/{case}/{mode}", rep)


@pytest.mark.parametrize("mode,tol", STOCK_MODES)
def test_forward_stock_random_init(mode, tol):
    """The north star's literal condition: reference-style random-init weights, heat maps within 1e-3 (fp32 mode) /
    2e-2 (bf16) RELATIVE tolerance of the reference's own output; top-k teacher-forced (scores are near-tied)."""
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    g = _golden("forward_stock_s224_b2_p3.npz")
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    m = MultiPersonKeypointModel(cfg.model, cfg.training, precision=mode)
    m.load_state_dict(synth_state_dict(seed=0, style="stock"), strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(synth_images(2, 1, 224, seed=1234))
    out = m.predict({"image": x, "bboxes": torch.from_numpy(g["boxes"])}, topk_override=torch.from_numpy(g["topk"]).to(torch.int32))
    hm, ref = out["heatmap"].cpu().numpy(), g["heatmap"]
    live = ref > 0
    rel = float((np.abs(hm - ref)[live] / ref[live]).max())
    ds = float(np.abs(m.stage_tensors()["scores"].cpu().numpy() - g["scores"]).max())
    same = int((m.stage_tensors()["topk"].cpu().numpy() == g["topk"]).all(axis=1).sum())
    _report(f"stock/{mode}", {"heatmap_pointwise_rel": rel, "scores_abs": ds, "topk_images_equal_free_running": same})
    assert rel < tol, f"heat-map relative error {rel} >= {tol}"
    assert float(np.abs(hm[~live]).max()) == 0.0
    np.testing.assert_allclose(out["keypoints"].cpu().numpy(), g["keypoints"], atol=max(1e-5, tol * 1e-2))
    np.testing.assert_array_equal(out["visibilities"].cpu().numpy(), g["visibilities"])
    assert ds < (2e-6 if mode == "fp32_strict" else 1e-4)


def test_forward_matches_golden_reference_outputs(model):
    """Directly against the reference's own outputs (not via the oracle): strict fp32 mode, free-running."""
    g = _golden("forward_s224_b2_p3.npz")
    kw = GOLDEN_CASES["s224_b2_p3"]
    model.set_precision("fp32_strict")
    x = synth_images(kw["batch"], 1, kw["size"], seed=1234)
    out = model.predict({"image": torch.from_numpy(x), "bboxes": torch.from_numpy(g["boxes"])})
    np.testing.assert_array_equal(model.stage_tensors()["topk"].cpu().numpy(), g["topk"])
    rr, pw = heatmap_errors(out["heatmap"].cpu().numpy(), g["heatmap"])
    assert rr < 1e-4 and pw < 1e-3, (rr, pw)
    np.testing.assert_allclose(out["keypoints"].cpu().numpy(), g["keypoints"], atol=1e-5)
    np.testing.assert_array_equal(out["visibilities"].cpu().numpy(), g["visibilities"])
    # padded slot (zero box in the middle of image 0) and output ranks
    assert out["heatmap"].shape == (2, 3, 17, 56, 56) and out["keypoints"].shape == (2, 3, 1, 17, 2)
    assert float(out["heatmap"][0, 2].abs().sum()) == 0.0 and float(out["visibilities"][0, 2].abs().sum()) == 0.0


def test_no_boxes_and_dummy_person(model):
    g = _golden("forward_no_boxes.npz")
    x = torch.from_numpy(synth_images(2, 1, 224, seed=1234))
    out = model.predict({"image": x, "bboxes": None})
    assert list(out["keypoints"].shape) == list(g["keypoints_shape"])
    assert list(out["visibilities"].shape) == list(g["visibilities_shape"])
    assert list(out["heatmap"].shape) == list(g["heatmap_shape"])
    g2 = _golden("forward_s224_b2_p2_empty_image.npz")
    model.set_precision("fp32_strict")
    out = model.predict({"image": x, "bboxes": [torch.from_numpy(g2["boxes"])]})       # list-wrapped form
    np.testing.assert_array_equal(out["visibilities"].cpu().numpy(), g2["visibilities"])
    assert out["visibilities"][0, 0, 0, :, 0].sum().item() == 17.0                      # dummy person: vis=[1,0,0]
    np.testing.assert_allclose(out["keypoints"].cpu().numpy(), g2["keypoints"], atol=1e-5)


def test_error_behaviour(model):
    with pytest.raises(TypeError):
        model.predict({"image": [1, 2, 3], "bboxes": None})
    x = torch.zeros(1, 1, 224, 224)
    with pytest.raises(ValueError):
        model.predict({"image": x, "bboxes": torch.zeros(1, 4)})
    with pytest.raises(ValueError):
        model.predict({"image": x, "bboxes": [torch.zeros(4)]})
    model.train()
    with pytest.raises(RuntimeError):
        model.forward({"image": x, "bboxes": torch.zeros(1, 1, 4)})
    model.eval()


def test_state_dict_reload_rebuilds_plan(model, sd):
    model.set_precision("fp32_strict")
    x = torch.from_numpy(synth_images(1, 1, 224, seed=3))
    b = torch.tensor([[[0.5, 0.5, 1.0, 1.0]]])
    h0 = model.predict({"image": x, "bboxes": b})["heatmap"].clone()
    sd2 = {k: (v * 1.5 if k == "heatmap_head.final_layer.3.weight" else v) for k, v in sd.items()}
    model.load_state_dict(sd2, strict=False)
    h1 = model.predict({"image": x, "bboxes": b})["heatmap"].clone()
    assert float((h0 - h1).abs().max()) > 1e-3
    model.load_state_dict({"model_state_dict": sd}["model_state_dict"], strict=False)    # checkpoint-dict form
    h2 = model.predict({"image": x, "bboxes": b})["heatmap"]
    assert float((h0 - h2).abs().max()) == 0.0      # the default path is bit-deterministic (no floating-point atomics)


# ------------------------------------------------------------------------------------------- decoders (A7, A8)
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_decoders(dtype):
    from dll.models import heatmap_head as HH
    g = _golden("decoders.npz")
    hm = g["heatmaps"]
    t = torch.from_numpy(hm).cuda()
    if dtype == "bf16":
        t = t.to(torch.bfloat16)
        hm = t.float().cpu().numpy()          # the oracle sees the same rounded values
        kp_ref, sc_ref, idx_ref = OD.decode_heatmaps(hm)
        sub_ref, subsc_ref, _ = OD.decode_heatmaps_subpixel(hm)
    else:
        kp_ref, sc_ref, idx_ref = g["plain_kp"], g["plain_sc"], g["argmax"]
        sub_ref, subsc_ref = g["subpix_kp"], g["subpix_sc"]
    full = HH.decode_heatmaps_full(t)
    np.testing.assert_array_equal(full["argmax"].cpu().numpy(), idx_ref)          # bit-exact incl. ties / corners
    kp, sc = HH.decode_heatmaps(t)
    np.testing.assert_array_equal(kp.cpu().numpy(), kp_ref)
    np.testing.assert_array_equal(sc.cpu().numpy(), sc_ref)
    kp, sc = HH.decode_heatmaps_subpixel(t)
    np.testing.assert_allclose(kp.cpu().numpy(), sub_ref, atol=2e-6)
    np.testing.assert_array_equal(sc.cpu().numpy(), subsc_ref)
    temp = float(g["soft_temperature"][0])
    soft_ref, softsc_ref = OD.decode_heatmaps_soft_argmax(hm, temp)
    kp, sc = HH.decode_heatmaps_soft_argmax(t, temperature=temp)
    np.testing.assert_allclose(kp.cpu().numpy(), soft_ref, atol=2e-5)
    np.testing.assert_array_equal(sc.cpu().numpy(), softsc_ref)
    mk_ref, mv_ref, _ = OD.model_decode(hm)
    kp, vis = HH.decode_model_heatmap(t)
    np.testing.assert_allclose(kp.cpu().numpy(), mk_ref, atol=2e-5)
    np.testing.assert_array_equal(vis.cpu().numpy(), mv_ref)
    if dtype == "fp32":
        np.testing.assert_allclose(kp.cpu().numpy(), g["model_kp"], atol=2e-5)
        np.testing.assert_array_equal(vis.cpu().numpy(), g["model_vis"])
        k3, s3 = HH.decode_heatmaps(t[0])                                           # 3-D input form
        assert k3.shape == (17, 2) and s3.shape == (17,)
        np.testing.assert_array_equal(k3.cpu().numpy(), g["plain3d_kp"])


def test_decode_large_batch_properties():
    """Size-independent properties on a larger seeded set: argmax == oracle, sub-pixel within half a pixel of argmax."""
    from dll.models import heatmap_head as HH
    hm = synth_blob_heatmaps(512, 17, 64, 48, seed=11)
    t = torch.from_numpy(hm).cuda()
    full = HH.decode_heatmaps_full(t)
    _, _, idx = OD.decode_heatmaps(hm)
    np.testing.assert_array_equal(full["argmax"].cpu().numpy(), idx)
    d = (full["kp_subpix"] - full["kp_plain"]).abs().cpu().numpy()
    assert d[..., 0].max() <= 1.0 / 47 and d[..., 1].max() <= 1.0 / 63
    assert torch.equal(full["vis"].sum(-1), torch.ones_like(full["vis"].sum(-1)))


# ------------------------------------------------------------------------------------------- person head (A10)
def test_nms_and_iou_bit_exact(model):
    g = _golden("nms.npz")
    ph = model.person_detector
    boxes, scores = torch.from_numpy(g["boxes"]), torch.from_numpy(g["scores"])
    keep, count = ph.batched_nms(boxes, scores, 0.3, 100)
    np.testing.assert_array_equal(count.cpu().numpy(), g["count"])
    np.testing.assert_array_equal(keep.cpu().numpy().astype(np.int64), g["keep"])
    k = ph.non_max_suppression(boxes[0], scores[0], 0.5, 20)
    assert k.dtype == torch.int64
    np.testing.assert_array_equal(k.numpy(), g["keep_thr05_max20"])
    iou = ph.box_iou(boxes[0][:16], boxes[1][:24])
    np.testing.assert_array_equal(iou.cpu().numpy(), OD.box_iou(g["boxes"][0][:16], g["boxes"][1][:24]))
    np.testing.assert_allclose(iou.cpu().numpy(), g["iou_16x24"], atol=1e-7)
    # full-size config C5: 128 images x 100 candidates against the oracle
    b, s = synth_nms_candidates(128, 100, seed=21)
    keep, count = ph.batched_nms(torch.from_numpy(b), torch.from_numpy(s), 0.3, 100)
    keep, count = keep.cpu().numpy(), count.cpu().numpy()
    for i in range(128):
        want = OD.non_max_suppression(b[i], s[i], 0.3, 100)
        assert count[i] == len(want)
        np.testing.assert_array_equal(keep[i, :count[i]], want)


def test_extract_roi_features_matches_torchvision_golden(model):
    """Stand-alone ROI-align method (A5) against torchvision's own outputs (tests/golden/roi_align.npz)."""
    g = _golden("roi_align.npz")
    feat = torch.from_numpy(g["feat"])                      # [2,8,64,64]
    for i, roi in enumerate(g["rois"]):
        b = int(roi[0]); x1, y1, x2, y2 = [float(v) / 64.0 for v in roi[1:]]
        box = torch.tensor([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dtype=torch.float32)
        corners = torch.stack([box[0] - box[2] / 2, box[1] - box[3] / 2, box[0] + box[2] / 2, box[1] + box[3] / 2]).clamp(0, 1) * 64
        want = OM.roi_align_legacy(feat[b], corners.view(1, 4).numpy())[0].numpy()   # oracle on the exact corners the method derives
        got = model.extract_roi_features(feat[b:b + 1], box)[0].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=5e-5)      # sample coordinates differ by 1 ulp (FMA contraction)
        if np.allclose(corners.numpy(), roi[1:], atol=1e-4):
            np.testing.assert_allclose(got, g["out"][i], rtol=1e-4, atol=1e-4)


def test_extract_roi_features_large_map_samplewise_kernel(model):
    """Maps beyond 392 pixels a side (adaptive grid > 7 samples per bin) take the sample-by-sample ROI-align kernel."""
    rng = np.random.default_rng(3)
    feat = torch.from_numpy(rng.standard_normal((1, 4, 460, 460)).astype(np.float32))
    for box in ([0.5, 0.5, 1.0, 1.0], [0.4, 0.55, 0.7, 0.5]):
        b = torch.tensor(box, dtype=torch.float32)
        corners = torch.stack([b[0] - b[2] / 2, b[1] - b[3] / 2, b[0] + b[2] / 2, b[1] + b[3] / 2]).clamp(0, 1) * 460
        want = OM.roi_align_legacy(feat[0], corners.view(1, 4).numpy())[0].numpy()
        got = model.extract_roi_features(feat, b)[0].cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=5e-5)


def test_person_heads(model, sd):
    g = _golden("person_heads.npz")
    box, cls = model.person_detector.heads(torch.from_numpy(g["x"]), level=0)
    np.testing.assert_allclose(box.cpu().numpy(), g["box"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(cls.cpu().numpy(), g["cls"], rtol=1e-4, atol=1e-5)


def test_standalone_modules(model, sd):
    """HeatmapHead.forward and MobileNetV3Wrapper.forward as standalone drop-ins."""
    x, boxes, ref = _oracle(sd, "s224_b2_p3")
    st = ref["stages"]
    head = model.heatmap_head
    head.precision = "fp32_strict"
    roi = torch.cat(st["roi"], 0)
    hm, (cw, sw) = head(roi.cuda())
    want, (cw_ref, sw_ref) = OM.heatmap_head_forward(sd, roi)
    rr, pw = heatmap_errors(hm.cpu().numpy(), want.numpy())
    assert rr < 1e-4 and pw < 1e-3
    np.testing.assert_allclose(cw.cpu().numpy(), cw_ref.numpy(), atol=1e-5)
    np.testing.assert_allclose(sw.cpu().numpy(), sw_ref.numpy(), atol=1e-5)
    outs = model.backbone(torch.from_numpy(x).cuda())
    refs = OM.fpn_forward(sd, OM.body_forward(sd, torch.from_numpy(x)))
    for a, b in zip(outs, refs):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_range_err(a.cpu().numpy(), b.numpy()) < 1e-3


# ------------------------------------------------------------------------------------------- tensor-core conv variants
@pytest.mark.parametrize("conv,n,hw", [(0, 2, 32), (0, 1, 112), (1, 3, 56), (2, 3, 56), (3, 2, 56), (3, 1, 40)])
def test_dense_conv_variants_agree(model, conv, n, hw):
    """kpd_dense_conv: classic tcgen05 (tf32 / bf16) and tap-reuse tcgen05 (chunk-major input) against the fp32 CUDA-core conv."""
    import ctypes as C
    from dll import _native as N
    model.set_precision("fp32_strict")
    plan, dev = model._plan()
    cin = {0: 128, 1: 64, 2: 256, 3: 256}[conv]; cout = {0: 128, 1: 256, 2: 256, 3: 64}[conv]
    g = torch.Generator(device="cuda").manual_seed(conv * 7 + hw)
    x = torch.randn(n, hw, hw, cin, device="cuda", generator=g).clamp_min(-0.5)

    def run(impl, inp, out_dtype):
        out = torch.empty(n, hw, hw, cout, device="cuda", dtype=out_dtype)
        N.check(N.load().kpd_dense_conv(plan.handle, conv, impl, N.ptr(inp), N.ptr(out), n, hw, hw, N.stream_ptr(dev)), "kpd_dense_conv")
        torch.cuda.synchronize()
        return out.float()

    def chunk_major(t):
        e = 16 // t.element_size()
        return t.view(n, hw, hw, cin // e, e).permute(0, 3, 1, 2, 4).contiguous()
    ref = run(0, x, torch.float32)
    peak = float(ref.abs().max())
    xb = x.to(torch.bfloat16)
    errs = {"tf32": float((run(1, x, torch.float32) - ref).abs().max()) / peak,
            "bf16": float((run(2, xb, torch.bfloat16) - ref).abs().max()) / peak,
            "tf32_halo": float((run(3, chunk_major(x), torch.float32) - ref).abs().max()) / peak,
            "bf16_halo": float((run(4, chunk_major(xb), torch.bfloat16) - ref).abs().max()) / peak}
    _report(f"dense_conv/{conv}/{n}x{hw}", errs)
    assert errs["tf32"] < 2e-3 and errs["tf32_halo"] < 2e-3, errs
    assert errs["bf16"] < 2e-2 and errs["bf16_halo"] < 2e-2, errs


def test_rgb_input_three_channels():
    """in_channels=3 (BackboneConfig default): stem conv over an NCHW RGB image, strict mode vs the oracle."""
    from dll import MultiPersonKeypointModel
    from dll.configs import ModelConfig, BackboneConfig, TrainingConfig
    from dll.configs.model_config import HeatmapHeadConfig
    cfg = ModelConfig(backbone=BackboneConfig(in_channels=3), heatmap_head=HeatmapHeadConfig(heatmap_size=(56, 56)))
    sd3 = synth_state_dict(seed=0, in_channels=3, style="stock")
    m = MultiPersonKeypointModel(cfg, TrainingConfig(), precision="fp32_strict")
    m.load_state_dict(sd3, strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(synth_images(1, 3, 224, seed=5))
    b = torch.from_numpy(synth_boxes(1, 2, seed=3))
    ref = OM.forward(sd3, x, b, return_stages=True)
    out = m.predict({"image": x, "bboxes": b}, topk_override=ref["stages"]["topk"].to(torch.int32))
    e = rel_range_err(nchw(m.stage_tensors()["feat0"]), ref["stages"]["feats"][0].numpy())
    assert e < 2e-5, e
    hm, r = out["heatmap"].cpu().numpy(), ref["heatmap"].numpy()
    assert float((np.abs(hm - r) / r).max()) < 1e-4
    np.testing.assert_allclose(out["keypoints"].cpu().numpy(), ref["keypoints"].numpy(), atol=1e-5)


@pytest.mark.parametrize("size,batch,persons", [(64, 3, 2), (96, 2, 1), (320, 2, 2)])
def test_forward_other_sizes_against_oracle(model, sd, size, batch, persons):
    """Sizes with no golden file (tiny 2x2 / 3x3 last maps, a non-224/512 size): CUDA path vs the oracle, strict mode
    free-running (top-k included) and fp32 mode teacher-forced."""
    x = torch.from_numpy(synth_images(batch, 1, size, seed=4321))
    boxes = torch.from_numpy(synth_boxes(batch, persons, seed=17, zero_slots=((0, persons - 1),) if persons > 1 else ()))
    ref = OM.forward(sd, x, boxes, return_stages=True)
    for mode, tol in (("fp32_strict", 1e-4), ("fp32", 1e-2)):
        model.set_precision(mode)
        out = model.predict({"image": x, "bboxes": boxes}, topk_override=ref["stages"]["topk"].to(torch.int32))
        st = model.stage_tensors()
        for i in range(4):
            e = rel_range_err(nchw(st[f"feat{i}"]), ref["stages"]["feats"][i].numpy())
            assert e < 1e-4, f"S={size} feat{i} {e}"
        rr, _ = heatmap_errors(out["heatmap"].cpu().numpy(), ref["heatmap"].numpy())
        assert rr < tol, f"S={size} {mode}: heat-map range error {rr}"
        assert float((out["keypoints"].cpu() - ref["keypoints"]).abs().max()) < max(tol, 1e-3)
        _report(f"other_sizes/S{size}/{mode}/heatmap_range_err", rr)
    model.set_precision("fp32_strict")
    out = model.predict({"image": x, "bboxes": boxes})
    srt = np.sort(ref["stages"]["scores"].numpy(), axis=1)[:, ::-1]
    clear = (srt[:, :64] - srt[:, 1:65]).min(axis=1) > 1e-5          # images whose top-64 boundary / order is not a near-tie
    same = (model.stage_tensors()["topk"].cpu().numpy() == ref["stages"]["topk"].numpy()).all(axis=1)
    assert (same | ~clear).all(), "free-running top-k differs on images with clear score gaps"


def test_forward_full_size_properties():
    """BASELINE config C2 at full size (B=256, 224x224) with ragged boxes, through size-independent properties:
    batch independence (an image's outputs do not depend on its batch-mates), run-to-run stability, zero padding,
    value ranges and the dummy-person rule.  (The oracle needs ~30 s per 8 images, so full size is checked this way.)"""
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    m = MultiPersonKeypointModel(cfg.model, cfg.training)
    m.load_state_dict(synth_state_dict(seed=0), strict=True)
    m = m.cuda().eval()
    B, P = 256, 2
    x = torch.from_numpy(synth_images(B, 1, 224, seed=77))
    zero = tuple((b, 1) for b in range(0, B, 3)) + ((5, 0), (5, 1))          # ragged: every third image has one box, image 5 none
    boxes = torch.from_numpy(synth_boxes(B, P, seed=5, zero_slots=zero))
    for mode in ("bf16", "fp32"):
        m.set_precision(mode)
        big = {k: v.cpu() for k, v in m.predict({"image": x, "bboxes": boxes}).items() if torch.is_tensor(v)}
        again = {k: v.cpu() for k, v in m.predict({"image": x, "bboxes": boxes}).items() if torch.is_tensor(v)}
        hm = big["heatmap"]
        assert hm.shape == (B, P, 17, 56, 56) and big["keypoints"].shape == (B, P, 1, 17, 2) and big["visibilities"].shape == (B, P, 1, 17, 3)
        assert torch.isfinite(hm).all() and float(hm.min()) >= 0.0 and float(hm.max()) <= 1.0
        assert float(big["keypoints"].min()) >= 0.0 and float(big["keypoints"].max()) <= 1.0
        # run-to-run: bit-identical (no floating-point atomics anywhere on the default path)
        assert torch.equal(hm, again["heatmap"]) and torch.equal(big["keypoints"], again["keypoints"])
        # padded person slots are zero; image 5 (no valid box) gets the dummy person: vis = [1,0,0] in slot 0
        for b in range(0, B, 3):
            if b != 5:
                assert float(hm[b, 1].abs().max()) == 0.0 and float(big["visibilities"][b, 1].abs().max()) == 0.0
        assert float(hm[5].abs().max()) == 0.0
        assert torch.equal(big["visibilities"][5, 0, 0], torch.tensor([[1.0, 0.0, 0.0]]).repeat(17, 1))
        # one-hot visibilities for every real person
        vis = big["visibilities"][:, :, 0]
        real = vis.sum(-1).sum(-1) > 0
        assert torch.all(vis[real].sum(-1) == 1.0)
        # batch independence: images 8..15 alone give the same heat maps (same kernels, different tiling of the batch)
        sub = m.predict({"image": x[8:16], "bboxes": boxes[8:16]})
        d = float((sub["heatmap"].cpu() - hm[8:16]).abs().max())
        assert d == 0.0, f"{mode}: batch-dependence {d}"
        _report(f"full_size/{mode}/batch_independence_max_abs", d)


# ------------------------------------------------------------------------------------------- benchmarked configurations
def _fresh_model(style, mode):
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    m = MultiPersonKeypointModel(cfg.model, cfg.training, precision=mode)
    sd = synth_state_dict(seed=0, style=style)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _compare_selected_with_oracle(m, sd, x, boxes, sel, mode, tol, relative, tag):
    """Full-size batch on the GPU, oracle on the images `sel` only (images are independent).  The channel order of the
    selected images is teacher-forced from the oracle, every other image keeps the GPU's own top-k."""
    xd, bd = x.cuda(), boxes.cuda()
    m.set_precision(mode)
    m.predict({"image": xd, "bboxes": bd})
    own = m.stage_tensors()["topk"].clone()
    ref = OM.forward(sd, x[sel], boxes[sel], return_stages=True)
    override = own.clone()
    override[torch.as_tensor(sel, device=own.device)] = ref["stages"]["topk"].to(torch.int32).to(own.device)
    out = m.predict({"image": xd, "bboxes": bd}, topk_override=override)
    hm = out["heatmap"][sel].cpu().numpy(); want = ref["heatmap"].numpy()
    if relative:                       # stock weights: point-wise relative error (north-star tolerance)
        live = want > 0
        err = float((np.abs(hm - want)[live] / want[live]).max())
        assert float(np.abs(hm[~live]).max() if (~live).any() else 0.0) == 0.0
    else:                              # calibrated stress weights: of the map's range
        err, _ = heatmap_errors(hm, want)
    kerr = float(np.abs(out["keypoints"][sel].cpu().numpy() - ref["keypoints"].numpy()).max())
    ds = float(np.abs(m.stage_tensors()["scores"][sel].cpu().numpy() - ref["stages"]["scores"].numpy()).max())
    for i in range(4):
        e = rel_range_err(nchw(m.stage_tensors()[f"feat{i}"][sel]), ref["stages"]["feats"][i].numpy())
        assert e < 1e-4, f"{tag}: feat{i} {e}"
    _report(tag, {"heatmap_err": err, "keypoints_abs": kerr, "scores_abs": ds, "images": list(map(int, sel)), "relative": relative})
    assert err < tol, f"{tag}: heat-map error {err} >= {tol}"
    assert kerr < max(1e-5, tol * 1e-1), f"{tag}: keypoint error {kerr}"
    # visibility classes: exact unless sigmoid(max) is within tol of a threshold
    mx = want.reshape(*want.shape[:3], -1).max(-1)
    conf = 1 / (1 + np.exp(-mx))
    near = (np.abs(conf - 0.3) < tol) | (np.abs(conf - 0.7) < tol)
    bad = (out["visibilities"][sel].cpu().numpy() != ref["visibilities"].numpy()).any(-1)[:, :, 0] & ~near
    assert not bad.any(), f"{tag}: visibility class differs away from the thresholds"
    return out


C2_SEL = [0, 1, 2, 3] + list(range(124, 132)) + [252, 253, 254, 255]      # both H2D chunks of the e2e path, tile edges, batch ends


@pytest.mark.parametrize("style,mode,tol", [("stock", "fp32", 1e-3), ("stock", "bf16", 2e-2),
                                            ("calibrated", "fp32", 1e-2), ("calibrated", "bf16", 1e-1)])
def test_bench_config_c2_against_oracle(style, mode, tol):
    """BASELINE config C2 exactly as bench.py runs it (B=256, 224x224, P=1 full-image boxes): 16 images spread over the
    batch against the oracle; and the host-input (two-chunk, two-stream) path gives the same bits as the resident path."""
    m, sd = _fresh_model(style, mode)
    B = 256
    x = torch.from_numpy(synth_images(B, 1, 224, seed=1234))
    boxes = torch.tensor([0.5, 0.5, 1.0, 1.0]).repeat(B, 1, 1)
    out = _compare_selected_with_oracle(m, sd, x, boxes, C2_SEL, mode, tol, style == "stock", f"bench_c2/{style}/{mode}")
    assert out["heatmap"].shape == (B, 1, 17, 56, 56)
    free = m.predict({"image": x.cuda(), "bboxes": boxes.cuda()})
    host = m.predict({"image": x.pin_memory(), "bboxes": boxes.pin_memory()})            # chunked H2D path (images 128..255 = chunk 2)
    assert torch.equal(free["heatmap"], host["heatmap"]) and torch.equal(free["keypoints"], host["keypoints"])
    assert torch.equal(free["visibilities"], host["visibilities"])
    # pipelined submit(): three batches in flight over the two input slots, each equal to the blocking call
    xp, bp = x.pin_memory(), boxes.pin_memory()
    x2 = torch.flip(x, dims=[0]).contiguous().pin_memory()
    pend = [m.submit({"image": xp, "bboxes": bp}), m.submit({"image": x2, "bboxes": bp}), m.submit({"image": xp, "bboxes": bp})]
    res = [p.result() for p in pend]
    assert torch.equal(res[0]["heatmap"], free["heatmap"]) and torch.equal(res[2]["keypoints"], free["keypoints"])
    assert torch.equal(res[1]["heatmap"], torch.flip(free["heatmap"], dims=[0]))     # reversed batch: per-image results, reversed


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-2), ("bf16", 1e-1)])
def test_bench_config_c5_shape_against_oracle(mode, tol):
    """C5-shaped batch: 512x512 images, B=16, P=3 seeded boxes (one padded slot), 6 images against the oracle."""
    m, sd = _fresh_model("calibrated", mode)
    B, P = 16, 3
    x = torch.from_numpy(synth_images(B, 1, 512, seed=4242))
    boxes = torch.from_numpy(synth_boxes(B, P, seed=31, zero_slots=((1, 2), (8, 0))))
    _compare_selected_with_oracle(m, sd, x, boxes, [0, 1, 7, 8, 14, 15], mode, tol, False, f"bench_c5_shape/{mode}")


@pytest.mark.parametrize("mode,tol", [("fp32_strict", 1e-4), ("fp32", 1e-3), ("bf16", 2e-2)])
def test_stock_weights_batch8(mode, tol):
    """North-star tolerances (relative) on reference-style random-init weights at B=8, P=3 -- every image against the oracle."""
    m, sd = _fresh_model("stock", mode)
    x = torch.from_numpy(synth_images(8, 1, 224, seed=99))
    boxes = torch.from_numpy(synth_boxes(8, 3, seed=7, zero_slots=((2, 1),)))
    _compare_selected_with_oracle(m, sd, x, boxes, list(range(8)), mode, tol, True, f"stock_b8/{mode}")


def test_detector_path_new_semantics(model, sd):
    """Bare-tensor forward.  NEW behaviour without a reference (upstream crashes there): the boxes must be
    exactly what oracle.detect derives from the GPU's own head outputs, and the keypoint stages must give what the box path
    gives for those boxes."""
    from oracle.detect import detect_persons
    B, S = 3, 224
    x = torch.from_numpy(synth_images(B, 1, S, seed=321)).cuda()
    model.set_precision("fp32")
    det = model.person_detector
    out = model.predict(x)
    torch.cuda.synchronize()
    P = model.max_detections
    assert out["heatmap"].shape == (B, P, 17, 56, 56) and out["keypoints"].shape == (B, P, 1, 17, 2) and out["visibilities"].shape == (B, P, 1, 17, 3)
    boxes = torch.stack(out["boxes"]).cpu().numpy()
    dbg = model._last_detect
    ref_boxes, ref_count = detect_persons(dbg["box"].cpu().numpy(), dbg["cls"].cpu().numpy(), S // 2, S // 2, S, tuple(det.anchor_sizes),
                                          det.conf_threshold, det.nms_iou_threshold, model.detect_top_m, P)
    assert np.array_equal(dbg["count"].cpu().numpy(), ref_count), (dbg["count"].cpu().numpy(), ref_count)
    assert ref_count.max() >= 1
    np.testing.assert_allclose(boxes, ref_boxes, rtol=0, atol=2e-6)
    # the same boxes through the box path: identical keypoint stages (same kernels, same inputs)
    again = model.predict({"image": x, "bboxes": torch.from_numpy(boxes).cuda()})
    assert torch.equal(again["heatmap"], out["heatmap"]) and torch.equal(again["keypoints"], out["keypoints"])
    assert torch.equal(again["visibilities"], out["visibilities"])
    # nothing passes the confidence threshold -> all-zero boxes -> the reference's dummy person per image (keypoint_model.py:171-179)
    old = det.conf_threshold
    try:
        det.conf_threshold = 1.1
        none = model.predict(x)
        assert float(torch.stack(none["boxes"]).abs().max()) == 0.0
        assert float(none["heatmap"].abs().max()) == 0.0
        assert torch.equal(none["visibilities"][:, 0, 0].cpu(), torch.tensor([[1.0, 0.0, 0.0]]).repeat(B, 17, 1))
    finally:
        det.conf_threshold = old


def test_cuda_graph_capture_is_bit_identical(model):
    """model.capture(): the whole forward replayed from a CUDA graph on new inputs of the captured shape."""
    prev = model.precision
    model.set_precision("bf16")
    x0 = torch.from_numpy(synth_images(8, 1, 224, seed=5)).cuda(); b0 = torch.from_numpy(synth_boxes(8, 2, seed=6)).cuda()
    x1 = torch.from_numpy(synth_images(8, 1, 224, seed=7)).cuda(); b1 = torch.from_numpy(synth_boxes(8, 2, seed=8, zero_slots=((3, 1),))).cuda()
    captured = model.capture({"image": x0, "bboxes": b0})
    for x, b in ((x1, b1), (x0, b0)):
        ref = model.predict({"image": x, "bboxes": b})
        ref = {k: ref[k].clone() for k in ("heatmap", "keypoints", "visibilities")}
        out = captured(x, b)
        torch.cuda.synchronize()
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    model.set_precision(prev)
