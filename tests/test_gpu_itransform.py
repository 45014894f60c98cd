"""GPU parity of `kpd_itransform` through `dll.data.ITransform(device="cuda")` -> ctypes -> C-ABI, against
the CPU oracle (oracle/itransform.py, itself pinned bit-exactly on OpenCV / Pillow and on the reference's own outputs) and
against the reference-generated vectors in tests/golden/itransform.npz.  Every stage is byte work: the bar is bit-exact."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

from oracle import itransform as O


def _cases():
    spec = importlib.util.spec_from_file_location("mk_it", os.path.join(GOLDEN, "make_itransform_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES


CASES = _cases()


def _t(size, clip, tiles, gray):
    from dll.data import ITransform
    return ITransform(img_size=size, clip_limit=clip, tile_size=tiles, grayscale=gray, device="cuda")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_matches_reference_vectors(case):
    name, h, w, seed, ch, gray, size, clip, tiles = case
    gold = np.load(os.path.join(GOLDEN, "itransform.npz"))
    img = O.synth_photo(h, w, seed, ch)
    out = _t(size, clip, tiles, gray)(img)                                     # the single-image call a reference user makes
    torch.cuda.synchronize()
    assert out.is_cuda and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), gold[name + ".final"])           # float32 of exact bytes: bit-exact


@pytest.mark.parametrize("hw", [(240, 320), (227, 301), (200, 333), (64, 64), (97, 131), (33, 500)])
def test_every_stage_against_oracle(hw):
    h, w = hw
    imgs = np.stack([O.synth_photo(h, w, 20 + i) for i in range(3)])
    rng = np.random.default_rng(h * w)
    imgs[2] = rng.integers(0, 256, imgs[2].shape, dtype=np.uint8)             # pure noise: dense edge map, long hysteresis chains
    out, st = _t(224, 2.0, (8, 8), True).batch(torch.from_numpy(imgs), return_stages=True)
    torch.cuda.synchronize()
    for b in range(3):
        ref = {}
        mixed = O.to_grayscale_clahe(imgs[b], 2.0, (8, 8), stages=ref)
        for key, okey in (("plane", "gray"), ("clahe", "clahe"), ("edges", "edges"), ("soft", "soft")):
            got = st[key][b, 0].cpu().numpy()
            bad = int((got != ref[okey]).sum())
            assert bad == 0, (hw, b, key, bad)
        bounds, kk = O.pil_bilinear_coeffs(w, 224)
        hp = O.pil_resize_bilinear(mixed, h, 224) if w != 224 else mixed
        assert np.array_equal(st["hpass"][b, 0].cpu().numpy(), hp), (hw, b, "hpass")
        assert np.array_equal(out[b].cpu().numpy(), O.to_tensor_normalize(O.pil_resize_bilinear(mixed, 224, 224))), (hw, b)


def test_fine_clahe_grid():
    """16x16 CLAHE tiles of 10x13 pixels: a 76-pixel Canny region blends between 8 x 7 tile LUTs."""
    img = O.synth_photo(160, 208, 77)
    out, st = _t(224, 3.0, (16, 16), True).batch(torch.from_numpy(img[None]), return_stages=True)
    ref = {}
    mixed = O.to_grayscale_clahe(img, 3.0, (16, 16), stages=ref)
    assert np.array_equal(st["clahe"][0, 0].cpu().numpy(), ref["clahe"]) and np.array_equal(st["edges"][0, 0].cpu().numpy(), ref["edges"])
    assert np.array_equal(out[0].cpu().numpy(), O.to_tensor_normalize(O.pil_resize_bilinear(mixed, 224, 224)))


def test_random_geometries():
    """Seeded sweep over image sizes (odd, prime, barely above the 16-pixel minimum, full HD), CLAHE grids that do / do not divide
    them, clip limits (0 = plain histogram equalisation) and output sizes (down- and up-sampling): final tensor bit-exact."""
    rng = np.random.default_rng(2024)
    cases = [(16, 16, (2, 2), 1.5, 32), (17, 401, (3, 8), 2.0, 224), (1080, 1920, (8, 8), 2.0, 224), (611, 389, (7, 5), 0.0, 96),
             (224, 224, (8, 8), 1.5, 224), (100, 100, (8, 8), 40.0, 300)]
    for _ in range(6):
        cases.append((int(rng.integers(20, 400)), int(rng.integers(20, 400)), (int(rng.integers(1, 9)), int(rng.integers(1, 9))),
                      float(rng.choice([0.0, 1.0, 1.5, 2.0, 4.0])), int(rng.choice([64, 128, 224]))))
    for k, (h, w, tiles, clip, size) in enumerate(cases):
        img = O.synth_photo(h, w, 300 + k) if k % 3 else rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = _t(size, clip, tiles, True)(img).cpu().numpy()
        assert np.array_equal(got, O.itransform(img, size, clip, tiles)), (h, w, tiles, clip, size)
    for k, (h, w, tiles, clip, size) in enumerate(cases[:2] + cases[3:8]):
        img = O.synth_photo(h, w, 400 + k)
        got = _t(size, clip, tiles, False)(img).cpu().numpy()
        assert np.array_equal(got, O.itransform(img, size, clip, tiles, grayscale=False)), ("rgb", h, w, tiles, clip, size)


def test_rgb_pipeline_and_input_forms():
    img = O.synth_photo(150, 210, 31)
    t = _t(96, 1.5, (8, 8), False)
    out, st = t.batch(torch.from_numpy(img[None]), return_stages=True)
    ref = O.to_rgb_clahe(img, 1.5, (8, 8))
    assert np.array_equal(st["clahe"][0].permute(1, 2, 0).cpu().numpy(), ref)
    assert np.array_equal(out[0].cpu().numpy(), O.itransform(img, 96, 1.5, (8, 8), grayscale=False))
    gray = O.rgb_to_gray(img)                                                 # 2-D input -> replicated channels (transforms.py:87-89)
    assert np.array_equal(t(gray).cpu().numpy(), O.itransform(gray, 96, 1.5, (8, 8), grayscale=False))
    rgba = O.synth_photo(150, 210, 31, channels=4)
    assert np.array_equal(_t(224, 1.5, (8, 8), True)(rgba).cpu().numpy(), O.itransform(rgba, 224, 1.5, (8, 8)))
    from PIL import Image
    assert np.array_equal(_t(224, 1.5, (8, 8), True)(Image.fromarray(img)).cpu().numpy(), O.itransform(img, 224, 1.5, (8, 8)))
    with pytest.raises(TypeError):
        t.batch(torch.zeros(1, 32, 32, 3))
    with pytest.raises(RuntimeError):
        t.batch(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))                   # below the 16-pixel minimum: argument error from the library


def test_batch_independence_and_determinism_at_bench_size():
    """B=64 640x480 (the bench shape): images 0, 31, 63 against the oracle, image 5 equal to its single-image result,
    two runs bit-identical (integer atomics only)."""
    imgs = np.stack([O.synth_photo(480, 640, 100 + (i % 8)) for i in range(64)])
    x = torch.from_numpy(imgs).cuda()
    t = _t(224, 2.0, (8, 8), True)
    a = t.batch(x).clone()
    b = t.batch(x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    for i in (0, 31, 63):
        assert np.array_equal(a[i].cpu().numpy(), O.itransform(imgs[i], 224, 2.0, (8, 8))), i
    assert torch.equal(t.batch(x[5:6])[0], a[5])
    assert torch.equal(a[5], a[13])                                            # same source image in another batch slot


def test_pinned_host_batch_is_chunk_pipelined_and_identical():
    imgs = torch.from_numpy(np.stack([O.synth_photo(120, 160, 60 + i) for i in range(10)]))
    t = _t(224, 2.0, (8, 8), True)
    resident = t.batch(imgs.cuda())
    for _ in range(2):                                                         # second call reuses the staging buffer
        piped = t.batch(imgs.pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(resident, piped)
    assert torch.equal(resident, t.batch(imgs))                                # pageable host tensor: plain copy path
    # explicit tickets: two batches in flight (both staging slots), consumed in order
    other = torch.from_numpy(np.stack([O.synth_photo(120, 160, 80 + i) for i in range(10)])).pin_memory()
    t1, t2 = t.prefetch(imgs.pin_memory()), t.prefetch(other)
    a, b = t.batch(t1), t.batch(t2)
    t3 = t.prefetch(imgs.pin_memory())                                         # re-uses the first slot after its consumer
    c = t.batch(t3)
    torch.cuda.synchronize()
    assert torch.equal(a, resident) and torch.equal(c, resident) and torch.equal(b, t.batch(other.cuda()))
    with pytest.raises(ValueError):
        t.prefetch(imgs)                                                       # not pinned


def test_feeds_the_model():
    """ITransform(device='cuda') output goes straight into MultiPersonKeypointModel.predict (the predict.py path)."""
    from dll import MultiPersonKeypointModel
    from dll.configs.config_loader import load_config
    from oracle.weights import synth_state_dict
    cfg = load_config(os.path.join(ROOT, "configs", "default_config.yaml"))
    model = MultiPersonKeypointModel(cfg.model, cfg.training).cuda().eval()
    model.load_state_dict(synth_state_dict(0), strict=True)
    imgs = torch.from_numpy(np.stack([O.synth_photo(240, 320, 40 + i) for i in range(2)]))
    x = _t(224, 2.0, (8, 8), True).batch(imgs)
    out = model.predict({"image": x, "bboxes": torch.tensor([[[0.5, 0.5, 1.0, 1.0]]] * 2)})
    assert out["keypoints"].shape == (2, 1, 1, 17, 2) and torch.isfinite(out["heatmap"]).all()
