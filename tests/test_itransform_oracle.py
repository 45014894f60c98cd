"""CPU tests of the ITransform oracle (oracle/itransform.py): pinned on vectors produced by the reference's own class
(tests/golden/itransform.npz, made by tests/golden/make_itransform_golden.py) and, where OpenCV / Pillow are importable,
stage by stage and bit-exactly on the libraries themselves."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import itransform as O                                             # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "itransform.npz")


def _cases():
    spec = importlib.util.spec_from_file_location("mk_it", os.path.join(ROOT, "tests", "golden", "make_itransform_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_vectors(case):
    name, h, w, seed, ch, gray, size, clip, tiles = case
    gold = np.load(GOLD)
    img = O.synth_photo(h, w, seed, ch)
    staged = O.to_grayscale_clahe(img, clip, tiles) if gray else O.to_rgb_clahe(img, clip, tiles)
    assert np.array_equal(staged, gold[name + ".staged"])                      # bytes: bit-exact
    final = O.itransform(img, size, clip, tiles, gray)
    assert final.dtype == np.float32 and np.array_equal(final, gold[name + ".final"])


def test_median_and_hysteresis_definitions():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (40, 50), dtype=np.uint8)
    m = O.median5(a)
    p = np.pad(a, 2, mode="edge")
    assert m[0, 0] == np.sort(p[0:5, 0:5].ravel())[12] and m[20, 30] == np.sort(p[20:25, 30:35].ravel())[12]
    pm = np.ones((8, 12), np.uint8)
    pm[2, 1:6] = 0; pm[2, 6] = 2; pm[5, 8:11] = 0; pm[3, 7] = 0          # chain touching a strong pixel (diagonal too), one orphan chain
    e = O.hysteresis(pm)
    assert e[2, 1] == 255 and e[3, 7] == 255 and e[2, 6] == 255 and e[5, 9] == 0 and e[0, 0] == 0


def test_stages_bit_exact_against_the_libraries():
    cv2 = pytest.importorskip("cv2")
    Image = pytest.importorskip("PIL.Image")
    for (h, w, seed) in [(120, 160, 11), (97, 131, 12)]:
        img = O.synth_photo(h, w, seed)
        gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
        assert np.array_equal(O.rgb_to_gray(img), gray)
        for clip, tiles in [(1.5, (8, 8)), (2.0, (5, 5)), (0.0, (4, 4))]:
            assert np.array_equal(O.clahe_apply(gray, clip, tiles), cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles).apply(gray))
        eq = O.clahe_apply(gray, 1.5, (8, 8))
        g5 = cv2.GaussianBlur(eq, (5, 5), 1.5)
        assert np.array_equal(O.gauss_fixed(eq, O.GAUSS5_S15), g5)
        md = cv2.medianBlur(g5, 5)
        assert np.array_equal(O.median5(g5), md)
        ed = cv2.Canny(md, 100, 200)
        assert ed.max() == 255 and np.array_equal(O.canny(md), ed)
        k = np.ones((3, 3), np.uint8)
        assert np.array_equal(O.dilate3(ed), cv2.dilate(ed, k)) and np.array_equal(O.erode3(cv2.dilate(ed, k)), cv2.erode(cv2.dilate(ed, k), k))
        assert np.array_equal(O.gauss_fixed(ed, O.GAUSS3_S0), cv2.GaussianBlur(ed, (3, 3), 0))
        assert np.array_equal(O.gauss_fixed(gray, O.GAUSS3_S05), cv2.GaussianBlur(gray, (3, 3), 0.5))
        assert np.array_equal(O.pil_resize_bilinear(eq, 224, 224), np.asarray(Image.fromarray(eq).resize((224, 224), Image.BILINEAR)))
        assert np.array_equal(O.pil_resize_bilinear(img, 50, 70), np.asarray(Image.fromarray(img).resize((70, 50), Image.BILINEAR)))
    # addWeighted over all 65,536 byte pairs (the fused multiply-add decides 128 of them)
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8))
    assert np.array_equal(O.add_weighted(a, b), cv2.addWeighted(a, 0.7, b, 0.3, 0))


def test_median_network_of_the_cuda_kernel_is_a_median():
    """Zero-one principle over all 2^25 binary inputs for the 99-exchange network compiled into csrc/kpd_itransform.cu."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_median_network.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
