"""CPU: the RoI-pipeline oracle (oracle/roi.py) against the golden vectors made by the reference's own functions + cv2
(tests/golden/roi.npz), against live cv2 where it is installed, and the HOST half of the product path
(gp_roi_affine_inverse, a pure C++ function of the C-ABI library) against the oracle.  Everything is bit-exact."""
import os

import numpy as np
import pytest

from oracle import roi as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "roi.npz"))
CASES = [tuple(r) for r in G["cases"]]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_equals_reference_golden(i):
    cx, cy, s, S, R, iid = CASES[i]
    roi_img, roi_mask, roi_coord = O.roi_tensors(G["image"], G["inst"], np.array([cx, cy]), s, int(iid), int(S), int(R))
    assert np.array_equal(roi_img, G[f"c{i}/roi_img"])
    assert np.array_equal(roi_mask, G[f"c{i}/roi_mask"])
    assert np.array_equal(roi_coord, G[f"c{i}/roi_coord_2d"])
    # the affine itself: cv2.getAffineTransform's bits
    assert np.array_equal(O.get_affine_transform_cv(*O.affine_points([cx, cy], s, int(S))), G[f"c{i}/trans_img"])
    assert np.array_equal(O.get_affine_transform_cv(*O.affine_points([cx, cy], s, int(R))), G[f"c{i}/trans_out"])


def _random_rois(n, seed, W=640, H=480):
    rng = np.random.default_rng(seed)
    c = np.stack([rng.uniform(-60, W + 60, n), rng.uniform(-60, H + 60, n)], 1)
    c[::3] = np.round(c[::3] * 2) / 2                      # bbox centres are multiples of 0.5 in practice
    s = rng.uniform(12, 800, n)
    s[::4] = np.round(s[::4]) * 1.5                        # max(h, w) * DZI_PAD_SCALE
    s[1::16] = float(max(H, W))                            # the clamp of load_data_eval.py:266
    return c, s


def test_host_affine_of_the_library_is_bit_identical_to_the_oracle():
    from givepose_b200 import roi
    c, s = _random_rois(4000, 0)
    for out in (256, 64, 17):
        got = roi.roi_affine_inverse(c, s, out)
        want = np.stack([O.affine_inverse(ci, si, out) for ci, si in zip(c, s)])
        assert got.dtype == np.float64 and np.array_equal(got, want)
    assert roi.roi_affine_inverse(np.zeros((0, 2)), np.zeros(0), 64).shape == (0, 6)
    with pytest.raises(RuntimeError):
        roi.roi_affine_inverse(np.zeros((1, 2)), np.zeros(1), 64)          # scale 0: singular, refused


def test_detection_geometry_matches_the_loader_expressions():
    from givepose_b200 import roi
    rng = np.random.default_rng(3)
    y1, x1 = rng.integers(-5, 300, 50), rng.integers(-5, 400, 50)
    b = np.stack([y1, x1, y1 + rng.integers(2, 300, 50), x1 + rng.integers(2, 400, 50)], 1)
    geo = roi.detection_geometry(b, 480, 640)
    for k, (yy1, xx1, yy2, xx2) in enumerate(b.tolist()):   # load_data_eval.py:258-268, eval_utils.py:243-249
        cx, cy = 0.5 * (xx1 + xx2), 0.5 * (yy1 + yy2)
        sc = min(max(yy2 - yy1, xx2 - xx1) * 1.5, max(480, 640)) * 1.0
        assert geo["bbox_center"][k].tolist() == [cx, cy] and geo["img_scale"][k] == sc and geo["resize_ratio"][k] == 64 / sc
        assert geo["roi_wh"][k].tolist() == [min(640, xx2) - max(0, xx1), min(480, yy2) - max(0, yy1)]


def test_oracle_equals_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    coord = O.get_2d_coord_np(640, 480).transpose(1, 2, 0)
    c, s = _random_rois(120, 2)
    for ci, si in zip(c, s):
        for out, src in ((256, img), (64, coord)):
            Mf = cv2.getAffineTransform(*[np.float32(p) for p in O.affine_points(ci, si, out)])
            assert np.array_equal(Mf, O.get_affine_transform_cv(*O.affine_points(ci, si, out)))
            want = cv2.warpAffine(src, Mf, (out, out), flags=cv2.INTER_NEAREST)
            assert np.array_equal(want, O.crop_resize_nearest(src, ci, si, out))


def test_full_img_resize_oracle_equals_reference_golden_and_live_cv2():
    """cv2.resize(frame, (w, h)) on uint8 (load_data_eval.py:336, FLAGS.resize_full default True) restated in fixed point."""
    assert np.array_equal(O.resize_linear_u8(G["full/frame"], 64, 48), G["full/resized_u8"])
    assert np.array_equal(O.full_img(G["full/frame"], (64, 48)), G["full/full_img"])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for H, W in ((480, 640), (375, 500), (256, 256), (720, 1280), (300, 257)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(O.resize_linear_u8(img, 256, 256), cv2.resize(img, (256, 256))), (H, W)
