"""GPU parity of the RoI input pipeline (givepose_b200/roi.py -> gp_roi_crop) -- index / byte work, so every comparison is
bit-exact: against the golden vectors made by the reference's own functions, against the NumPy oracle at the loader's full
sizes (640x480 frames, 256 / 64 crops), and against live cv2 where it is installed."""
import os

import numpy as np
import pytest
import torch

from oracle import roi as O
from test_roi_oracle import CASES, G, _random_rois

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def roi():
    from givepose_b200 import roi
    return roi


@pytest.mark.parametrize("i", range(len(CASES)))
def test_kernel_equals_reference_golden(roi, i):
    cx, cy, s, S, R, iid = CASES[i]
    out = roi.roi_crops(torch.from_numpy(G["image"]).cuda(), [[cx, cy]], [s], 0, torch.from_numpy(G["inst"]).cuda(), 0, int(iid),
                        img_size=int(S), out_res=int(R))
    for k in ("roi_img", "roi_mask", "roi_coord_2d"):
        assert out[k].dtype == torch.float32 and np.array_equal(out[k][0].cpu().numpy(), G[f"c{i}/{k}"]), k


def test_batch_of_frames_full_size_equals_oracle(roi):
    rng = np.random.default_rng(5)
    M, B, H, W = 3, 96, 480, 640
    imgs = rng.integers(0, 256, (M, H, W, 3), dtype=np.uint8)
    inst = rng.integers(0, 5, (M, H, W), dtype=np.uint8)
    c, s = _random_rois(B, 6)
    iidx = rng.integers(0, M, B)
    iid = rng.integers(-1, 5, B)
    out = roi.roi_crops(torch.from_numpy(imgs).cuda(), c, s, iidx, torch.from_numpy(inst).cuda(), iidx, iid)
    assert out["roi_img"].shape == (B, 3, 256, 256) and out["roi_mask"].shape == (B, 1, 256, 256) and out["roi_coord_2d"].shape == (B, 2, 64, 64)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    for b in range(B):
        want = O.roi_tensors(imgs[iidx[b]], inst[iidx[b]], c[b], s[b], int(iid[b]))
        for k, w in zip(("roi_img", "roi_mask", "roi_coord_2d"), want):
            assert np.array_equal(got[k][b], w), (b, k)


def test_kernel_equals_live_cv2(roi):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    c, s = _random_rois(64, 9)
    got = roi.roi_crops(torch.from_numpy(img).cuda(), c, s)["roi_img"].cpu().numpy()
    lut = roi.normalisation_table().numpy()
    for b in range(64):
        Mf = cv2.getAffineTransform(*[np.float32(p) for p in O.affine_points(c[b], s[b], 256)])
        crop = cv2.warpAffine(img, Mf, (256, 256), flags=cv2.INTER_NEAREST)
        want = np.stack([lut[ch][crop[..., ch]] for ch in range(3)])
        assert np.array_equal(got[b], want), b


def test_edge_cases_and_errors(roi):
    img = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")
    out = roi.roi_crops(img, np.zeros((0, 2)), np.zeros(0))                        # empty batch
    assert out["roi_img"].shape == (0, 3, 256, 256) and out["roi_coord_2d"].shape == (0, 2, 64, 64) and "roi_mask" not in out
    far = roi.roi_crops(img + 7, [[1e4, -1e4]], [5.0], img_size=16, out_res=4)   # RoI entirely outside: border value 0, normalised
    lut = roi.normalisation_table()
    assert torch.equal(far["roi_img"][0, :, 0, 0].cpu(), lut[:, 0]) and float(far["roi_coord_2d"].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        roi.roi_crops(img.cpu(), [[4, 4]], [4.0])
    with pytest.raises(RuntimeError):
        roi.roi_crops(img.float(), [[4, 4]], [4.0])
    with pytest.raises(RuntimeError):
        roi.roi_crops(img, [[4, 4]], [4.0], image_index=[1])
    with pytest.raises(RuntimeError):
        roi.roi_crops(img, [[4, 4]], [0.0])                                        # singular transform


def test_detections_to_posenet_forward(roi):
    """Frames + detections -> input dict -> PoseNet.forward: same poses as the forward on inputs prepared by the oracle."""
    from test_posenet_gpu import build
    from oracle import posenet as OP
    rng = np.random.default_rng(12)
    B, H, W = 8, 480, 640
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    masks = (rng.random((B, H, W)) > 0.5).astype(np.uint8)
    y1, x1 = rng.integers(0, 300, B), rng.integers(0, 400, B)
    bboxes = np.stack([y1, x1, y1 + rng.integers(20, 180, B), x1 + rng.integers(20, 240, B)], 1)
    syn = OP.make_inputs(B, seed=2)
    data = roi.posenet_inputs_from_detections(torch.from_numpy(frame).cuda(), bboxes, torch.from_numpy(masks).cuda(), syn["cam_K"], syn["mean_size"])
    geo = roi.detection_geometry(bboxes, H, W)
    want = [O.roi_tensors(frame, masks[b], geo["bbox_center"][b], geo["img_scale"][b]) for b in range(B)]
    for j, k in enumerate(("roi_img", "roi_mask", "roi_coord_2d")):
        assert np.array_equal(data[k].cpu().numpy(), np.stack([w[j] for w in want])), k
    _, net = build(OP, "o1", precision="fp32")
    with torch.no_grad():
        out = net(data, "cuda")
        ref_in = dict(data)
        ref_in.update({k: torch.from_numpy(np.stack([w[j] for w in want])) for j, k in enumerate(("roi_img", "roi_mask", "roi_coord_2d"))})
        ref = net(ref_in, "cuda")
    assert out["rot"].shape == (B, 3, 3) and torch.equal(out["rot"], ref["rot"]) and torch.equal(out["trans"], ref["trans"])


def test_full_image_tensor_matches_the_loaders(roi):
    """load_data_eval.py:336-338,348: cv2.resize(frame, (256, 256)) -> normalise -> CHW, repeated per RoI; bit-exact against the
    golden made by cv2, the oracle at the loaders' frame size, and live cv2; without resize it is the plain normalisation."""
    frame = torch.from_numpy(G["full/frame"]).cuda()
    got = roi.full_image_tensor(frame, resize=(64, 48))
    assert got.shape == (1, 3, 48, 64) and np.array_equal(got[0].cpu().numpy(), G["full/full_img"])
    rng = np.random.default_rng(4)
    frames = rng.integers(0, 256, (2, 480, 640, 3), dtype=np.uint8)
    got = roi.full_image_tensor(torch.from_numpy(frames).cuda(), image_index=[1, 0, 1]).cpu().numpy()
    want = np.stack([O.full_img(f) for f in frames])[[1, 0, 1]]
    assert got.shape == (3, 3, 256, 256) and np.array_equal(got, want)
    try:
        import cv2
        for k, f in enumerate(frames):
            ref = ((cv2.resize(f, (256, 256)) / 255.0 - roi.IMG_MEAN) / roi.IMG_STD).transpose(2, 0, 1).astype(np.float32)
            assert np.array_equal(want[[1, 0, 1].index(k)], ref)
    except ImportError:
        pass
    small = rng.integers(0, 256, (2, 48, 64, 3), dtype=np.uint8)
    plain = roi.full_image_tensor(torch.from_numpy(small).cuda(), image_index=[1, 0, 1], resize=None).cpu().numpy()
    want = np.stack([((f / 255.0 - roi.IMG_MEAN) / roi.IMG_STD).transpose(2, 0, 1).astype(np.float32) for f in small])[[1, 0, 1]]
    assert np.array_equal(plain, want)
    with pytest.raises(RuntimeError):
        roi.full_image_tensor(torch.from_numpy(small).cuda())      # 48 x 64 -> 256 x 256 would be an upscale: refused
