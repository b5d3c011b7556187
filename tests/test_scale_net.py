"""Scale_net mirror (givepose_b200/scale_net.py) and the test-time pose assembly against golden outputs of the reference's own
``network/scale_net.py::Scale_net`` and ``evaluation/evaluate.py:114-127`` (tests/golden/make_golden_scalenet.py)."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scalenet.npz"))


def _net():
    from givepose_b200.scale_net import Scale_net
    torch.manual_seed(0)
    return Scale_net(pretrained=False).eval()


def test_state_dict_and_initialisation_match_the_reference_class():
    net = _net()
    sd = net.state_dict()
    assert len(sd) == int(G["n_keys"])                      # same keys -> a reference scale checkpoint loads with strict=True
    assert np.array_equal(sd["feat_encoder_bbox.0.0.0.weight"].numpy(), G["first_conv"])   # same construction order / RNG stream
    assert np.array_equal(sd["line3.weight"].numpy(), G["line3"])
    with pytest.raises(RuntimeError):
        net({k: torch.from_numpy(G[k]) for k in ("roi_img", "full_img", "one_hot", "roi_wh", "mean_size")}, "cpu")


@pytest.mark.gpu
def test_forward_and_pose_assembly_match_the_reference_golden():
    from givepose_b200.scale_net import assemble_pred_RT
    net = _net().cuda()
    data = {k: torch.from_numpy(G[k]) for k in ("roi_img", "full_img", "one_hot", "roi_wh", "mean_size")}
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            scale = net(data, "cuda", "test")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert scale.shape == (3,) and np.allclose(scale.cpu().numpy(), G["scale"], rtol=1e-4, atol=1e-5)
    RT, size = assemble_pred_RT(torch.from_numpy(G["rot"]), torch.from_numpy(G["trans"]).cuda(), torch.from_numpy(G["size"]).cuda(),
                                torch.from_numpy(G["scale"]))
    assert np.array_equal(RT.cpu().numpy(), G["pred_RT"]) and np.allclose(size.cpu().numpy(), G["pred_size"], rtol=1e-6)


@pytest.mark.gpu
def test_test_time_pipeline_frames_to_scaled_poses():
    """evaluate.py:100-127 end to end on the device: frames + detections -> RoI crops + full_img -> Scale_net -> PoseNet.forward
    -> pred_RT / pred_size; checks the plumbing (shapes, devices, finiteness) -- each stage has its own parity test."""
    from givepose_b200 import roi
    from givepose_b200.posenet import PoseNet, PoseNetConfig
    from givepose_b200.scale_net import assemble_pred_RT
    rng = np.random.default_rng(0)
    B, H, W = 5, 480, 640
    frames = torch.from_numpy(rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)).cuda()
    masks = torch.from_numpy((rng.random((B, H, W)) > 0.5).astype(np.uint8)).cuda()
    y1, x1 = rng.integers(0, 300, B), rng.integers(0, 400, B)
    bboxes = np.stack([y1, x1, y1 + rng.integers(30, 170, B), x1 + rng.integers(30, 230, B)], 1)
    frame_of_roi = [0, 0, 1, 1, 1]
    cam_K = torch.tensor([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])
    data = roi.posenet_inputs_from_detections(frames, bboxes, masks, cam_K, torch.rand(B, 3) + 0.1, image_index=frame_of_roi)
    data["full_img"] = roi.full_image_tensor(frames, frame_of_roi)
    data["one_hot"] = torch.eye(6)[torch.tensor([0, 1, 2, 3, 5])]
    torch.manual_seed(0)
    net = PoseNet(PoseNetConfig(precision="bf16")).eval().cuda()
    with torch.no_grad():
        scale = _net().cuda()(data, "cuda", "test")
        out = net(data, "cuda", pred_scale=scale.cpu())
        RT, size = assemble_pred_RT(out["rot"], out["trans"], out["size"], scale)
    assert data["full_img"].shape == (B, 3, 256, 256) and scale.shape == (B,)
    assert RT.shape == (B, 4, 4) and RT.is_cuda and size.shape == (B, 3) and bool(torch.isfinite(RT).all())
    assert torch.allclose(RT[:, 3], torch.tensor([0.0, 0, 0, 1], device="cuda").expand(B, 4))
