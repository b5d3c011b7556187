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
