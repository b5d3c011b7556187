#!/usr/bin/env python
"""Golden outputs of the reference's own ``network/scale_net.py::Scale_net`` (eval mode, torch.manual_seed(0) initialisation,
pretrained=False) and of the pose assembly of ``evaluation/evaluate.py:114-127``.  Build container only (needs /root/reference)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
import ref_loader  # noqa: E402

ref_loader.install_stubs(lambda *a: None)
try:
    import torch.utils.tensorboard  # noqa: F401
except Exception:
    m = types.ModuleType("torch.utils.tensorboard")
    m.SummaryWriter = object
    sys.modules["torch.utils.tensorboard"] = m
import config.config  # noqa: E402,F401
from absl import flags  # noqa: E402

flags.FLAGS(["x"])
import network.scale_net as S  # noqa: E402

torch.manual_seed(0)
net = S.Scale_net(pretrained=False).eval()
g = torch.Generator().manual_seed(1)
B = 3
data = {"roi_img": torch.randn(B, 3, 96, 96, generator=g), "full_img": torch.randn(B, 3, 64, 80, generator=g),
        "one_hot": torch.eye(6)[torch.tensor([0, 3, 5])], "roi_wh": torch.rand(B, 2, generator=g) * 100 + 50,
        "mean_size": torch.rand(B, 3, generator=g) + 0.1}
with torch.no_grad():
    scale = net(data, "cpu", "test")
# evaluate.py:114-127
rot = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
trans, size = torch.randn(B, 3, generator=g), torch.rand(B, 3, generator=g) + 0.2
pred_RT = torch.zeros([B, 4, 4])
pred_RT[:, :3, :3] = rot
pred_RT[:, :3, 3] = trans
pred_RT[:, 3, 3] = 1
pred_RT[:, :3, :] = pred_RT[:, :3, :] * scale[:, None, None]
out = {k: v.numpy() for k, v in data.items()}
out.update(scale=scale.numpy(), rot=rot.numpy(), trans=trans.numpy(), size=size.numpy(), pred_RT=pred_RT.numpy(),
           pred_size=torch.nn.functional.normalize(size, p=2, dim=1).numpy(), n_keys=np.array(len(net.state_dict())),
           first_conv=net.state_dict()["feat_encoder_bbox.0.0.0.weight"].numpy(), line3=net.state_dict()["line3.weight"].numpy())
np.savez_compressed(os.path.join(HERE, "scalenet.npz"), **out)
print("wrote scalenet.npz", scale)
