"""``Scale_net`` (``network/scale_net.py:22-65``) and the pose assembly that follows ``PoseNet.forward`` at test time
(``evaluation/evaluate.py:111-127``) -- the step AFTER the hot path.

``Scale_net`` is a composition of library blocks (two torchvision MobileNetV3-small trunks + three Linear layers); it is mirrored
here so that a reference scale checkpoint loads with ``strict=True`` (identical state-dict keys) and the whole test-time
pipeline -- RoI crops (``givepose_b200.roi``), ``Scale_net``, ``PoseNet.forward``, ``assemble_pred_RT`` -- runs on the device
without the reference's host round trips.  There is no hand-written kernel in it: cuDNN/cuBLAS through torch, channels_last.
torchvision's pretrained weights need the network: ``pretrained=True`` raises unless they are in the local torch hub cache.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Scale_net(nn.Module):
    def __init__(self, feat_dim=8, use_hw=True, backbone="mobilenetv3s", pretrained=False, cats_num=6):
        super().__init__()
        try:
            import torchvision
        except ImportError as e:   # pragma: no cover
            raise RuntimeError("givepose_b200.Scale_net needs torchvision (MobileNetV3-small trunks)") from e
        weights = "IMAGENET1K_V1" if pretrained else None
        bbox = torchvision.models.mobilenet_v3_small(weights=weights)      # construction order == reference (:25-26)
        full = torchvision.models.mobilenet_v3_small(weights=weights)
        self.feat_encoder_bbox = nn.Sequential(bbox.features, bbox.avgpool, nn.Flatten())
        self.feat_encoder_full = nn.Sequential(full.features, full.avgpool, nn.Flatten())
        in_dim = bbox.features[-1].out_channels * 2
        self.drop = nn.Dropout(p=0.2, inplace=True)
        self.line1 = nn.Linear(in_dim, 128)
        self.line2 = nn.Linear(128 + cats_num, feat_dim)
        self.relu = nn.ReLU(inplace=True)
        self.use_hw = use_hw
        if use_hw:
            feat_dim += 2
        self.line3 = nn.Linear(feat_dim + cats_num, 1)
        self.head = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten())   # built but unused in the reference (:43-44)

    def forward(self, data, device, mode=""):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("givepose_b200.Scale_net: Not implemented on the CPU (there is no CPU fallback)")
        cl = lambda t: t.to(dev, non_blocking=True).contiguous(memory_format=torch.channels_last)
        one_hot = data["one_hot"].to(dev, non_blocking=True)
        feat_roi = self.drop(self.feat_encoder_bbox(cl(data["roi_img"])))
        feat_full = self.drop(self.feat_encoder_full(cl(data["full_img"])))
        x = self.relu(self.line1(torch.cat([feat_roi, feat_full], dim=1)))
        x = self.relu(self.line2(torch.cat([x, one_hot], dim=1)))
        x = torch.cat([x, one_hot], dim=1)
        if self.use_hw:
            x = torch.cat([x, data["roi_wh"].to(dev, non_blocking=True) / 100], dim=1)
        resi_scale = self.line3(x).squeeze()
        return resi_scale + data["mean_size"].to(dev, non_blocking=True).norm(dim=1)


def assemble_pred_RT(rot: torch.Tensor, trans: torch.Tensor, size: torch.Tensor, pred_scale: torch.Tensor):
    """``evaluate.py:114-127`` batched on the device: ``pred_RT`` (B,4,4) with the scaled rotation / translation rows and the
    L2-normalised size (``pred_scales`` of the detection dict).  Inputs may live on different devices (``rot`` is on the host in
    the reference's test path); everything is moved to ``trans``'s device."""
    dev = trans.device
    bs = rot.shape[0]
    RT = torch.zeros(bs, 4, 4, dtype=torch.float32, device=dev)
    RT[:, :3, :3] = rot.to(dev)
    RT[:, :3, 3] = trans
    RT[:, 3, 3] = 1
    RT[:, :3, :] = RT[:, :3, :] * pred_scale.to(dev).reshape(bs, 1, 1)
    return RT, torch.nn.functional.normalize(size.to(dev), p=2, dim=1)
