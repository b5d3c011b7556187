"""``PoseLoss`` of the training step (reference ``losses/pose_loss.py:14-205``) without per-sample host loops.

The reference walks the batch in Python three times per step -- building ``sym_infos`` (``:56-60``), choosing the closest
symmetric ground-truth rotation per RoI in numpy (``get_closest_rot_batch`` ``:401-428`` -> ``get_closest_rot`` ``:329-352``,
360 candidates each) and, for ``'sym'`` rotation types, masking axes (``:101-104``, ``:165-168``) -- and synchronises on
``sym_mask.sum() > 0`` (``:50``).  Here everything is batched tensor algebra on the device the predictions live on, with no
``.item()`` / ``.cpu()`` on the way, so the step stays asynchronous; the selection is evaluated in float64 like the numpy code.

Same flags (``config/config.py:50-59,101-102,116-117``: ``pose_loss_type='l1'``, ``r_loss='l1'``, ``rot_1_w = tran_w = size_w =
prop_pm_w = 1``, ``coor_w = 0.1``, ``coor_gt_sym='rot'``), same dict keys in and out
(``Rot1, Tran, Size, Point_matching, nocs_coor, sp2d_coor``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class PoseLossConfig:
    pose_loss_type: str = "l1"     # 'l1' | 'smoothl1' (beta 0.5)
    r_loss: str = "l1"             # 'l1' | 'angle'
    r_type: str = "allo_rot6d"
    rot_1_w: float = 1.0
    tran_w: float = 1.0
    size_w: float = 1.0
    prop_pm_w: float = 1.0
    coor_w: float = 0.1
    out_res: int = 64
    threshold: float = 0.03        # PoseLoss.threshold (:27)
    sym_candidates: int = 360      # symmetry_rotation_matrix_y(number=360) (:24)


def symmetry_rotations_y(number: int, dtype=torch.float64) -> torch.Tensor:
    """``symmetry_rotation_matrix_y`` (``:319-326``): rotations about y by 2 pi i / number."""
    th = 2.0 * math.pi / number * torch.arange(number, dtype=torch.float64)
    c, s, z, o = torch.cos(th), torch.sin(th), torch.zeros_like(th), torch.ones_like(th)
    return torch.stack([c, z, s, z, o, z, -s, z, c], dim=1).reshape(number, 3, 3).to(dtype)


def closest_symmetric_rotation(pred_rot: torch.Tensor, gt_rot: torch.Tensor, sym_mask: torch.Tensor, sym_rots: torch.Tensor):
    """``get_closest_rot_batch``: for symmetric RoIs the candidate ``gt @ S_k`` with the smallest rotation error to the
    prediction (``re`` ``:446-461`` is decreasing in ``trace(pred (gt S_k)^T)``, ``k = 0`` is the identity, ties keep the
    first), for the others ``gt`` itself.  float64 like the numpy reference, result in ``gt_rot.dtype``."""
    g = gt_rot.detach().double()
    cand = torch.matmul(g[:, None], sym_rots.to(g.device)[None])                      # (B, K, 3, 3)
    tr = torch.einsum("bij,bkij->bk", pred_rot.detach().double(), cand)
    err = torch.acos(torch.clamp(0.5 * (torch.clamp(tr, max=3.0) - 1.0), -1.0, 1.0))  # the reference compares angles, not traces
    idx = torch.argmin(err, dim=1)
    best = cand[torch.arange(g.shape[0], device=g.device), idx]
    return torch.where(sym_mask.view(-1, 1, 1), best.to(gt_rot.dtype), gt_rot)


class PoseLoss(nn.Module):
    def __init__(self, cfg: PoseLossConfig | None = None):
        super().__init__()
        self.cfg = cfg or PoseLossConfig()
        if self.cfg.pose_loss_type not in ("l1", "smoothl1") or self.cfg.r_loss not in ("l1", "angle"):
            raise NotImplementedError(f"pose_loss_type={self.cfg.pose_loss_type!r} r_loss={self.cfg.r_loss!r}")
        if "sym" in self.cfg.r_type:
            raise NotImplementedError("GIVEPose trains with r_type='allo_rot6d' (config/config.py:116)")
        self.register_buffer("sym_rots", symmetry_rotations_y(self.cfg.sym_candidates), persistent=False)

    def _elem(self, a, b):
        if self.cfg.pose_loss_type == "l1":
            return (a - b).abs()
        return F.smooth_l1_loss(a, b, beta=0.5, reduction="none")

    def coor_loss(self, pred, gt, mask):
        """``cal_coor_loss`` + ``cal_coor_loss_for_batch`` (``:182-204``): masked Huber with threshold 0.03, normalised per
        RoI by the mask area."""
        th = self.cfg.threshold
        diff = (pred * mask - gt * mask).abs()
        m = torch.where(diff > th, diff - th / 2.0, diff.pow(2) / (2.0 * th)) * mask
        return (m.sum(dim=[1, 2, 3]) / (mask.sum(dim=[1, 2, 3]) + 1e-5)).mean()

    def forward(self, pred_dict: Dict[str, torch.Tensor], data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        c = self.cfg
        dev = pred_dict["rot"].device
        to = lambda k: data[k].to(dev, non_blocking=True)
        gt_rot0, gt_t, gt_size = to("rotation"), to("translation"), to("real_size")
        gt_mask, gt_mask_sp, sym = to("roi_mask_output"), to("roi_ivfc_mask_output"), to("sym_info")
        nocs_scale = to("nocs_scale").unsqueeze(-1)
        gt_nocs, gt_ivfc = to("nocs_coord"), to("ivfc_coord")
        bs = gt_rot0.shape[0]
        sym_mask = sym[:, 0] == 1

        # :50-77 -- the reference rotates the coordinate targets of EVERY RoI (identity up to rounding for the asymmetric
        # ones) when the batch holds a symmetric object, and leaves them untouched otherwise; both branches, selected on device
        gt_rot_sym = closest_symmetric_rotation(pred_dict["rot"], gt_rot0, sym_mask, self.sym_rots)
        rot_sym = torch.bmm(gt_rot_sym.transpose(1, 2), gt_rot0)
        any_sym = sym_mask.any()
        rotated = lambda x: torch.bmm(rot_sym, x.reshape(bs, 3, -1)).reshape(bs, 3, c.out_res, c.out_res)
        gt_nocs_s = torch.where(any_sym, rotated(gt_nocs), gt_nocs)
        gt_ivfc_s = torch.where(any_sym, rotated(gt_ivfc), gt_ivfc)
        gt_rot = torch.where(any_sym, gt_rot_sym, gt_rot0)

        out = {}
        if c.r_loss == "l1":
            out["Rot1"] = c.rot_1_w * self._elem(pred_dict["rot"], gt_rot).mean()
        else:   # cal_loss_Rot_angle :110-115
            tr = torch.diagonal(torch.bmm(gt_rot, pred_dict["rot"].transpose(1, 2)), dim1=-2, dim2=-1).sum(-1)
            ang = torch.acos(torch.clip((tr - 1) / 2, -0.99999, 0.99999))
            out["Rot1"] = c.rot_1_w * F.smooth_l1_loss(ang, torch.zeros_like(ang), beta=0.2, reduction="none").mean()
        out["Tran"] = c.tran_w * self._elem(pred_dict["trans"], gt_t / nocs_scale).mean()
        out["Size"] = c.size_w * self._elem(pred_dict["size"], gt_size / nocs_scale).mean()
        pts = to("model_point").permute(0, 2, 1)                                        # :160-171 (translations are commented out there)
        out["Point_matching"] = c.prop_pm_w * self._elem(torch.bmm(pred_dict["rot"], pts), torch.bmm(gt_rot, pts)).mean()
        out["nocs_coor"] = c.coor_w * self.coor_loss(pred_dict["nocs_coor"], gt_nocs_s, gt_mask)
        out["sp2d_coor"] = c.coor_w * self.coor_loss(pred_dict["ivfc_coor"], gt_ivfc_s, gt_mask_sp)
        return out


def make_loss_inputs(B: int, seed: int = 0, sym_every: int = 3) -> Dict[str, torch.Tensor]:
    """Synthetic ground truth in the layout of ``datasets/load_data_nocs.py:355-386``; every ``sym_every``-th RoI is a
    symmetric object (``sym_info[:, 0] == 1``; 0 disables symmetry)."""
    g = torch.Generator().manual_seed(3000 + seed)
    r = lambda *s: torch.randn(*s, generator=g)
    q, _ = torch.linalg.qr(r(B, 3, 3))
    q = q * torch.sign(torch.linalg.det(q)).view(-1, 1, 1)                              # proper rotations
    sym = torch.zeros(B, 4)
    if sym_every:
        sym[::sym_every, 0] = 1
    return {"rotation": q.contiguous(), "translation": r(B, 3) * 0.3 + torch.tensor([0.0, 0.0, 1.0]),
            "real_size": torch.rand(B, 3, generator=g) * 0.3 + 0.1, "nocs_scale": torch.rand(B, generator=g) * 0.5 + 0.2,
            "roi_mask_output": (torch.rand(B, 1, 64, 64, generator=g) > 0.4).float(),
            "roi_ivfc_mask_output": (torch.rand(B, 1, 64, 64, generator=g) > 0.5).float(),
            "sym_info": sym, "nocs_coord": torch.rand(B, 3, 64, 64, generator=g) - 0.5,
            "ivfc_coord": torch.rand(B, 3, 64, 64, generator=g) - 0.5, "model_point": r(B, 1024, 3) * 0.2}
