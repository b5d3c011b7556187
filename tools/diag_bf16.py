#!/usr/bin/env python
"""Where does the bf16 PoseNet error come from?  (the bf16 bar was 1e-1 / 15 degrees.)

Runs the fp32 parity mode and the bf16 mode of givepose_b200.posenet.PoseNet on the same 64 synthetic RoIs (weights 'o1'),
stage by stage, and prints the max-norm relative error of every intermediate tensor plus the geodesic rotation error; then
cross-feeds (fp32 maps into the bf16 PnP head and bf16 maps into the fp32 head) to attribute the rotation outliers.
Output: gpurun_out/<tag>_diag_bf16.json.  GPU box only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_posenet, posenet_inputs  # noqa: E402
from givepose_b200 import ops  # noqa: E402
from givepose_b200.posenet import _conv1x1_rows  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def stages(net, data, dev):
    cd = torch.bfloat16 if net.cfg.precision == "bf16" else torch.float32
    out = {}
    with torch.no_grad(), net._precision():
        feat = net.backbone(data["roi_img"].to(dev).float().contiguous(), cd)
        f = feat[0].permute(0, 2, 3, 1).contiguous()
        out["backbone"] = f
        out["nocs"] = net.xyz_nocs_head.forward_nhwc(f)
        out["nocs_feat"] = net.nocs_encoder.forward_nhwc(out["nocs"])
        cf = _conv1x1_rows(f, net.feat_reducer)
        out["ivfc"] = net.xyz_deform_head.forward_nhwc(torch.cat([cf, out["nocs_feat"].to(cf.dtype)], -1))
        coord = data["roi_coord_2d"].to(dev).permute(0, 2, 3, 1)
        out["pnp_in"] = torch.cat([out["ivfc"], coord.to(out["ivfc"].dtype)], -1)
        out["rot6"], out["t"], _ = net.pnp_net.forward_nhwc(out["pnp_in"])
    return out


def head(net, pnp_in):
    with torch.no_grad(), net._precision():
        cd = torch.bfloat16 if net.cfg.precision == "bf16" else torch.float32
        r, t, _ = net.pnp_net.forward_nhwc(pnp_in.to(cd))
    return r.float(), t.float()


def rotmat(r6):
    x = torch.nn.functional.normalize(r6[:, 0:3].double(), dim=-1)
    z = torch.nn.functional.normalize(torch.cross(x, r6[:, 3:6].double(), dim=-1), dim=-1)
    return torch.stack((x, torch.cross(z, x, dim=-1), z), dim=-1)


def angles(a6, b6):
    A, B = rotmat(a6), rotmat(b6)
    return torch.rad2deg(torch.acos(((torch.einsum("bij,bij->b", A, B) - 1) / 2).clamp(-1, 1)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "diag"
    dev = torch.device("cuda", 0)
    B = 64
    data = posenet_inputs(B, seed=0)
    _, n32 = build_posenet("fp32", dev)
    _, n16 = build_posenet("bf16", dev)
    s32, s16 = stages(n32, data, dev), stages(n16, data, dev)
    res = {"B": B, "stage_rel_err_bf16_vs_fp32": {k: rel(s16[k].float(), s32[k].float()) for k in s32}}
    a = angles(s16["rot6"].float(), s32["rot6"].float())
    res["rot_deg"] = {"median": a.median().item(), "max": a.max().item(), "sorted_top5": sorted(a.tolist())[-5:]}
    # degeneracy of the 6-D representation per RoI: angle between the two predicted axes (Gram-Schmidt is ill-conditioned near 0/180)
    r = s32["rot6"].double()
    cosab = torch.nn.functional.cosine_similarity(r[:, 0:3], r[:, 3:6], dim=-1)
    res["axis_angle_deg_of_worst_rois"] = [round(float(torch.rad2deg(torch.acos(cosab[i].clamp(-1, 1)))), 2) for i in a.argsort(descending=True)[:5]]
    res["rot6_norms_of_worst_rois"] = [[round(float(r[i, 0:3].norm()), 4), round(float(r[i, 3:6].norm()), 4)] for i in a.argsort(descending=True)[:5]]
    # attribution: which half produces the rotation error
    r_a, t_a = head(n16, s32["pnp_in"])     # fp32 maps -> bf16 head
    r_b, t_b = head(n32, s16["pnp_in"])     # bf16 maps -> fp32 head
    for name, (rr, tt) in (("fp32_maps_into_bf16_head", (r_a, t_a)), ("bf16_maps_into_fp32_head", (r_b, t_b))):
        an = angles(rr, s32["rot6"].float())
        res[name] = {"rot_deg_median": an.median().item(), "rot_deg_max": an.max().item(), "t_rel": rel(tt, s32["t"].float())}
    # the tcgen05 trunk against cuBLAS inside the bf16 head
    n16.pnp_net.tc_linear = False
    r_c, _ = head(n16, s16["pnp_in"])
    n16.pnp_net.tc_linear = True
    an = angles(r_c, s16["rot6"].float())
    res["bf16_head_tcgen05_vs_cublas_rot_deg_max"] = an.max().item()
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"{tag}_diag_bf16.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
