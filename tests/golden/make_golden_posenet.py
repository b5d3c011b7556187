"""Generate the PoseNet golden vectors from the REFERENCE's own ``network/PoseNet.py``.

Run once in the build container (``/root/reference`` exists only there):

    python tests/golden/make_golden_posenet.py

``ref_loader.py`` injects leaf stubs for the third-party modules the reference imports but this image lacks
(mmcv, timm, detectron2, transforms3d, open3d, ... -- SURVEY.md Appendix B) and imports ``network.PoseNet``
unchanged.  Two things are substituted, both documented in DESIGN.md:

* the compiled ``DCNv3`` extension (CUDA-only, ``src/dcnv3.h:37``) is replaced by the reference's own
  ``dcnv3_core_pytorch`` behind the flat-slice adapter of SURVEY.md 0.1 (what the CUDA kernel computes for the
  stride-2 in-model calls);
* ``convnext_backbone`` (timm + pretrained download) is replaced by the reference's own ``network/resnet.py``
  ResNet-34 trunk plus a 1x1 neck to the hard-coded 1024 channels (``PoseNet.py:144``).

Weights come from ``oracle.posenet.init_weights`` (seeded), loaded into the reference module with
``strict=True`` -- which also proves state-dict compatibility.  Inputs come from ``oracle.posenet.make_inputs``.
Stored: the reference outputs (small), plus SHA-256 of inputs and weights so a test can tell "RNG drifted"
from "result differs".
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from oracle import posenet as OP  # noqa: E402
from oracle.dcnv3 import flat_slice, out_size  # noqa: E402

_ref_func = {}


def dcnv3_forward_stub(input, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, scale, im2col_step, rc=0):
    """What the CUDA extension computes, expressed with the reference's own PyTorch path."""
    N, H, W, _ = input.shape
    Ho, Wo = out_size(H, kh, sh, ph, dh), out_size(W, kw, sw, pw, dw)
    return _ref_func["core"](input, flat_slice(offset, N, Ho, Wo), flat_slice(mask, N, Ho, Wo), kh, kw, sh, sw, ph, pw,
                             dh, dw, group, gc, scale, rc)


def sha(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    ref_loader.install_stubs(dcnv3_forward_stub)
    PN = ref_loader.load_posenet_module()
    from network.ops_dcnv3.functions import dcnv3_func
    _ref_func["core"] = dcnv3_func.dcnv3_core_pytorch
    from network.resnet import resnet34

    class RefBackbone(nn.Module):   # reference ResNet-34 trunk (network/resnet.py) + neck, emitting [B,1024,8,8]
        def __init__(self):
            super().__init__()
            r = resnet34()
            del r.fc, r.avgpool
            self.trunk, self.neck = r, nn.Conv2d(512, 1024, 1)

        def forward(self, x):
            t = self.trunk
            x = t.maxpool(t.relu(t.bn1(t.conv1(x))))
            return [self.neck(t.layer4(t.layer3(t.layer2(t.layer1(x)))))]

    PN.convnext_backbone = lambda: RefBackbone()
    torch.manual_seed(0)
    ref = PN.PoseNet().eval()

    blob, B = {}, 8
    for mode in ("reference", "o1"):
        mine = OP.PoseNet().eval()
        OP.init_weights(mine, mode, seed=0)
        sd = mine.state_dict()
        missing = ref.load_state_dict(sd, strict=True)   # key / shape compatibility with the reference module
        data = OP.make_inputs(B, seed=0)
        with torch.no_grad():
            out_ref = ref.forward({k: v.clone() for k, v in data.items()}, "cpu")
            out_mine = mine.forward(data, "cpu")
        for k in ("rot", "trans", "size", "mask", "nocs_coor", "ivfc_coor"):
            a, b = out_ref[k].double(), out_mine[k].double()
            rel = ((a - b).abs().max() / a.abs().max().clamp_min(1e-30)).item()
            print(f"[{mode}] {k:10s} ref absmax {a.abs().max().item():.3e}  oracle-vs-ref rel {rel:.2e}")
            if k != "mask":
                blob[f"{mode}/{k}"] = out_ref[k].numpy()
        if mode == "o1":   # training-path pose decode (do_loss=True: pose_from_predictions_train + allo_to_ego_mat_torch)
            dtrain = {k: v.clone() for k, v in data.items()}
            dtrain["roi_mask_deform"] = dtrain["roi_mask"].clone()
            with torch.no_grad():
                tr_ref = ref.forward({k: v.clone() for k, v in dtrain.items()}, "cpu", do_loss=True)
                tr_mine = mine.forward(dtrain, "cpu", do_loss=True)
            for k in ("rot", "trans"):
                a, b = tr_ref[k].double(), tr_mine[k].double()
                print(f"[o1 train path] {k:6s} oracle-vs-ref rel {((a - b).abs().max() / a.abs().max()).item():.2e}")
                blob[f"o1_train/{k}"] = tr_ref[k].numpy()
        blob[f"{mode}/sha_inputs"] = np.array(sha(data[k] for k in sorted(data)))
        blob[f"{mode}/sha_weights"] = np.array(sha(sd[k] for k in sorted(sd)))
    blob["__provenance__"] = np.array(["reference network/PoseNet.py:173-231 forward (CPU, eval), DCNv3 core = reference "
                                       "dcnv3_core_pytorch + flat-slice adapter, backbone = reference resnet34 trunk + neck; "
                                       f"B={B}; torch {torch.__version__}"])
    path = os.path.join(HERE, "posenet.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
