"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch CPU restatement of the reference ``PoseNet.forward``.

Never imported by ``givepose_b200/``; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it (see ``oracle/__init__.py``).

What is restated (reference file:line, relative to /root/reference):

* ``PoseNet.__init__/forward``               network/PoseNet.py:134-231
* ``TopDownXyzHead``                         network/xyz_head.py:195-366
* ``ConvModule`` (conv -> GN -> act)         network/torch_utils/layers/conv_module.py:57-234
* ``MAPEncoder`` / ``ConvPnPNet``            network/conv_pnp_net.py:203-332 / :18-201
* ``DCNv3_C`` / ``DCNv3``                    network/dcnv3.py:23-38 / network/ops_dcnv3/modules/dcnv3.py:221-356
* ``SizeHead``                               network/pose_head.py:17-51
* ``rot6d_to_mat_batch``                     network/pose_utils/rot_reps.py:34-55
* ``pose_from_predictions_test``             network/pose_utils/pose_from_pred_centroid_z.py:59-157
* ``allocentric_to_egocentric`` (mat->mat)   network/pose_utils/utils.py:29-84  (float64 numpy, per RoI, like the reference)
* ``axangle2mat``                            transforms3d 0.4.1 (GIVEPose_env.yml:185), published Rodrigues formula
* backbone for the synthetic runs            network/resnet.py:96-212 (ResNet-34 trunk), + a 1x1 ``neck`` 512 -> 1024
                                             (the reference hard-codes feature_channel = 1024, PoseNet.py:144)

State-dict keys equal the reference's (``tests/golden/make_golden_posenet.py`` loads this module's weights into
the reference ``PoseNet`` with ``strict=True`` before producing the golden outputs).

The DCNv3 core inside ``DCNv3`` is ``oracle.dcnv3.forward``: the C restatement of the reference CUDA kernel, which
reads ``offset`` / ``mask`` through their flat ``[N*Ho*Wo]``-row prefix -- the stride-2 behaviour of the reference
extension.  Pinning: tests/golden/posenet.npz (reference ``PoseNet`` run in the build
container with the core = reference ``dcnv3_core_pytorch`` + flat-slice adapter).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dcnv3 as core

CAM_K = [[591.0125, 0.0, 322.525], [0.0, 590.16775, 244.11084], [0.0, 0.0, 1.0]]   # NOCS Real intrinsics


# ------------------------------------------------------------------------------------------------------
# DCNv3 module (modules/dcnv3.py:221-356) and its NCHW wrapper (network/dcnv3.py:23-38)
# ------------------------------------------------------------------------------------------------------
class _LNChannelsLast(nn.Sequential):
    """``build_norm_layer(dim, 'LN', 'channels_first', 'channels_last')`` (modules/dcnv3.py:37-58): index 0 is the
    NCHW->NHWC permute, index 1 the LayerNorm -- hence the key ``dw_conv.1.1.weight``."""

    class _ToLast(nn.Module):
        def forward(self, x):
            return x.permute(0, 2, 3, 1)

    def __init__(self, dim):
        super().__init__(self._ToLast(), nn.LayerNorm(dim, eps=1e-6))


class DCNv3(nn.Module):
    differentiable = False   # True: core = C oracle forward + backward behind autograd (CoreFunction)

    def __init__(self, channels, kernel_size=3, stride=1, pad=1, dilation=1, group=4, offset_scale=1.0):
        super().__init__()
        self.channels, self.kernel_size, self.stride, self.pad, self.dilation = channels, kernel_size, stride, pad, dilation
        self.group, self.group_channels, self.offset_scale = group, channels // group, offset_scale
        P = kernel_size * kernel_size
        self.dw_conv = nn.Sequential(nn.Conv2d(channels, channels, kernel_size, 1, (kernel_size - 1) // 2, groups=channels),
                                     _LNChannelsLast(channels), nn.GELU())
        self.offset = nn.Linear(channels, group * P * 2)
        self.mask = nn.Linear(channels, group * P)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)

    def forward(self, x):   # (N, H, W, C) -> (N, Ho, Wo, C), modules/dcnv3.py:318-356
        N, H, W, _ = x.shape
        xp = self.input_proj(x)
        x1 = self.dw_conv(x.permute(0, 3, 1, 2))
        offset = self.offset(x1)
        mask = F.softmax(self.mask(x1).reshape(N, H, W, self.group, -1), -1).reshape(N, H, W, -1).type(xp.dtype)
        k, s, p, d = self.kernel_size, self.stride, self.pad, self.dilation
        if self.differentiable:
            y = core.CoreFunction.apply(xp.contiguous(), offset.contiguous(), mask.contiguous(),
                                        (k, k, s, s, p, p, d, d, self.group, self.group_channels, self.offset_scale), 0)
        else:
            y = core.forward(xp.contiguous(), offset.contiguous(), mask.contiguous(), k, k, s, s, p, p, d, d, self.group,
                             self.group_channels, self.offset_scale, 0)
        return self.output_proj(y)


class DCNv3_C(nn.Module):
    def __init__(self, cin, cout, stride=2, groups=4):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)
        self.dcnv3 = DCNv3(cout, kernel_size=3, stride=stride, group=groups)
        self.bn = nn.BatchNorm2d(cout)   # built, never used (network/dcnv3.py:29,37)

    def forward(self, x):
        return self.dcnv3(self.conv(x).permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class MAPEncoder(nn.Module):   # conv_pnp_net.py:203-332: 3 x [DCNv3_C s2 -> GN32 -> ReLU]
    def __init__(self, cin=3, featdim=256):
        super().__init__()
        self.features = nn.ModuleList()
        for i in range(3):
            self.features += [DCNv3_C(cin if i == 0 else featdim, featdim), nn.GroupNorm(32, featdim), nn.ReLU()]

    def forward(self, x):
        for layer in self.features:
            x = layer(x)
        return x


# ------------------------------------------------------------------------------------------------------
# coordinate-map decoder (xyz_head.py:195-366)
# ------------------------------------------------------------------------------------------------------
class ConvModule(nn.Module):   # conv3x3 (no bias) -> GN32 -> GELU; the GroupNorm is registered twice (conv_module.py:181-183)
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.norm = nn.GroupNorm(32, cout)
        self.gn = self.norm

    def forward(self, x):
        return F.gelu(self.norm(self.conv(x)))


class TopDownXyzHead(nn.Module):
    def __init__(self, in_dim, feat_dim=256):
        super().__init__()
        f = [nn.ConvTranspose2d(in_dim, feat_dim, 3, stride=2, padding=1, output_padding=1, bias=False),
             nn.GroupNorm(32, feat_dim), nn.GELU(), ConvModule(feat_dim, feat_dim), ConvModule(feat_dim, feat_dim)]
        for _ in range(2):
            f += [nn.UpsamplingBilinear2d(scale_factor=2), ConvModule(feat_dim, feat_dim), ConvModule(feat_dim, feat_dim)]
        self.features = nn.ModuleList(f)
        self.out_layer = nn.Conv2d(feat_dim, 3, 1)

    def forward(self, x):
        for layer in self.features:
            x = layer(x)
        out = self.out_layer(x)   # (B, 3, 64, 64): x, y, z maps (xyz_head.py:352-360)
        return out[:, 0:1], out[:, 1:2], out[:, 2:3]


# ------------------------------------------------------------------------------------------------------
# PnP regression head (conv_pnp_net.py:18-201), size head (pose_head.py:17-51)
# ------------------------------------------------------------------------------------------------------
class ConvPnPNet(nn.Module):
    def __init__(self, cin=5, featdim=128):
        super().__init__()
        self.features = nn.ModuleList()
        for i in range(3):
            self.features += [nn.Conv2d(cin if i == 0 else featdim, featdim, 3, 2, 1, bias=False),
                              nn.GroupNorm(32, featdim), nn.ReLU()]
        self.fc1, self.fc2 = nn.Linear(featdim * 64, 1024), nn.Linear(1024, 256)
        self.fc1_z, self.fc2_z = nn.Linear(featdim * 64, 1024), nn.Linear(1024, 256)
        self.fc_z, self.fc_r, self.fc_t = nn.Linear(256, 1), nn.Linear(256, 6), nn.Linear(256, 2)

    def forward(self, x):
        for layer in self.features:
            x = layer(x)
        flat = x.flatten(1)   # NCHW order (conv_pnp_net.py:168-170)
        h = F.leaky_relu(self.fc2(F.leaky_relu(self.fc1(flat), 0.1)), 0.1)
        hz = F.leaky_relu(self.fc2_z(F.leaky_relu(self.fc1_z(flat), 0.1)), 0.1)
        return self.fc_r(h), torch.cat([self.fc_t(h), self.fc_z(hz)], 1)


class SizeHead(nn.Module):
    def __init__(self, in_dim=1024, feat=128):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv1d(in_dim, feat, 1), nn.Conv1d(feat, 3, 1)
        self.bn1 = nn.BatchNorm1d(feat)

    def forward(self, x):   # eval mode: Dropout(0.2) is the identity
        x = x.flatten(2).max(-1, keepdim=True).values
        return self.conv2(F.relu(self.bn1(self.conv1(x)))).squeeze(2)


# ------------------------------------------------------------------------------------------------------
# backbone of the synthetic runs: ResNet-34 trunk (network/resnet.py:18-43, :96-148) + 1x1 neck to 1024
# ------------------------------------------------------------------------------------------------------
class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout)
        self.conv2, self.bn2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        r = x if self.downsample is None else self.downsample(x)
        return F.relu(self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x))))) + r)


class _Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64)
        cin = 64
        for li, (c, n, s) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), 1):
            setattr(self, f"layer{li}", nn.Sequential(*[_BasicBlock(cin if b == 0 else c, c, s if b == 0 else 1) for b in range(n)]))
            cin = c

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class Backbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk, self.neck = _Trunk(), nn.Conv2d(512, 1024, 1)

    def forward(self, x):
        return [self.neck(self.trunk(x))]   # features_only list, like timm's convnext (backbone.py:36-46)


# ------------------------------------------------------------------------------------------------------
# pose decode
# ------------------------------------------------------------------------------------------------------
def rot6d_to_mat(d6):   # rot_reps.py:34-55
    x = F.normalize(d6[..., 0:3], p=2, dim=-1)
    z = F.normalize(torch.cross(x, d6[..., 3:6], dim=-1), p=2, dim=-1)
    return torch.stack((x, torch.cross(z, x, dim=-1), z), dim=-1)


def axangle2mat(axis, angle):   # transforms3d.axangles.axangle2mat (Rodrigues), float64
    x, y, z = axis
    n = math.sqrt(x * x + y * y + z * z)
    x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return np.array([[x * x * C + c, x * y * C - z * s, z * x * C + y * s],
                     [x * y * C + z * s, y * y * C + c, y * z * C - x * s],
                     [z * x * C - y * s, y * z * C + x * s, z * z * C + c]])


def allo_to_ego_mat(rot, trans):   # utils.py:29-60, src/dst "mat", cam_ray (0,0,1); numpy like the reference
    cam_ray = np.asarray((0, 0, 1.0))
    obj_ray = trans.copy() / np.linalg.norm(trans)
    angle = math.acos(cam_ray.dot(obj_ray))
    if angle > 0:
        return np.dot(axangle2mat(np.cross(cam_ray, obj_ray), angle), rot)
    return rot.copy()


def pose_from_predictions_test(rots, centroids, z_vals, cams, centers, resize_ratios, whs):
    """pose_from_pred_centroid_z.py:59-157 with z_type 'REL', is_allo, dataset 'Real'."""
    if cams.dim() == 2:
        cams = cams.unsqueeze(0)
    cx = centroids[:, 0:1] * whs[:, 0:1] + centers[:, 0:1]
    cy = centroids[:, 1:2] * whs[:, 1:2] + centers[:, 1:2]
    z = z_vals * resize_ratios.view(-1, 1)
    trans = torch.cat([z * (cx - cams[:, 0:1, 2]) / cams[:, 0:1, 0], z * (cy - cams[:, 1:2, 2]) / cams[:, 1:2, 1], z], 1)
    r = rots.detach().cpu().numpy()
    t = trans.detach().cpu().numpy()
    ego = np.zeros_like(r)
    for i in range(r.shape[0]):   # host loop of the reference (:142-156)
        ego[i] = allo_to_ego_mat(r[i], t[i])
    return torch.from_numpy(ego), trans


def quat2mat(q):   # pose_utils/pose_utils.py:348-412 (eps = 0)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    X, Y, Z = 2 * x, 2 * y, 2 * z
    return torch.stack([1 - (y * Y + z * Z), x * Y - w * Z, x * Z + w * Y, x * Y + w * Z, 1 - (x * X + z * Z), y * Z - w * X,
                        x * Z - w * Y, y * Z + w * X, 1 - (x * X + y * Y)], dim=1).reshape(-1, 3, 3)


def pose_from_predictions_train(rots, centroids, z_vals, cams, centers, resize_ratios, whs, eps=1e-4):
    """pose_from_pred_centroid_z.py:160-249 (z_type 'REL', is_allo) + allo_to_ego_mat_torch (pose_utils/utils.py:198-229)."""
    if cams.dim() == 2:
        cams = cams.unsqueeze(0)
    cx = centroids[:, 0:1] * whs[:, 0:1] + centers[:, 0:1]
    cy = centroids[:, 1:2] * whs[:, 1:2] + centers[:, 1:2]
    z = z_vals * resize_ratios.view(-1, 1)
    trans = torch.cat([z * (cx - cams[:, 0:1, 2]) / cams[:, 0:1, 0], z * (cy - cams[:, 1:2, 2]) / cams[:, 1:2, 1], z], 1)
    cam_ray = torch.tensor([0, 0, 1.0], dtype=trans.dtype, device=trans.device)
    obj_ray = trans / (torch.norm(trans, dim=1, keepdim=True) + eps)
    angle = obj_ray[:, 2:3].acos()
    axis = torch.cross(cam_ray.expand_as(obj_ray), obj_ray, dim=-1)
    axis = axis / (torch.norm(axis, dim=1, keepdim=True) + eps)
    q = torch.cat([torch.cos(angle / 2.0), axis * torch.sin(angle / 2.0)], dim=1)
    return torch.matmul(quat2mat(q), rots), trans


# ------------------------------------------------------------------------------------------------------
# MAPTransformerEncoer (attention_pnp_net.py:126-157, `--nocsmap_encoder=att`).  PARITY UNPINNED: its blocks are
# `timm.models.vision_transformer.Block` (timm==0.9.6, GIVEPose_env.yml), a third-party dependency that is neither vendored
# under /root/reference nor installed here, so the reference cannot be run for this branch.  Restated from timm 0.9.6's
# published source: Block(dim, num_heads) = x + proj(softmax(q k^T hd^-0.5) v) on LayerNorm(x) [qkv_bias=False, eps 1e-5],
# then x + fc2(GELU(fc1(LayerNorm(x)))) with hidden = 4 dim; LayerScale / DropPath are identities at the defaults.
# ------------------------------------------------------------------------------------------------------
class _OAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.scale = num_heads, (dim // num_heads) ** -0.5
        self.qkv, self.proj = nn.Linear(dim, dim * 3, bias=False), nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class _OMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _OBlock(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm1, self.attn, self.norm2, self.mlp = nn.LayerNorm(dim), _OAttention(dim, num_heads), nn.LayerNorm(dim), _OMlp(dim, 4 * dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _OPatchEmbed(nn.Module):
    def __init__(self, patch, cin, dim):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, patch, patch)

    def forward(self, x):   # attention_pnp_net.py:296-303
        return self.proj(x).flatten(2).transpose(1, 2)


class MAPTransformerEncoer(nn.Module):
    def __init__(self, img_size=64, patch_size=8, in_chans=3, embed_dim=256, depth=3, num_heads=8):
        super().__init__()
        self.embed_dim = embed_dim
        self.norm = nn.LayerNorm(embed_dim)
        self.patch_embed = _OPatchEmbed(patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, (img_size // patch_size) ** 2, embed_dim))
        self.block = nn.ModuleList([_OBlock(embed_dim, num_heads) for _ in range(depth)])

    def forward(self, x):   # :143-157
        x = self.patch_embed(x) + self.pos_embed
        for blk in self.block:
            x = blk(x)
        x = self.norm(x).permute(0, 2, 1)
        return x.reshape(x.shape[0], self.embed_dim, 8, 8)


# ------------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------------
class PoseNet(nn.Module):
    def __init__(self, nocsmap_encoder="conv"):
        super().__init__()
        self.backbone = Backbone()
        self.xyz_nocs_head = TopDownXyzHead(1024)
        self.size_head = SizeHead(1024)
        self.nocs_encoder = MAPEncoder(3, 256) if nocsmap_encoder == "conv" else MAPTransformerEncoer()   # PoseNet.py:152-157
        self.feat_reducer = nn.Conv2d(1024, 256, 1)
        self.xyz_deform_head = TopDownXyzHead(512)
        self.pnp_net = ConvPnPNet(5, 128)

    def forward(self, data, device="cpu", do_loss=False, pred_scale=None):   # PoseNet.py:173-231
        img = data["roi_img"].to(device)
        mask_out = data["roi_mask_deform" if do_loss else "roi_mask"].to(device)[:, :, ::4, ::4]   # Resize(64, NEAREST) of a 256x256 map: src = floor(4*i) (PoseNet.py:170,180)
        feat = self.backbone(img)
        pred_size = self.size_head(feat[0])
        nocs = torch.cat(self.xyz_nocs_head(feat[0]), 1)
        nocs_feat = self.nocs_encoder(nocs)
        feat_cat = torch.cat([self.feat_reducer(feat[0]), nocs_feat], 1)
        ivfc = torch.cat(self.xyz_deform_head(feat_cat), 1)
        rot6, t = self.pnp_net(torch.cat([ivfc, data["roi_coord_2d"].to(device)], 1))
        mean_size = data["mean_size"].to(device)
        pred_size = pred_size + mean_size / mean_size.norm(dim=1).unsqueeze(-1)
        decode = pose_from_predictions_train if do_loss else pose_from_predictions_test
        rot, trans = decode(rot6d_to_mat(rot6), t[:, :2], t[:, 2:3], data["cam_K"].to(device), data["bbox_center"].to(device),
                            data["resize_ratio"].to(device), data["roi_wh"].to(device))
        return {"rot": rot, "trans": trans, "size": pred_size, "mask": mask_out, "nocs_coor": nocs, "ivfc_coor": ivfc}


# ------------------------------------------------------------------------------------------------------
# weights and synthetic inputs -- shared by the golden generator, tests and bench
# ------------------------------------------------------------------------------------------------------
def init_weights(net: nn.Module, mode: str, seed: int = 0) -> None:
    """``reference``: the reference's initialisation (std 1e-3 everywhere in the heads) -- nearly vacuous
    for the DCNv3 branch.  ``o1``: O(1) activations, offsets ~ N(0,1) px and non-uniform masks, so that the sampler
    leaves the integer grid.  Deterministic given ``seed`` (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in net.named_modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d, nn.ConvTranspose2d, nn.Linear)):
                fan_in = m.weight[0].numel() if not isinstance(m, nn.ConvTranspose2d) else m.weight.shape[0] * 9 // 4
                if name.startswith("backbone"):
                    std = math.sqrt(1.0 / fan_in)
                elif mode == "reference":
                    std = 0.01 if name.endswith(("out_layer", "fc_r", "fc_t")) else 0.001
                else:
                    std = math.sqrt(1.0 / fan_in)
                    if name.endswith(("dcnv3.offset", "dcnv3.mask")):
                        std = 1.0 / math.sqrt(fan_in)   # inputs are LN+GELU outputs (O(1)) -> offsets / logits ~ N(0, ~0.4..1)
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * (0.0 if mode == "reference" else 0.1))
            elif isinstance(m, (nn.GroupNorm, nn.LayerNorm, nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.fill_(1.0)
                m.bias.zero_()
        if mode != "reference":   # make the offsets a few pixels wide
            for m in net.modules():
                if isinstance(m, DCNv3) or type(m).__name__ == "DCNv3":
                    m.offset.weight.mul_(2.0)
        for name, p in net.named_parameters():   # MAPTransformerEncoer: trunc_normal_(pos_embed, std=.02) (attention_pnp_net.py:139)
            if name.endswith("pos_embed"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).clamp_(-2.0, 2.0))


def make_inputs(B: int, seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(1000 + seed)
    r = lambda *s: torch.rand(*s, generator=g)
    return {
        "roi_img": torch.randn(B, 3, 256, 256, generator=g),
        "roi_mask": (r(B, 1, 256, 256) > 0.5).float(),
        "roi_coord_2d": r(B, 2, 64, 64) * 2 - 1,
        "cam_K": torch.tensor(CAM_K).repeat(B, 1, 1),
        "mean_size": r(B, 3) + 0.1,
        "roi_wh": r(B, 2) * 100 + 50,
        "bbox_center": r(B, 2) * 300 + 100,
        "resize_ratio": r(B) * 0.3 + 0.2,
    }
