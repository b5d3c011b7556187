"""Per-RoI ``PoseNet`` inference forward on B200: host-side mirror of the reference model surface.

Mirrors (reference file:line, relative to /root/reference):

* ``PoseNet.__init__/forward(data, device, do_loss=False, pred_scale=None)``   network/PoseNet.py:134-231
* ``DCNv3`` / ``DCNv3_C``                     network/ops_dcnv3/modules/dcnv3.py:221-356, network/dcnv3.py:23-38
* ``MAPEncoder`` / ``ConvPnPNet``             network/conv_pnp_net.py:203-332 / :18-201
* ``TopDownXyzHead`` / ``ConvModule``         network/xyz_head.py:195-366, torch_utils/layers/conv_module.py:57-234
* ``SizeHead``                                network/pose_head.py:17-51
* pose decode                                 network/pose_utils/{rot_reps.py:34-55, pose_from_pred_centroid_z.py:59-157,
                                              utils.py:29-84}

Same module tree and state-dict keys (a reference checkpoint of the heads loads with ``strict=True``), same input dict,
same output dict -- ``rot`` comes back on the CPU like the reference's test path (``pose_from_pred_centroid_z.py:157``)
unless ``cfg.rot_on_cpu`` is cleared.

What runs where (inference, ``torch.no_grad``):

* DCNv3 core, its ``dw_conv -> LayerNorm -> GELU`` prologue, every ``GroupNorm -> ReLU/GELU`` (+ the bilinear x2
  upsampling that follows it in the decoder) and the whole pose decode are hand-written sm_100a kernels behind the C ABI
  (``givepose_b200/ops.py``, ``functions.py``); activations stay channel-last end to end, so the reference's
  NCHW<->NHWC permutes (``network/dcnv3.py:33-36``) disappear.
* The mask softmax is fused into the sampler (``mask_is_logits``), and ``dw_conv/LN/GELU/offset/mask`` are evaluated only
  for the ``N*Ho*Wo`` pixel rows the sampler actually reads: 4x less work at stride 2, same result.
* Dense projections / convolutions (1x1, Linear, 3x3, deconv, backbone) are library GEMMs/convs (cuBLAS / cuDNN through
  torch) -- fp32 with TF32 disabled in ``precision='fp32'`` (the 1e-4 parity mode), bf16 tensor cores in ``'bf16'``.
* With autograd enabled (training step) the glue falls back to differentiable torch CUDA ops around ``DCNv3Function``
  (custom backward kernel); there is no CPU path in either mode.
"""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .functions import DCNv3Function, dcnv3_forward, dcnv3_forward_packed

# decoder 3x3 convolutions at bf16 inference: 'tc' = the hand-written tcgen05 implicit GEMM (conv3x3_tc.cu), 'cudnn' = library
DECODER_CONV = os.environ.get("GP_DECODER_CONV", "tc")
# ConvModule -> ConvModule at one resolution: apply the first module's GroupNorm + GELU inside the second module's convolution
# (conv3x3_tc.cu XFORM; bit-identical, tests/test_conv3x3_gpu.py).  OFF by default: measured 33 % slower per fused convolution
# (profiles/r02_it12_fused_input_norm.json) -- every operand element passes through three horizontal-tap slabs with a 2x row halo,
# so the GELU is evaluated six times per element and the transform needs all of the SM's issue slots.
DECODER_FUSE_NORM = os.environ.get("GP_DECODER_FUSE_NORM", "0") != "0"


@dataclass
class PoseNetConfig:
    """The absl FLAGS the reference reads at construction / call time (config/config.py defaults)."""
    img_size: int = 256
    out_res: int = 64
    feat_ts: int = 128            # SizeHead hidden width (FLAGS.feat_ts)
    size_head_out_dim: int = 3
    r_type: str = "allo_rot6d"
    t_type: str = "site"
    dataset: str = "Real"         # 'wild6d' rescales z by fx/590 (pose_from_pred_centroid_z.py:110-111)
    precision: str = "fp32"       # 'fp32' (TF32 off, parity mode) | 'bf16'
    rot_on_cpu: bool = True       # reference behaviour at test time
    nocsmap_encoder: str = "conv" # 'conv' (MAPEncoder, DCNv3) | 'att' (MAPTransformerEncoer) -- FLAGS.nocsmap_encoder
    tc_linear: bool = True        # bf16 inference: PnP regression trunk (fc1||fc1_z, fc2, fc2_z + LeakyReLU) on the hand-written
                                  # tcgen05 dense layer (ops.linear_bf16) instead of cuBLAS + separate activation kernels
    h2d_chunk_rois: int = 256     # host inputs: RoI crops are uploaded in chunks of this many RoIs on a copy stream while the
                                  # backbone runs on the previous chunk (0 = one blocking upload like the reference)


def _fused(x: torch.Tensor) -> bool:
    return not torch.is_grad_enabled()


def _cached(p: torch.Tensor, dtype, fn=None, tag=""):
    """Derived inference copy of a parameter (dtype cast, channels_last layout, reshapes ...), rebuilt when the parameter
    changes (in-place updates bump ``_version``).  Keeps the fp32 master weights / state-dict untouched in bf16 mode and
    avoids re-casting every weight on every forward (what autocast would do)."""
    cache = getattr(p, "_gp_cache", None)
    if cache is None:
        cache = {}
        p._gp_cache = cache
    key = (dtype, tag)
    hit = cache.get(key)
    if hit is None or hit[0] != p._version or hit[1].device != p.device:
        with torch.no_grad():
            t = p.detach()
            t = (fn(t) if fn is not None else t).to(dtype)
            t = t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.contiguous()
        cache[key] = (p._version, t)
        return t
    return hit[1]


def _lin(x, lin: nn.Linear):
    """Linear on channel-last rows; inference uses the cached weight copy in the activation dtype."""
    if not _fused(x):
        return lin(x)
    return F.linear(x, _cached(lin.weight, x.dtype), None if lin.bias is None else _cached(lin.bias, x.dtype))


def _ln(x, ln: nn.LayerNorm):
    """LayerNorm over the last dimension; inference uses cached parameter copies in the activation dtype."""
    if not _fused(x):
        return ln(x)
    return F.layer_norm(x, ln.normalized_shape, _cached(ln.weight, x.dtype), _cached(ln.bias, x.dtype), ln.eps)


def _conv1x1_rows(x, conv: nn.Conv2d):
    """A 1x1 convolution is a Linear over the channel-last rows."""
    if not _fused(x):
        return F.linear(x, conv.weight.flatten(1), conv.bias)
    return F.linear(x, _cached(conv.weight, x.dtype, lambda w: w.flatten(1), "rows"),
                    None if conv.bias is None else _cached(conv.bias, x.dtype))


# ----------------------------------------------------------------------------------------------------------
# DCNv3 module + NCHW wrapper
# ----------------------------------------------------------------------------------------------------------
class _ToChannelsLast(nn.Module):
    def forward(self, x):
        return x.permute(0, 2, 3, 1)


class DCNv3(nn.Module):
    """Drop-in for ``ops_dcnv3.modules.dcnv3.DCNv3``: channel-last ``(N,H,W,C) -> (N,Ho,Wo,C)``."""

    def __init__(self, channels=64, kernel_size=3, dw_kernel_size=None, stride=1, pad=1, dilation=1, group=4,
                 offset_scale=1.0, act_layer="GELU", norm_layer="LN", center_feature_scale=False, remove_center=False):
        super().__init__()
        if channels % group != 0:
            raise ValueError(f"channels must be divisible by group, but got {channels} and {group}")
        if act_layer != "GELU" or norm_layer != "LN" or center_feature_scale:
            raise NotImplementedError("GIVEPose builds DCNv3 with GELU / LN / center_feature_scale=False (network/dcnv3.py:26-28)")
        dw_kernel_size = kernel_size if dw_kernel_size is None else dw_kernel_size
        self.channels, self.kernel_size, self.dw_kernel_size = channels, kernel_size, dw_kernel_size
        self.stride, self.pad, self.dilation, self.group = stride, pad, dilation, group
        self.group_channels, self.offset_scale = channels // group, offset_scale
        self.center_feature_scale, self.remove_center = False, int(remove_center)
        if self.remove_center and kernel_size % 2 == 0:
            raise ValueError("remove_center is only compatible with odd kernel size.")
        P = kernel_size * kernel_size - self.remove_center
        self.dw_conv = nn.Sequential(
            nn.Conv2d(channels, channels, dw_kernel_size, 1, (dw_kernel_size - 1) // 2, groups=channels),
            nn.Sequential(_ToChannelsLast(), nn.LayerNorm(channels, eps=1e-6)), nn.GELU())
        self.offset = nn.Linear(channels, group * P * 2)
        self.mask = nn.Linear(channels, group * P)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        self._reset_parameters()
        self._cache = {}
        self.fuse_whole_module = True   # first-layer (K = 3) inference path: ops.dcnv3_smallk_fused
        self.fuse_offset_mask = True    # bf16 inference: offset || mask as one tcgen05 GEMM, rows consumed packed by the sampler

    def _reset_parameters(self):   # modules/dcnv3.py:308-316
        for m in (self.offset, self.mask):
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)
        for m in (self.input_proj, self.output_proj):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)

    def _out_hw(self, H, W):
        f = lambda n: (n + 2 * self.pad - (self.dilation * (self.kernel_size - 1) + 1)) // self.stride + 1
        return f(H), f(W)

    def _dw_params(self, device):
        key = (device, self.dw_conv[0].weight._version, self.dw_conv[1][1].weight._version)
        if self._cache.get("key") != key:
            conv, ln = self.dw_conv[0], self.dw_conv[1][1]
            self._cache = {"key": key,
                           "w_t": conv.weight.detach().float().reshape(self.channels, -1).t().contiguous(),
                           "b": conv.bias.detach().float().contiguous(),
                           "ln_w": ln.weight.detach().float().contiguous(), "ln_b": ln.bias.detach().float().contiguous()}
        return self._cache

    def _fusable(self, x, C):
        return (_fused(x) and self.dw_kernel_size == 3 and C in (128, 256, 512) and x.is_cuda
                and x.dtype in (torch.float32, torch.bfloat16, torch.float16))

    def _geom(self):
        k, s, p, d = self.kernel_size, self.stride, self.pad, self.dilation
        return (k, k, s, s, p, p, d, d, self.group, self.group_channels, self.offset_scale)

    def _offmask_params(self, device):
        """``offset`` and ``mask`` read the same activation (modules/dcnv3.py:330-334): one [G*P*3 (+ pad to a multiple of 8), C]
        bf16 weight and one fp32 bias, so both run as ONE tcgen05 GEMM whose rows the sampler consumes in place."""
        ps = (self.offset.weight, self.offset.bias, self.mask.weight, self.mask.bias)
        key = (device,) + tuple(p._version for p in ps)
        if self._cache.get("omkey") != key:
            with torch.no_grad():
                n = self.offset.out_features + self.mask.out_features
                pad = (-n) % 8
                w = torch.cat([self.offset.weight, self.mask.weight, self.offset.weight.new_zeros(pad, self.channels)])
                b = torch.cat([self.offset.bias, self.mask.bias, self.offset.bias.new_zeros(pad)])
                self._cache.update({"omkey": key, "om_w": w.detach().to(torch.bfloat16).contiguous(), "om_b": b.detach().float().contiguous()})
        return self._cache["om_w"], self._cache["om_b"]

    def _sample(self, x, x1):
        """offset / mask-logit Linears on the computed rows, fused-softmax sampler, output projection."""
        if (self.fuse_offset_mask and _fused(x1) and x1.dtype == torch.bfloat16 and x1.is_cuda and self.channels % 8 == 0
                and self.group_channels % 4 == 0):
            w, b = self._offmask_params(x1.device)
            om = ops.linear_bf16(x1.reshape(-1, self.channels), w, b)        # one launch instead of two cuBLAS GEMMs
            x = dcnv3_forward_packed(x.contiguous(), om, *self._geom(), 256, self.remove_center)
            return _lin(x, self.output_proj)
        offset = _lin(x1, self.offset)
        logits = _lin(x1, self.mask)
        x = dcnv3_forward(x.contiguous(), offset.contiguous(), logits.contiguous(), *self._geom(), 256, self.remove_center,
                          mask_is_logits=True)
        return _lin(x, self.output_proj)

    def _composed_params(self, conv):
        """``conv`` (1x1, K -> C) feeds this module and both of its consumers are linear in its output: compose
        ``input_proj o conv`` and ``dw_conv o conv`` (fp64 products, stored fp32); see ``include/givepose_b200.h``."""
        dw, ln = self.dw_conv[0], self.dw_conv[1][1]
        ps = (conv.weight, conv.bias, self.input_proj.weight, self.input_proj.bias, dw.weight, dw.bias, ln.weight, ln.bias,
              self.output_proj.weight, self.output_proj.bias)
        key = (conv.weight.device,) + tuple(p._version for p in ps)
        if self._cache.get("ckey") != key:
            with torch.no_grad():
                Wc, bc = conv.weight.detach().double().flatten(1), conv.bias.detach().double()            # (C,K), (C,)
                Wip, bip = self.input_proj.weight.detach().double(), self.input_proj.bias.detach().double()
                wdw = dw.weight.detach().double().reshape(self.channels, 9)                              # (C, 9), tap = ky*3+kx
                w_eff = torch.cat([wdw.t()[:, None, :] * Wc.t()[None, :, :], (wdw * bc[:, None]).t()[:, None, :]], dim=1)
                # whole-module composition (ops.dcnv3_smallk_fused): out = W2^T [S_g,j ; S0_g] + b_out
                Wp, bp = Wip @ Wc, Wip @ bc + bip                                                        # (C,K), (C,)
                Wo, G, gc = self.output_proj.weight.detach().double(), self.group, self.group_channels
                WoG = Wo.reshape(Wo.shape[0], G, gc)                                                     # (O, G, gc)
                w2 = torch.cat([torch.einsum("ogc,gcj->gjo", WoG, Wp.reshape(G, gc, -1)),
                                torch.einsum("ogc,gc->go", WoG, bp.reshape(G, gc))[:, None, :]], dim=1)  # (G, K+1, O)
                self._cache.update({"w2": w2.reshape(G * (Wc.shape[1] + 1), -1).float().contiguous(),
                                    "b_out": self.output_proj.bias.detach().float().contiguous()})
                self._cache.update({"ckey": key, "wp_t": (Wip @ Wc).t().float().contiguous(), "bp": (Wip @ bc + bip).float().contiguous(),
                                    "w_eff": w_eff.float().contiguous(), "dw_b": dw.bias.detach().float().contiguous(),
                                    "cln_w": ln.weight.detach().float().contiguous(), "cln_b": ln.bias.detach().float().contiguous()})
        return self._cache

    def forward_after_conv1x1(self, x_small, conv):
        """``self(conv(x_small))`` for a 1x1 ``conv`` with K = 3 input channels (first MAPEncoder layer) without writing the
        C-channel convolution output."""
        N, H, W, K = x_small.shape
        if not (K == 3 and conv.bias is not None and self._fusable(x_small, self.channels)):
            return self(_conv1x1_rows(x_small, conv))
        Ho, Wo = self._out_hw(H, W)
        rows = min(N * Ho * Wo, N * H * W)
        c = self._composed_params(conv)
        x_small = x_small.contiguous()
        x1 = ops.smallk_dwconv3x3_ln_gelu(x_small, c["w_eff"], c["dw_b"], c["cln_w"], c["cln_b"], rows, eps=1e-6)
        if self.fuse_whole_module and self.group == 4 and self.kernel_size == 3 and not self.remove_center and self.channels == 256:
            # input_proj, the core and output_proj collapse into one sampling kernel over the 3-channel map + a 16 -> 256 map:
            # the 256-channel input_proj tensor is never written
            return ops.dcnv3_smallk_fused(x_small, _lin(x1, self.offset).contiguous(), _lin(x1, self.mask).contiguous(), c["w2"], c["b_out"],
                                          self._geom(), self.remove_center)
        x = ops.small_k_linear(x_small, c["wp_t"], c["bp"])
        return self._sample(x, x1)

    def forward(self, input):
        N, H, W, C = input.shape
        geom = self._geom()
        if self._fusable(input, C):
            # the sampler reads offset / mask through their flat [N*Ho*Wo]-row prefix (cuh:229,243-244): compute only those rows
            Ho, Wo = self._out_hw(H, W)
            rows = min(N * Ho * Wo, N * H * W)
            c = self._dw_params(input.device)
            x1 = ops.dwconv3x3_ln_gelu(input.contiguous(), c["w_t"], c["b"], c["ln_w"], c["ln_b"], rows, eps=1e-6)
            return self._sample(_lin(input, self.input_proj), x1)
        x = _lin(input, self.input_proj)
        x1 = self.dw_conv(input.permute(0, 3, 1, 2))
        offset = self.offset(x1)
        mask = F.softmax(self.mask(x1).reshape(N, H, W, self.group, -1), -1).reshape(N, H, W, -1).type(x.dtype)
        x = DCNv3Function.apply(x.contiguous(), offset.contiguous(), mask.contiguous(), *geom, 256, self.remove_center)
        return _lin(x, self.output_proj)


class DCNv3_C(nn.Module):
    """``network/dcnv3.py:23-38``: 1x1 conv -> DCNv3 (channel-last inside).  NCHW in / NCHW out like the reference;
    ``forward_nhwc`` is the permute-free entry the fused encoder uses."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, groups=4, dilation=1, padding=1, bias=False):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.dcnv3 = DCNv3(out_channels, kernel_size=kernel_size, stride=stride, group=groups, dilation=dilation)
        self.bn = nn.BatchNorm2d(out_channels)   # built but unused in the reference (:29,37); kept for checkpoint keys
        self.gelu = nn.GELU()

    def forward_nhwc(self, x):
        if x.shape[-1] == 3 and _fused(x):
            return self.dcnv3.forward_after_conv1x1(x, self.conv)
        return self.dcnv3(_conv1x1_rows(x, self.conv))

    def forward(self, x):
        return self.forward_nhwc(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def _gn_act_nhwc(x, gn: nn.GroupNorm, act: str, upsample2x=False):
    """GroupNorm -> act (-> bilinear x2) on channel-last activations: fused kernel at inference, torch ops under autograd."""
    if _fused(x):
        return ops.groupnorm_act(x.contiguous(), _cached(gn.weight, torch.float32), _cached(gn.bias, torch.float32),
                                 gn.num_groups, gn.eps, act, upsample2x)
    # training step: one autograd node with our forward + backward kernels (storage dtype in and out: no fp32 round trip)
    y = ops.GroupNormAct.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, act)
    if upsample2x:
        y = ops.UpsampleBilinear2x.apply(y)
    return y


def _conv_nhwc(x, conv: nn.Module):
    """cuDNN convolution on a channel-last activation without layout copies: (N,H,W,C) viewed as channels_last NCHW."""
    xn = x.permute(0, 3, 1, 2)
    if not _fused(x):
        return conv(xn).permute(0, 2, 3, 1)
    w = _cached(conv.weight, x.dtype)
    b = None if conv.bias is None else _cached(conv.bias, x.dtype)
    if isinstance(conv, nn.ConvTranspose2d):
        y = F.conv_transpose2d(xn, w, b, conv.stride, conv.padding, conv.output_padding, conv.groups, conv.dilation)
    else:
        y = F.conv2d(xn, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    return y.permute(0, 2, 3, 1)


class MAPEncoder(nn.Module):
    """``conv_pnp_net.py:203-332`` with ``FLAGS.use_dcn == 'dcnv3'``: 3 x [DCNv3_C(stride 2) -> GN(32) -> ReLU]."""

    def __init__(self, nIn, featdim=128, outdim=256, num_stride2_layers=3, num_gn_groups=32):
        super().__init__()
        self.features = nn.ModuleList()
        for i in range(num_stride2_layers):
            cin = nIn if i == 0 else featdim
            featdim = outdim if i == num_stride2_layers - 1 else featdim
            self.features += [DCNv3_C(cin, featdim, kernel_size=3, stride=2, padding=1, bias=False),
                              nn.GroupNorm(num_gn_groups, featdim), nn.ReLU(inplace=True)]
        for m in self.modules():   # :291-300
            if isinstance(m, (nn.Conv2d, nn.Conv1d, nn.ConvTranspose2d, nn.Linear)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.GroupNorm, nn.BatchNorm2d)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward_nhwc(self, x):
        for i in range(0, len(self.features), 3):
            x = _gn_act_nhwc(self.features[i].forward_nhwc(x), self.features[i + 1], "relu")
        return x

    def forward(self, coor_feat=None, mask_attention=None, cat_id=None, sp2d=None):
        return self.forward_nhwc(coor_feat.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class _Attention(nn.Module):
    """timm 0.9.6 ``vision_transformer.Attention`` at ``Block``'s defaults (``qkv_bias=False``, no q/k norm, no dropout)."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = _lin(x, self.qkv)
        if _fused(x) and N == 64 and self.head_dim == 32 and x.is_cuda:
            o = ops.mhsa_tokens(qkv.contiguous(), self.num_heads)
        else:
            q, k, v = qkv.reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4).unbind(0)
            a = ((q * self.head_dim ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
            o = (a @ v).transpose(1, 2).reshape(B, N, C)
        return _lin(o, self.proj)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return _lin(F.gelu(_lin(x, self.fc1)), self.fc2)


class ViTBlock(nn.Module):
    """timm 0.9.6 ``vision_transformer.Block(dim, num_heads)`` as ``attention_pnp_net.py:141`` builds it: pre-norm attention
    and MLP (ratio 4, exact GELU) with residuals; ``nn.LayerNorm`` default eps 1e-5; LayerScale / DropPath are identities."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1, self.attn = nn.LayerNorm(dim), _Attention(dim, num_heads)
        self.norm2, self.mlp = nn.LayerNorm(dim), _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(_ln(x, self.norm1))
        return x + self.mlp(_ln(x, self.norm2))


class _PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class MAPTransformerEncoer(nn.Module):
    """``attention_pnp_net.py:126-157`` (sic), the ``--nocsmap_encoder=att`` alternative to ``MAPEncoder``: 8x8 patch embedding
    of the 64x64x3 NOCS map -> + pos_embed -> 3 ViT blocks (dim 256, 8 heads) -> LayerNorm -> (B, 256, 8, 8).
    State-dict keys follow the reference / timm (``patch_embed.proj``, ``pos_embed``, ``block.i.{norm1,attn.qkv,attn.proj,norm2,
    mlp.fc1,mlp.fc2}``, ``norm``).  Attention over the 64 tokens runs in ``ops.mhsa_tokens`` at inference."""

    def __init__(self, img_size=64, patch_size=8, in_chans=3, embed_dim=256, depth=3, num_heads=8):
        super().__init__()
        self.embed_dim, self.patch, self.grid = embed_dim, patch_size, img_size // patch_size
        self.norm = nn.LayerNorm(embed_dim)
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.grid * self.grid, embed_dim))
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.block = nn.ModuleList([ViTBlock(embed_dim, num_heads) for _ in range(depth)])

    def forward_nhwc(self, x):
        """x: (B, 64, 64, 3) channel-last -> (B, 8, 8, 256) channel-last (token order = patch raster order)."""
        B, H, W, Cin = x.shape
        p, g = self.patch, self.grid
        # a stride = kernel convolution is a Linear over the flattened patch, in the weight's (c, ky, kx) order
        patches = x.reshape(B, g, p, g, p, Cin).permute(0, 1, 3, 5, 2, 4).reshape(B, g * g, Cin * p * p)
        conv = self.patch_embed.proj
        if _fused(x):
            t = F.linear(patches, _cached(conv.weight, x.dtype, lambda w: w.flatten(1), "rows"), _cached(conv.bias, x.dtype))
            t = t + _cached(self.pos_embed, x.dtype)
        else:
            t = F.linear(patches, conv.weight.flatten(1), conv.bias) + self.pos_embed
        for blk in self.block:
            t = blk(t)
        return _ln(t, self.norm).reshape(B, g, g, self.embed_dim)

    def forward(self, x):
        return self.forward_nhwc(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------------------
# coordinate-map decoder
# ----------------------------------------------------------------------------------------------------------
class ConvModule(nn.Module):
    def __init__(self, cin, cout, kernel_size=3, padding=1, num_gn_groups=32):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, padding=padding, bias=False)
        self.norm = nn.GroupNorm(num_gn_groups, cout)
        self.gn = self.norm   # the reference registers the same GroupNorm under both names (conv_module.py:181-183)
        self.activate = nn.GELU()
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode="fan_out", nonlinearity="relu")

    def _tc(self, x):
        """bf16 inference: the hand-written tcgen05 implicit-GEMM convolution (GroupNorm statistics from its epilogue)."""
        return (DECODER_CONV == "tc" and _fused(x) and self.conv.kernel_size == (3, 3) and self.conv.stride == (1, 1)
                and self.conv.padding == (1, 1) and self.conv.dilation == (1, 1) and self.conv.groups == 1 and self.conv.bias is None
                and self.norm.num_groups == 32 and ops.conv3x3_gn_supported(x, self.conv.out_channels))

    def conv_gn_stats(self, x, in_norm=None):
        """conv -> (y, (mean, rstd) pairs of GroupNorm(32) over y); only valid when ``_tc(x)``.  ``in_norm``: ``x`` is the raw
        convolution output of the previous ConvModule, whose GroupNorm + GELU is applied inside this convolution."""
        return ops.conv3x3_gn_bf16(x.contiguous(), _cached(self.conv.weight, torch.bfloat16, ops.pack_conv3x3_weight, "k-major"),
                                   self.norm.num_groups, self.norm.eps, in_norm=in_norm)

    def norm_params(self):
        return _cached(self.norm.weight, torch.float32), _cached(self.norm.bias, torch.float32)

    def forward_nhwc(self, x, upsample2x=False):
        if self._tc(x):
            y, stats = self.conv_gn_stats(x)
            return ops.groupnorm_apply(y, stats, _cached(self.norm.weight, torch.float32), _cached(self.norm.bias, torch.float32),
                                       self.norm.num_groups, self.norm.eps, "gelu", upsample2x)
        return _gn_act_nhwc(_conv_nhwc(x, self.conv), self.norm, "gelu", upsample2x)

    def forward(self, x):
        return self.forward_nhwc(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class TopDownXyzHead(nn.Module):
    """``xyz_head.py:195-366`` at its defaults: deconv -> GN -> GELU, 2 ConvModules, [bilinear x2, 2 ConvModules] x 2,
    1x1 out layer -> (x, y, z) maps."""

    def __init__(self, in_dim, feat_dim=256, num_gn_groups=32, xyz_num_classes=1):
        super().__init__()
        f = [nn.ConvTranspose2d(in_dim, feat_dim, 3, stride=2, padding=1, output_padding=1, bias=False),
             nn.GroupNorm(num_gn_groups, feat_dim), nn.GELU(), ConvModule(feat_dim, feat_dim), ConvModule(feat_dim, feat_dim)]
        for _ in range(2):
            f += [nn.UpsamplingBilinear2d(scale_factor=2), ConvModule(feat_dim, feat_dim), ConvModule(feat_dim, feat_dim)]
        self.features = nn.ModuleList(f)
        self.out_layer = nn.Conv2d(feat_dim, 3 * xyz_num_classes, 1)
        for m in self.modules():   # :334-347
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, std=0.001)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.out_layer.weight, std=0.01)
        nn.init.zeros_(self.out_layer.bias)

    def forward_nhwc(self, x):
        """x: (B, 8, 8, in_dim) channel-last -> (B, 64, 64, 3) channel-last."""
        fs = self.features
        x = _gn_act_nhwc(_conv_nhwc(x, fs[0]), fs[1], "gelu")

        def pair_of_modules(x, a, b):
            """ConvModule a -> ConvModule b at one resolution: (raw conv output of b, its GroupNorm statistics), or None when
            the fused path does not apply.  b's convolution applies a's GroupNorm + GELU to its operand (conv3x3_tc.cu), so
            a's activation is never written."""
            if not (DECODER_FUSE_NORM and a._tc(x) and ops.conv3x3_fused_in_supported(x) and a.norm.eps == b.norm.eps):
                return None
            ya, sa = a.conv_gn_stats(x)
            if not b._tc(ya):
                return None
            return b.conv_gn_stats(ya, in_norm=(sa,) + a.norm_params())

        def stage(x, a, b, upsample2x):
            r = pair_of_modules(x, a, b)
            if r is None:
                return b.forward_nhwc(a.forward_nhwc(x), upsample2x=upsample2x)
            return ops.groupnorm_apply(r[0], r[1], *b.norm_params(), b.norm.num_groups, b.norm.eps, "gelu", upsample2x)

        x = stage(x, fs[3], fs[4], True)               # features.5: bilinear x2 after the GN/GELU apply pass
        x = stage(x, fs[6], fs[7], True)               # features.8
        last = pair_of_modules(x, fs[9], fs[10]) if (x.shape[-1] == 256 and self.out_layer.out_channels == 3) else None
        if last is not None:
            # last ConvModule: conv -> [GN -> GELU -> out_layer 1x1] in one pass; the 256-channel activation is never written
            return ops.groupnorm_apply_conv1x1(last[0], last[1], *fs[10].norm_params(),
                                               _cached(self.out_layer.weight, torch.float32, lambda w: w.flatten(1), "rows"),
                                               _cached(self.out_layer.bias, torch.float32), fs[10].norm.num_groups, fs[10].norm.eps, "gelu")
        x = fs[9].forward_nhwc(x)
        if _fused(x) and x.shape[-1] == 256 and self.out_layer.out_channels == 3:
            # last ConvModule: conv -> [GN -> GELU -> out_layer 1x1] in one pass; the 256-channel activation is never written
            last = fs[10]
            gnw, gnb = _cached(last.norm.weight, torch.float32), _cached(last.norm.bias, torch.float32)
            ow = _cached(self.out_layer.weight, torch.float32, lambda w: w.flatten(1), "rows")
            if last._tc(x):
                y, stats = last.conv_gn_stats(x)
                return ops.groupnorm_apply_conv1x1(y, stats, gnw, gnb, ow, _cached(self.out_layer.bias, torch.float32),
                                                   last.norm.num_groups, last.norm.eps, "gelu")
            return ops.groupnorm_act_conv1x1(_conv_nhwc(x, last.conv).contiguous(), _cached(last.norm.weight, torch.float32),
                                             _cached(last.norm.bias, torch.float32),
                                             _cached(self.out_layer.weight, torch.float32, lambda w: w.flatten(1), "rows"),
                                             _cached(self.out_layer.bias, torch.float32), last.norm.num_groups, last.norm.eps, "gelu")
        x = fs[10].forward_nhwc(x)
        return _conv1x1_rows(x, self.out_layer)

    def forward(self, x):
        if isinstance(x, (tuple, list)) and len(x) == 1:
            x = x[0]
        out = self.forward_nhwc(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return out[:, 0:1], out[:, 1:2], out[:, 2:3]


# ----------------------------------------------------------------------------------------------------------
# PnP regression head, size head
# ----------------------------------------------------------------------------------------------------------
class ConvPnPNet(nn.Module):
    def __init__(self, nIn, featdim=128, rot_dim=6, num_gn_groups=32, mask_attention_type="none", flat_op="flatten"):
        super().__init__()
        if mask_attention_type != "none" or flat_op != "flatten":
            raise NotImplementedError("GIVEPose runs ConvPnPNet with mask_attention_type='none', flat_op='flatten' (config.py)")
        self.features = nn.ModuleList()
        for i in range(3):
            self.features += [nn.Conv2d(nIn if i == 0 else featdim, featdim, 3, 2, 1, bias=False),
                              nn.GroupNorm(num_gn_groups, featdim), nn.ReLU(inplace=True)]
        self.fc1, self.fc2 = nn.Linear(featdim * 64, 1024), nn.Linear(1024, 256)
        self.fc1_z, self.fc2_z = nn.Linear(featdim * 64, 1024), nn.Linear(1024, 256)
        self.fc_z, self.fc_r, self.fc_t = nn.Linear(256, 1), nn.Linear(256, rot_dim), nn.Linear(256, 2)
        self.tc_linear = True
        for m in self.modules():   # :124-134
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.fc_r.weight, std=0.01)
        nn.init.normal_(self.fc_t.weight, std=0.01)

    def forward_nhwc(self, x):
        for i in range(0, 9, 3):
            conv = self.features[i]
            if i == 0 and _fused(x) and x.shape[-1] % 8:
                # 5 input channels (3 IVFC + 2 image coordinates) send cuDNN to a SIMT kernel: zero-pad activation and weight
                # to 8 channels (exact) so the stride-2 3x3 runs as a tensor-core implicit GEMM
                pad = 8 - x.shape[-1] % 8
                w = _cached(conv.weight, x.dtype, lambda t: F.pad(t, (0, 0, 0, 0, 0, pad)), "cin8")
                y = F.conv2d(F.pad(x, (0, pad)).permute(0, 3, 1, 2), w, None, conv.stride, conv.padding).permute(0, 2, 3, 1)
            else:
                y = _conv_nhwc(x, conv)
            x = _gn_act_nhwc(y, self.features[i + 1], "relu")
        pnp_feat = x.permute(0, 3, 1, 2)
        flat = pnp_feat.reshape(x.shape[0], -1)   # NCHW flatten order (conv_pnp_net.py:168-170): checkpoint compatible
        # fc1 || fc1_z read the same 8192-wide activation: one GEMM over the concatenated weights
        tc = _fused(x) and self.tc_linear and flat.dtype == torch.bfloat16 and flat.is_cuda
        if _fused(x):
            vers = (self.fc1.weight._version, self.fc1_z.weight._version, self.fc1.bias._version, self.fc1_z.bias._version, flat.dtype, flat.device)
            if getattr(self, "_fc1_cat", (None,))[0] != vers:
                bcat = torch.cat([self.fc1.bias, self.fc1_z.bias]).detach()
                self._fc1_cat = (vers, torch.cat([self.fc1.weight, self.fc1_z.weight]).detach().to(flat.dtype).contiguous(),
                                 bcat.to(flat.dtype).contiguous(), bcat.float().contiguous())
            if tc:   # tcgen05 GEMM, bias + LeakyReLU(0.1) in its epilogue (conv_pnp_net.py:172-199)
                h = ops.linear_bf16(flat.contiguous(), self._fc1_cat[1], self._fc1_cat[3], "lrelu", 0.1)
            else:
                h = F.leaky_relu(F.linear(flat, self._fc1_cat[1], self._fc1_cat[2]), 0.1)
        else:
            h = F.leaky_relu(F.linear(flat, torch.cat([self.fc1.weight, self.fc1_z.weight]), torch.cat([self.fc1.bias, self.fc1_z.bias])), 0.1)
        if tc:
            hr = ops.linear_bf16(h[:, :1024].contiguous(), _cached(self.fc2.weight, torch.bfloat16), _cached(self.fc2.bias, torch.float32), "lrelu", 0.1)
            hz = ops.linear_bf16(h[:, 1024:].contiguous(), _cached(self.fc2_z.weight, torch.bfloat16), _cached(self.fc2_z.bias, torch.float32), "lrelu", 0.1)
        else:
            hr = F.leaky_relu(_lin(h[:, :1024], self.fc2), 0.1)
            hz = F.leaky_relu(_lin(h[:, 1024:], self.fc2_z), 0.1)
        rot = _lin(hr, self.fc_r)
        t = torch.cat([_lin(hr, self.fc_t), _lin(hz, self.fc_z)], dim=1)
        return rot, t, pnp_feat

    def forward(self, coor_feat=None, mask_attention=None, cat_id=None, sp2d=None):
        return self.forward_nhwc(coor_feat.permute(0, 2, 3, 1))


class SizeHead(nn.Module):
    def __init__(self, in_dim, out_dim, feat_dim=128):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv1d(in_dim, feat_dim, 1), nn.Conv1d(feat_dim, out_dim, 1)
        self.drop1, self.bn1 = nn.Dropout(0.2), nn.BatchNorm1d(feat_dim)
        for m in (self.conv1, self.conv2):
            nn.init.normal_(m.weight, std=0.001)
            nn.init.zeros_(m.bias)

    def forward(self, x):
        if isinstance(x, (tuple, list)) and len(x) == 1:
            x = x[0]
        x = x.flatten(2, 3).max(dim=-1, keepdim=True).values
        if _fused(x) and not self.training:   # eval: BN = affine with running stats, Dropout = identity
            bn = self.bn1
            h = F.linear(x.squeeze(2).float(), self.conv1.weight.squeeze(2), self.conv1.bias)
            h = F.relu((h - bn.running_mean) * torch.rsqrt(bn.running_var + bn.eps) * bn.weight + bn.bias)
            return F.linear(h, self.conv2.weight.squeeze(2), self.conv2.bias)[:, :3]
        x = self.conv2(self.drop1(F.relu(self.bn1(self.conv1(x)))))
        return x.squeeze(2).contiguous()[:, :3]


# ----------------------------------------------------------------------------------------------------------
# backbone of the synthetic runs: ResNet-34 trunk + 1x1 neck to the 1024 channels PoseNet hard-codes (PoseNet.py:144).
# Plain torch modules (cuDNN); the reference uses timm's pretrained ConvNeXt-B here, which needs the network.
# ----------------------------------------------------------------------------------------------------------
class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout)
        self.conv2, self.bn2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        return F.relu(self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x))))) + idt)


class ResNet34Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64)
        cin = 64
        for li, (c, n, s) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), 1):
            setattr(self, f"layer{li}", nn.Sequential(*[_Block(cin if b == 0 else c, c, s if b == 0 else 1) for b in range(n)]))
            cin = c

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


def _folded(conv: nn.Conv2d, bn: nn.BatchNorm2d, dtype):
    """eval-mode BatchNorm folded into the preceding convolution: w' = w * g/sqrt(var+eps), b' = beta - mean * g/sqrt(var+eps)."""
    vers = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, dtype,
            conv.weight.device)
    hit = getattr(conv, "_gp_fold", None)
    if hit is None or hit[0] != vers:
        with torch.no_grad():
            scale = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
            w = (conv.weight.float() * scale.view(-1, 1, 1, 1)).to(dtype).contiguous(memory_format=torch.channels_last)
            b = (bn.bias.float() - bn.running_mean.float() * scale).to(dtype).contiguous()
        hit = (vers, w, b)
        conv._gp_fold = hit
    return hit[1], hit[2]


def _conv_bn_act(x, conv, bn, relu=True, residual=None):
    """conv -> folded BN (-> + residual) (-> ReLU) with cuDNN's fused epilogue when this torch build exposes it.
    Two shapes where the library has no good sm_100 kernel take other routes (measured per 1024 RoIs, B200, bf16):
    * 1x1 stride-2 downsample convolutions (a legacy SIMT kernel: 1.22 ms for 64 -> 128 at 64x64): subsample the channel-last
      rows, then one row-GEMM (0.15 ms);
    * residual blocks with <= 128 channels (cuDNN's conv+add+ReLU runs at a third of the plain convolution's rate: 0.68 vs
      0.25 ms): plain convolution + ``ops.bias_add_relu_`` (one extra pass over the activation, still faster)."""
    w, b = _folded(conv, bn, x.dtype)
    if (conv.kernel_size == (1, 1) and conv.stride == (2, 2) and conv.padding == (0, 0) and conv.groups == 1 and residual is None
            and x.dtype in (torch.bfloat16, torch.float16)):
        rows = x.permute(0, 2, 3, 1)[:, ::2, ::2, :].contiguous()                      # (B, H/2, W/2, Cin) channel-last rows
        y = F.linear(rows, w.flatten(1), b)
        y = y.relu_() if relu else y
        return y.permute(0, 3, 1, 2)
    if (relu and residual is not None and conv.out_channels <= 128 and x.dtype in (torch.bfloat16, torch.float16)
            and residual.is_contiguous(memory_format=torch.channels_last)):
        y = F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if y.is_contiguous(memory_format=torch.channels_last):
            ops.bias_add_relu_(y.permute(0, 2, 3, 1), residual.permute(0, 2, 3, 1), _cached(b, torch.float32, tag="f32bias"))
            return y
        return y.add_(b.view(1, -1, 1, 1)).add_(residual).relu_()
    if relu and hasattr(torch, "cudnn_convolution_relu") and _conv_bn_act.fused_ok:
        try:
            if residual is None:
                return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
            return torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        except RuntimeError:
            _conv_bn_act.fused_ok = False   # fall through to conv + elementwise (still cuDNN, still on the device)
    y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    if residual is not None:
        y = y.add_(residual)
    return y.relu_() if relu else y


_conv_bn_act.fused_ok = True


def _stem_s2d(img, conv, bn, dtype):
    """7x7/2 stem over the fp32 NCHW RoI crops as a 4x4/1 convolution over the 2x2 space-to-depth image (12 -> 16 channels,
    K = 256): cuDNN's 3-channel 7x7 kernels take 10.6 ms per 1024 RoIs on B200, this form 2.6 ms.  ``ops.stem_s2d_pack`` reads
    the crops once (it replaces the dtype cast and the channels_last copy); the weight is the folded 7x7 kernel padded to
    8x8 with a zero tap in front and regrouped the same way.  Same sums as the direct convolution, different order."""
    vers = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version, dtype,
            conv.weight.device)
    hit = getattr(conv, "_gp_fold_s2d", None)
    if hit is None or hit[0] != vers:
        with torch.no_grad():
            scale = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
            w = conv.weight.float() * scale.view(-1, 1, 1, 1)                             # (O, 3, 7, 7)
            O = w.shape[0]
            w = F.pad(w, (1, 0, 1, 0)).reshape(O, 3, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4)   # o, c, ry, rx, a, b
            w = F.pad(w.reshape(O, 12, 4, 4), (0, 0, 0, 0, 0, 4)).to(dtype).contiguous(memory_format=torch.channels_last)
            b = (bn.bias.float() - bn.running_mean.float() * scale).to(dtype).contiguous()
        hit = (vers, w, b)
        conv._gp_fold_s2d = hit
    _, w, b = hit
    packed = ops.stem_s2d_pack(img, dtype)
    if dtype == torch.bfloat16 and packed.shape[2] - 3 == 128 and w.shape[0] == 64 and _stem_s2d.tc_ok:
        # hand-written tcgen05 implicit GEMM (TMA im2col through an overlapping-row tensor map): 2.8 -> 0.8 ms per 1024 RoIs
        try:
            y = ops.stem_s2d_gemm(packed, w.permute(0, 2, 3, 1).reshape(64, 256), _cached(b, torch.float32, tag="f32bias"), pool=True)
            return y.permute(0, 3, 1, 2)   # bias + ReLU + the 3x3/2 max-pool happened in the kernel's epilogue
        except RuntimeError as e:
            import warnings
            warnings.warn(f"givepose_b200: tcgen05 stem kernel unavailable ({e}); the stem runs through cuDNN from now on")
            _stem_s2d.tc_ok = False   # e.g. the driver refuses the tensor map: cuDNN path below (still on the device)
    x = packed.permute(0, 3, 1, 2)
    if hasattr(torch, "cudnn_convolution_relu") and _conv_bn_act.fused_ok:
        try:
            y = torch.cudnn_convolution_relu(x, w, b, (1, 1), (0, 0), (1, 1), 1)
        except RuntimeError:
            _conv_bn_act.fused_ok = False
            y = F.conv2d(x, w, b)
    else:
        y = F.conv2d(x, w, b)
    y = ops.maxpool3x3s2(y.permute(0, 2, 3, 1).contiguous(), relu=True)
    return y.permute(0, 3, 1, 2)


_stem_s2d.tc_ok = True


def _stem(x, conv, bn):
    """7x7/2 stem conv + folded BN, then ReLU + 3x3/2 max-pool in one channel-last kernel (ReLU commutes with max)."""
    w, b = _folded(conv, bn, x.dtype)
    y = F.conv2d(x, w, b, conv.stride, conv.padding)
    y = ops.maxpool3x3s2(y.permute(0, 2, 3, 1).contiguous(), relu=True)
    return y.permute(0, 3, 1, 2)


class ResNet34Backbone(nn.Module):
    def __init__(self, out_channels=1024):
        super().__init__()
        self.trunk, self.neck = ResNet34Trunk(), nn.Conv2d(512, out_channels, 1)

    def forward(self, x, compute_dtype=None):
        """``x``: RoI crops (B,3,H,W).  At inference ``compute_dtype`` (default: ``x.dtype``) selects the activation type; fp32
        NCHW crops go through the packed space-to-depth stem, so the caller need not cast / re-layout them first."""
        if self.training or torch.is_grad_enabled():
            return [self.neck(self.trunk(x))]
        t = self.trunk   # inference: BN folded, ReLU / residual add in the convolution epilogue
        dtype = compute_dtype or x.dtype
        c1 = t.conv1
        if (x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0
                and c1.kernel_size == (7, 7) and c1.stride == (2, 2) and c1.padding == (3, 3)):
            x = _stem_s2d(x, c1, t.bn1, dtype)
        else:
            x = _stem(x.to(dtype).contiguous(memory_format=torch.channels_last), c1, t.bn1)
        for layer in (t.layer1, t.layer2, t.layer3, t.layer4):
            for blk in layer:
                idt = x if blk.downsample is None else _conv_bn_act(x, blk.downsample[0], blk.downsample[1], relu=False)
                x = _conv_bn_act(_conv_bn_act(x, blk.conv1, blk.bn1), blk.conv2, blk.bn2, residual=idt)
        return [F.conv2d(x, _cached(self.neck.weight, x.dtype), _cached(self.neck.bias, x.dtype))]


# ----------------------------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------------------------
class PoseNet(nn.Module):
    def __init__(self, cfg: PoseNetConfig | None = None, backbone: nn.Module | None = None):
        super().__init__()
        self.cfg = cfg or PoseNetConfig()
        feature_channel = 1024
        self.backbone = backbone if backbone is not None else ResNet34Backbone(feature_channel)
        self.xyz_nocs_head = TopDownXyzHead(in_dim=feature_channel, xyz_num_classes=1)
        self.size_head = SizeHead(feature_channel, self.cfg.size_head_out_dim, self.cfg.feat_ts)
        if self.cfg.nocsmap_encoder == "conv":   # PoseNet.py:152-157
            self.nocs_encoder = MAPEncoder(3, featdim=256)
        elif self.cfg.nocsmap_encoder == "att":
            self.nocs_encoder = MAPTransformerEncoer()
        else:
            raise NotImplementedError(self.cfg.nocsmap_encoder)
        self.feat_reducer = nn.Conv2d(feature_channel, 256, kernel_size=1)
        self.xyz_deform_head = TopDownXyzHead(in_dim=512, xyz_num_classes=1)
        self.pnp_net = ConvPnPNet(5, featdim=128, rot_dim=4 if "quat" in self.cfg.r_type else 6)
        self.pnp_net.tc_linear = self.cfg.tc_linear
        self.out_res, self.ROT_TYPE, self.TRANS_TYPE, self.Z_TYPE = self.cfg.out_res, self.cfg.r_type, "centroid_z", "REL"
        if "rot6d" not in self.cfg.r_type:
            raise NotImplementedError("GIVEPose's config uses r_type='allo_rot6d' (config/config.py)")
        self.to(memory_format=torch.channels_last)   # 4-D weights in the layout cuDNN consumes for channel-last activations

    @contextlib.contextmanager
    def _precision(self):
        """fp32: TF32 off (the 1e-4 parity mode).  bf16: inference runs on cached bf16 weight copies with bf16 activations
        (no autocast re-casting); under autograd (training step) it is autocast over the fp32 master weights."""
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
        if self.cfg.precision == "fp32":
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.deterministic = True   # parity mode: no split-K / atomic convolution algorithms
        try:
            if self.cfg.precision == "bf16" and torch.is_grad_enabled():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    yield
            else:
                yield
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old

    def _backbone_pipelined(self, img_h, late, dev, dtype):
        """Host-resident RoI crops (the reference uploads them inside ``forward`` too, PoseNet.py:174): chunks of
        ``cfg.h2d_chunk_rois`` RoIs go up on a copy stream while the backbone -- per-RoI independent, unlike the DCNv3
        encoder with its batch-coupled offset rows -- runs on the previous chunk.  ``late`` (mask, 2-D
        coordinates: read only after the heads) follows on the same copy stream; returns the features and
        ``{name: (device tensor, event to wait for)}``."""
        main = torch.cuda.current_stream(dev)
        side = getattr(self, "_copy_stream", None)
        if side is None or side.device != dev:
            side = self._copy_stream = torch.cuda.Stream(dev)
        B, chunk = img_h.shape[0], self.cfg.h2d_chunk_rois
        img_d = torch.empty(img_h.shape, dtype=img_h.dtype, device=dev)
        late_d = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in late.items()}
        side.wait_stream(main)   # the buffers above belong to `main`'s allocation order
        events, late_out = [], {}
        with torch.cuda.stream(side):
            for c0 in range(0, B, chunk):
                img_d[c0:c0 + chunk].copy_(img_h[c0:c0 + chunk], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                events.append(ev)
            for k, v in late.items():
                late_d[k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
                late_out[k] = (late_d[k], ev)
        feats = []
        for i, c0 in enumerate(range(0, B, chunk)):
            main.wait_event(events[i])
            feats.append(self.backbone(img_d[c0:c0 + chunk], dtype)[0])
        return [torch.cat(feats)], late_out

    def forward(self, data, device, do_loss=False, pred_scale=None):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("givepose_b200.PoseNet: Not implemented on the CPU (there is no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        img, mask = data["roi_img"], data["roi_mask_deform" if do_loss else "roi_mask"]
        fast_backbone = isinstance(self.backbone, ResNet34Backbone) and not torch.is_grad_enabled()
        cdtype = torch.bfloat16 if self.cfg.precision == "bf16" else torch.float32
        feat, coord2d, late = None, data["roi_coord_2d"], {}
        if (fast_backbone and not img.is_cuda and img.is_pinned() and not mask.is_cuda and mask.is_pinned()
                and 0 < self.cfg.h2d_chunk_rois < img.shape[0] and img.dtype == torch.float32):
            with self._precision():
                host_late = {"mask": mask}
                if not coord2d.is_cuda and coord2d.is_pinned():
                    host_late["coord2d"] = coord2d
                feat, late = self._backbone_pipelined(img, host_late, dev, cdtype)
        else:
            img = img.to(dev, non_blocking=True)
            mask = mask.to(dev, non_blocking=True)

        def arrived(name, fallback):   # a tensor uploaded on the copy stream: wait for it right before its first use
            if name not in late:
                return fallback.to(dev, non_blocking=True)
            t, ev = late[name]
            torch.cuda.current_stream(dev).wait_event(ev)
            return t

        with self._precision():
            if feat is not None:
                pass
            elif fast_backbone:
                # fp32 NCHW crops straight into the packed stem: no separate cast / channels_last copies
                feat = self.backbone(img.float().contiguous(), cdtype)
            elif self.cfg.precision == "bf16" and not torch.is_grad_enabled():
                # user-supplied backbone (e.g. the reference's timm ConvNeXt-B, network/backbone.py:36-46): its parameters are
                # fp32 masters we hold no bf16 copies of, so it runs under autocast; whatever it returns is cast to bf16
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    feat = self.backbone(img.float().contiguous(memory_format=torch.channels_last))
                feat = [f.to(torch.bfloat16) for f in (feat if isinstance(feat, (list, tuple)) else [feat])]
            else:
                feat = self.backbone(img.contiguous(memory_format=torch.channels_last))
                feat = list(feat) if isinstance(feat, (list, tuple)) else [feat]
            f_nhwc = feat[0].permute(0, 2, 3, 1)                              # (B, 8, 8, 1024) channel-last view
            if not f_nhwc.is_contiguous():
                f_nhwc = f_nhwc.contiguous()
            pred_size = self.size_head(feat).float()
            nocs = self.xyz_nocs_head.forward_nhwc(f_nhwc)                    # (B, 64, 64, 3)
            nocs_feat = self.nocs_encoder.forward_nhwc(nocs)                  # (B, 8, 8, 256)
            conv_feat256 = _conv1x1_rows(f_nhwc, self.feat_reducer)
            ivfc = self.xyz_deform_head.forward_nhwc(torch.cat([conv_feat256, nocs_feat.to(conv_feat256.dtype)], dim=-1))
            coord2d = arrived("coord2d", coord2d).permute(0, 2, 3, 1)
            rot6, t, _ = self.pnp_net.forward_nhwc(torch.cat([ivfc, coord2d.to(ivfc.dtype)], dim=-1))
        mask = arrived("mask", mask)
        # Resize(out_res, NEAREST) of the square mask: src index = floor(dst * in/out) (PoseNet.py:170,180)
        step = mask.shape[-1] // self.out_res
        mask_out = mask[..., ::step, ::step] if mask.shape[-1] % self.out_res == 0 else F.interpolate(mask, size=self.out_res, mode="nearest")
        coor_xyz_nocs = nocs.float().permute(0, 3, 1, 2)
        coor_xyz_ivfc = ivfc.float().permute(0, 3, 1, 2)
        mean_size = data["mean_size"].to(dev)
        pred_size = pred_size + mean_size / mean_size.norm(dim=1).unsqueeze(-1)
        cams = data["cam_K"].to(dev)
        is_allo = "allo" in self.ROT_TYPE
        t = t.float()
        if self.cfg.t_type != "site":
            t = torch.cat([t[:, :2] * 0, t[:, 2:3]], 1)
        args = (cams, data["bbox_center"].to(dev), data["roi_wh"].to(dev), data["resize_ratio"].to(dev))
        if do_loss or torch.is_grad_enabled():
            rot, trans = _pose_decode_torch(rot6.float(), t, *args, is_allo)   # differentiable (train path, :160-249)
        else:
            z_calib = float(cams.reshape(-1, 9)[0, 0]) / 590.0 if self.cfg.dataset == "wild6d" else 1.0
            rot, trans = ops.pose_decode(rot6.float(), t, cams, *args[1:], is_allo=is_allo, z_calib=z_calib)
            if self.cfg.rot_on_cpu:
                rot = rot.cpu()   # one batched D2H instead of the reference's per-RoI loop
        return {"rot": rot, "trans": trans, "size": pred_size, "mask": mask_out, "nocs_coor": coor_xyz_nocs,
                "ivfc_coor": coor_xyz_ivfc}


def _pose_decode_torch(rot6, t, cams, centers, whs, ratios, is_allo, eps=1e-4):
    """Differentiable twin of ``ops.pose_decode`` for the training step: ``pose_from_predictions_train``
    (pose_from_pred_centroid_z.py:160-249) with ``allo_to_ego_mat_torch`` (pose_utils/utils.py:198-229: ray and axis are
    normalised with ``+ eps``, the rotation goes through a quaternion, ``pose_utils.py:348-412``)."""
    x = F.normalize(rot6[..., 0:3], p=2, dim=-1)
    z = F.normalize(torch.cross(x, rot6[..., 3:6], dim=-1), p=2, dim=-1)
    R = torch.stack((x, torch.cross(z, x, dim=-1), z), dim=-1)
    if cams.dim() == 2:
        cams = cams.unsqueeze(0)
    cx = t[:, 0:1] * whs[:, 0:1] + centers[:, 0:1]
    cy = t[:, 1:2] * whs[:, 1:2] + centers[:, 1:2]
    zz = t[:, 2:3] * ratios.view(-1, 1)
    trans = torch.cat([zz * (cx - cams[:, 0:1, 2]) / cams[:, 0:1, 0], zz * (cy - cams[:, 1:2, 2]) / cams[:, 1:2, 1], zz], 1)
    if is_allo:
        ray = trans / (trans.norm(dim=1, keepdim=True) + eps)
        angle = ray[:, 2:3].acos()
        # cross((0, 0, 1), ray) written out: no host-built constant, so the step can be captured in a CUDA graph
        axis = torch.stack((-ray[:, 1], ray[:, 0], torch.zeros_like(ray[:, 2])), dim=1)
        axis = axis / (axis.norm(dim=1, keepdim=True) + eps)
        q = torch.cat([torch.cos(angle / 2.0), axis * torch.sin(angle / 2.0)], dim=1)
        q = q / q.norm(p=2, dim=1, keepdim=True)
        w, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        X, Y, Z = 2 * qx, 2 * qy, 2 * qz
        M = torch.stack([1 - (qy * Y + qz * Z), qx * Y - w * Z, qx * Z + w * Y, qx * Y + w * Z, 1 - (qx * X + qz * Z),
                         qy * Z - w * X, qx * Z - w * Y, qy * Z + w * X, 1 - (qx * X + qy * Y)], dim=1).reshape(-1, 3, 3)
        R = torch.matmul(M, R)
    return R, trans


class GraphedPoseNet:
    """``PoseNet.forward`` for a FIXED number of RoIs captured once as a CUDA graph and replayed (serving a frame's handful of
    detections: at B = 8 the eager forward is ~300 kernel launches of a few microseconds each, i.e. launch-bound).
    ``__call__(data)`` copies the inputs (host or device tensors, any subset of the keys may already be the static buffers)
    into static device buffers on the current stream and replays; the returned tensors are the graph's static outputs --
    consume or clone them before the next call.  ``rot`` stays on the device (one D2H of the caller's choice instead of the
    reference's per-RoI host loop, ``pose_from_pred_centroid_z.py:139-157``)."""

    KEYS = ("roi_img", "roi_mask", "roi_coord_2d", "cam_K", "mean_size", "roi_wh", "bbox_center", "resize_ratio")

    def __init__(self, net: PoseNet, example: dict, device, warmup: int = 2):
        dev = torch.device(device)
        if net.training:
            raise RuntimeError("GraphedPoseNet: call net.eval() first")
        self.net, self.dev = net, dev
        self.static = {k: example[k].to(dev).clone() for k in self.KEYS}
        rot_on_cpu, net.cfg.rot_on_cpu = net.cfg.rot_on_cpu, False   # no D2H inside the capture
        try:
            with torch.no_grad():
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(max(1, warmup)):   # cuDNN algorithm selection, weight caches
                        net(self.static, dev)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self.out = net(self.static, dev)
        finally:
            net.cfg.rot_on_cpu = rot_on_cpu

    def __call__(self, data: dict) -> dict:
        for k in self.KEYS:
            v = data[k]
            if v is not self.static[k]:
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
