"""Leaf stubs for the third-party packages the reference imports but this image lacks: mmcv,
detectron2, timm, transforms3d (with a real ``axangle2mat`` -- it is on the test-time pose path, pose_utils/utils.py:58),
open3d, ipdb, termcolor, matplotlib, plus the NumPy-2 shim for ``pose_utils/RT_transform.py:297``.  The reference's own files
are imported unchanged on top of these.  Test / reference-arm infrastructure only."""
import math
import sys
import types

import numpy as np
import torch.nn as nn


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m

def install(dcnv3_forward, dcnv3_backward=None):
    # numpy 2 shim (pose_utils/RT_transform.py:297)
    if not hasattr(np, "maximum_sctype"):
        np.maximum_sctype = lambda t: np.float64
    # compiled ext + dist version
    _mod("DCNv3", dcnv3_forward=dcnv3_forward, dcnv3_backward=dcnv3_backward)
    import pkg_resources
    real = pkg_resources.get_distribution
    pkg_resources.get_distribution = lambda n: types.SimpleNamespace(version="1.1") if n == "DCNv3" else real(n)
    # mmcv
    def normal_init(m, mean=0, std=1, bias=0):
        if hasattr(m, "weight") and m.weight is not None: nn.init.normal_(m.weight, mean, std)
        if hasattr(m, "bias") and m.bias is not None: nn.init.constant_(m.bias, bias)
    def constant_init(m, val, bias=0):
        if hasattr(m, "weight") and m.weight is not None: nn.init.constant_(m.weight, val)
        if hasattr(m, "bias") and m.bias is not None: nn.init.constant_(m.bias, bias)
    def kaiming_init(m, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
        if hasattr(m, "weight") and m.weight is not None:
            nn.init.kaiming_normal_(m.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if hasattr(m, "bias") and m.bias is not None: nn.init.constant_(m.bias, bias)
    class Registry(dict):
        def register_module(self, name=None, module=None, force=False):
            if module is not None:
                self[name] = module; return module
            def deco(cls):
                self[name or cls.__name__] = cls; return cls
            return deco
    CONV_LAYERS = Registry(); CONV_LAYERS["Conv2d"] = nn.Conv2d; CONV_LAYERS["Conv"] = nn.Conv2d
    def build_conv_layer(cfg, *a, **k):
        if cfg is None: return nn.Conv2d(*a, **k)
        cfg = dict(cfg); return CONV_LAYERS[cfg.pop("type")](*a, **k, **cfg)
    def build_padding_layer(cfg, *a, **k): raise NotImplementedError
    mm = _mod("mmcv", Config=dict)
    cnn = _mod("mmcv.cnn", normal_init=normal_init, constant_init=constant_init, kaiming_init=kaiming_init,
               build_conv_layer=build_conv_layer, build_padding_layer=build_padding_layer, CONV_LAYERS=CONV_LAYERS)
    _mod("mmcv.cnn.utils", normal_init=normal_init, constant_init=constant_init, kaiming_init=kaiming_init)
    _mod("mmcv.cnn.bricks"); _mod("mmcv.cnn.bricks.conv", build_conv_layer=build_conv_layer, CONV_LAYERS=CONV_LAYERS)
    _mod("mmcv.cnn.bricks.padding", build_padding_layer=build_padding_layer)
    _mod("mmcv.runner", obj_from_dict=lambda *a, **k: None)
    mm.cnn = cnn
    # detectron2
    _mod("detectron2"); _mod("detectron2.layers"); _mod("detectron2.utils")
    _mod("detectron2.layers.batch_norm", BatchNorm2d=nn.BatchNorm2d, FrozenBatchNorm2d=nn.BatchNorm2d,
         NaiveSyncBatchNorm=nn.BatchNorm2d)
    _mod("detectron2.utils.env", TORCH_VERSION=(2, 11))
    # timm (import-only on the default path)
    class _Dummy(nn.Module):
        def __init__(self, *a, **k): super().__init__()
    def _cfg(**k): return k
    t = _mod("timm", create_model=lambda *a, **k: None, list_modules=lambda *a, **k: [])
    tm = _mod("timm.models"); 
    lay = dict(StdConv2d=nn.Conv2d, DropPath=_Dummy, to_2tuple=lambda x: (x, x), trunc_normal_=nn.init.trunc_normal_, Mlp=_Dummy)
    _mod("timm.models.layers", **lay); _mod("timm.layers", **lay)
    _mod("timm.models.registry", register_model=lambda f: f)
    _mod("timm.models.vision_transformer", Block=_Dummy, _cfg=_cfg, Mlp=_Dummy, PatchEmbed=_Dummy, Attention=_Dummy)
    # transforms3d: axangle2mat must be real
    def axangle2mat(axis, angle, is_normalized=False):
        x, y, z = axis
        if not is_normalized:
            n = math.sqrt(x * x + y * y + z * z); x, y, z = x / n, y / n, z / n
        c, s = math.cos(angle), math.sin(angle); C = 1 - c
        xs, ys, zs = x * s, y * s, z * s
        xC, yC, zC = x * C, y * C, z * C
        xyC, yzC, zxC = x * yC, y * zC, z * xC
        return np.array([[x * xC + c, xyC - zs, zxC + ys], [xyC + zs, y * yC + c, yzC - xs], [zxC - ys, yzC + xs, z * zC + c]])
    any_fn = lambda *a, **k: None
    _mod("transforms3d")
    _mod("transforms3d.quaternions", mat2quat=any_fn, quat2mat=any_fn, qmult=any_fn, quat2axangle=any_fn, axangle2quat=any_fn, qinverse=any_fn)
    _mod("transforms3d.axangles", axangle2mat=axangle2mat, mat2axangle=any_fn)
    _mod("transforms3d.euler", _AXES2TUPLE={}, _NEXT_AXIS=[1, 2, 0, 1], _TUPLE2AXES={}, euler2mat=any_fn, mat2euler=any_fn, euler2quat=any_fn, quat2euler=any_fn)
    for n in ("open3d", "ipdb", "termcolor", "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "cv2"):
        if n == "cv2":
            try:
                import cv2  # noqa
                continue
            except Exception:
                pass
        _mod(n, colored=lambda s, *a, **k: s, set_trace=any_fn)
