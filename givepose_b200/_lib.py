"""ctypes loader for ``libgivepose_b200.so`` (the C ABI of ``include/givepose_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C givepose_b200/csrc``.  There is
deliberately no fallback: if the shared object is missing or does not export the ABI, import fails.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GIVEPOSE_B200_LIB") or os.path.join(_HERE, "lib", "libgivepose_b200.so")   # override: tuning builds only
CSRC = os.path.join(_HERE, "csrc")

GP_F32, GP_BF16, GP_F16, GP_F64 = 0, 1, 2, 3


class DCNv3Desc(ctypes.Structure):
    """``gp_dcnv3_desc`` (include/givepose_b200.h)."""

    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "H", "W", "G", "gc", "kh", "kw", "sh", "sw", "ph", "pw", "dh", "dw", "remove_center", "Ho", "Wo")
                ] + [("offset_scale", ctypes.c_float)]


# every symbol include/givepose_b200.h declares: (restype, argtypes)
_VP, _SZ, _I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
_DP = ctypes.POINTER(DCNv3Desc)
EXPORTS = {
    "gp_abi_version": (_I, []),
    "gp_error_string": (ctypes.c_char_p, [_I]),
    "gp_dcnv3_out_size": (_I, [_I, _I, _I, _I, _I]),
    "gp_dcnv3_forward": (_I, [_VP, _VP, _VP, _VP, _DP, _I, _VP]),
    "gp_dcnv3_forward_softmax": (_I, [_VP, _VP, _VP, _VP, _DP, _I, _VP]),
    "gp_dcnv3_forward_softmax_packed": (_I, [_VP, _VP, _VP, ctypes.c_longlong, _DP, _I, _VP]),
    "gp_dcnv3_backward_workspace": (_SZ, [_DP, _I]),
    "gp_dcnv3_backward": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _VP, _SZ, _DP, _I, _VP]),
    "gp_dcnv3_sample_index": (_I, [_VP, _VP, _VP, _DP, _I, _VP]),
    "gp_dcnv3_forward_host": (_I, [_VP, _VP, _VP, _VP, _SZ, _SZ, _DP, _I, _I]),
    "gp_dcnv3_backward_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _DP, _I, _I]),
    "gp_dcnv3_forward_backward_host": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _DP, _I, _I, _I]),
    "gp_host_cache_release": (_I, []),
    "gp_dwconv3x3_ln_gelu": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, ctypes.c_longlong, ctypes.c_float, _I, _VP]),
    "gp_small_k_linear": (_I, [_VP, _VP, _VP, _VP, ctypes.c_longlong, _I, _I, _I, _VP]),
    "gp_smallk_dwconv3x3_ln_gelu": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, ctypes.c_longlong, ctypes.c_float, _I, _VP]),
    "gp_dcnv3_smallk_fused": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _SZ, _SZ, _DP, _I, _I, _I, _VP]),
    "gp_groupnorm_workspace_floats": (_SZ, [_I, _I, _I, _I]),
    "gp_groupnorm_act": (_I, [_VP, _VP, _VP, _SZ, _VP, _VP, _I, _I, _I, _I, _I, ctypes.c_float, _I, _I, _VP]),
    "gp_groupnorm_apply": (_I, [_VP, _VP, _VP, _SZ, _VP, _VP, _I, _I, _I, _I, _I, ctypes.c_float, _I, _I, _VP]),
    "gp_groupnorm_apply_conv1x1": (_I, [_VP, _VP, _VP, _SZ, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, ctypes.c_float, _I, _I, _I, _VP]),
    "gp_groupnorm_finalize": (_I, [_VP, _VP, _I, _I, _I, ctypes.c_longlong, ctypes.c_float, _VP]),
    "gp_conv3x3_gn_slabs": (_SZ, [_I, _I]),
    "gp_conv3x3_set_pair": (_I, [_I]),
    "gp_conv3x3_gn_bf16_fused_in": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "gp_conv3x3_gn_bf16": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "gp_groupnorm_backward_workspace_floats": (_SZ, [_I, _I, _I, _I, _I]),
    "gp_groupnorm_act_backward": (_I, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "gp_groupnorm_act_conv1x1": (_I, [_VP, _VP, _VP, _SZ, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, ctypes.c_float, _I, _I, _I, _VP]),
    "gp_linear_bf16": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, ctypes.c_float, _VP]),
    "gp_linear_set_pair": (_I, [_I]),
    "gp_mhsa_tokens": (_I, [_VP, _VP, _I, _I, _I, _I, ctypes.c_float, _I, _VP]),
    "gp_stem_s2d_pack": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "gp_upsample_bilinear2x": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "gp_upsample_bilinear2x_backward": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "gp_bias_add_relu": (_I, [_VP, _VP, _VP, ctypes.c_longlong, _I, _I, _VP]),
    "gp_stem_s2d_gemm": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "gp_maxpool3x3s2": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "gp_pose_decode": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, ctypes.c_float, _VP]),
    "gp_roi_affine_inverse": (_I, [_VP, _VP, _I, _I, _VP]),
    "gp_roi_crop": (_I, [_VP, _I, _I, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "gp_resize_linear_u8_normalize": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "gp_set_tuning": (_I, [_I, _I, _I, _I]),
    "gp_set_option": (_I, [_I, _I]),
    "gp_get_option": (_I, [_I]),
    "gp_launch_count": (ctypes.c_uint64, []),
    "gp_launch_count_reset": (None, []),
}


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"givepose_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C givepose_b200/csrc`.  There is no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.gp_abi_version() != 1:
        raise ImportError("givepose_b200: ABI version mismatch between the Python host and the shared library")
    return lib


lib = _load()


def check(code: int, what: str) -> None:
    """Turn a C-ABI return code into the RuntimeError the reference extension would raise."""
    if code != 0:
        raise RuntimeError(f"{what}: {lib.gp_error_string(code).decode()} (code {code})")
