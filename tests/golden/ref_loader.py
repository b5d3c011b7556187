"""Loads the reference PoseNet from /root/reference with leaf stubs (build container only).  The stub installer lives in
``baseline/stubs.py`` (shared with the reference arm of bench.py)."""
import os
import sys

REF = "/root/reference"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from baseline.stubs import install as install_stubs  # noqa: E402,F401


def load_posenet_module():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import config.config  # noqa
    import absl.flags as flags
    if not flags.FLAGS.is_parsed():
        flags.FLAGS(["x"])
    import network.PoseNet as PN
    return PN
