"""``install()`` puts the ``DCNv3`` drop-in module (and its ``DCNv3-1.1.dist-info``) on ``sys.path``."""
import os
import sys


def install() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    return here
