"""``install()`` puts the ``DCNv3`` drop-in module (and its ``DCNv3-1.1.dist-info``) on ``sys.path``."""
import os
import sys


def install() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    # the reference asks ``pkg_resources.get_distribution('DCNv3')`` (functions/dcnv3_func.py:17-19); a pkg_resources that was
    # imported before this call has already scanned sys.path, so tell its working set about the new entry
    pr = sys.modules.get("pkg_resources")
    if pr is not None:
        pr.working_set.add_entry(here)
    return here
