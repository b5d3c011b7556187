"""Drop-in for the reference's compiled extension module ``DCNv3``.

The reference does ``import DCNv3`` and calls ``DCNv3.dcnv3_forward(*args)`` / ``DCNv3.dcnv3_backward(*args)``
(``network/ops_dcnv3/functions/dcnv3_func.py:16,53,74``; pybind definitions in ``src/vision.cpp:14-17``), and
reads ``pkg_resources.get_distribution('DCNv3').version`` (``dcnv3_func.py:18-19``; ``setup.py:63-64`` says 1.1).
Putting this directory on ``sys.path`` (``givepose_b200.dropin.install()``) provides both: this module and the
``DCNv3-1.1.dist-info`` next to it, so the reference's ``dcnv3_func.py`` / ``modules/dcnv3.py`` / ``PoseNet.py``
run unchanged on the B200 kernels.
"""
from givepose_b200.functions import dcnv3_backward, dcnv3_forward  # noqa: F401

__version__ = "1.1"
