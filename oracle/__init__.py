"""CPU oracle for the GIVEPose DCNv3 + PoseNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``givepose_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do.  The product path fails loudly when its CUDA library is missing -- it never falls back here.

Contents
--------
``dcnv3_oracle.c`` / ``dcnv3_oracle_impl.h``
    C restatement (f32 + f64) of the reference CUDA kernels' arithmetic, in the reference's operation
    order (``network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh``).  Pins indices / bounds bit-exactly and
    values for every configuration, including the stride-2 flat-offset addressing that the reference's
    PyTorch path cannot express.
``dcnv3.py``
    ctypes wrapper around the C oracle + ``dcnv3_core_torch``: a torch/grid_sample restatement of the
    reference's ``dcnv3_core_pytorch`` (``network/ops_dcnv3/functions/dcnv3_func.py:172-220``), used as
    the multi-threaded CPU baseline ("port") and as a second, independently-derived value check.
``posenet.py``
    plain-PyTorch CPU restatement of the reference ``PoseNet.forward`` (``network/PoseNet.py:173-231``).

Pinning: the golden vectors in ``tests/golden/`` were produced by importing the *reference's own*
``dcnv3_core_pytorch`` / ``PoseNet`` from ``/root/reference`` in the build container
(``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every oracle function
against them.
"""
