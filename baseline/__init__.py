"""Reference arm plumbing (bench.py --impl reference, drop-in tests): stages and loads the reference's OWN Python files.

Nothing here is product code: ``givepose_b200/`` never imports it (tests/test_host_abi.py enforces that for ``oracle`` and
``baseline`` alike).  ``baseline/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the staged files travel to the GPU box like ``oracle/_ref/DCNv3_ref.so`` does.
"""
