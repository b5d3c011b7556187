#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Builds the reference's own DCNv3 CUDA extension for sm_100a into ``oracle/_ref/``.

Sources are compiled where they lie under ``/root/reference/network/ops_dcnv3/src`` (vision.cpp, cpu/dcnv3_cpu.cpp,
cuda/dcnv3_cuda.cu + dcnv3_im2col_cuda.cuh); nothing is copied into this repository.  The reference's own build
(``setup.py:35-46``) refuses to run without a visible GPU and its dispatch macro does not compile against torch 2.11, so
the three translation units are compiled directly with g++ / nvcc; ``oracle/ref_ext/shim_cuda.cu`` re-points the one
macro and #includes the reference .cu verbatim.  The result is the pybind module ``DCNv3_ref`` (same two functions as the
reference's ``DCNv3`` module, ``src/vision.cpp:14-17``) used by the GPU parity tests and as the "reference kernels on the
same B200" timing in bench.py.  ``/root/reference`` does not exist on the GPU box: only the prebuilt .so travels.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/network/ops_dcnv3/src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "DCNv3_ref.so")


def build(force: bool = False) -> str | None:
    """Returns the path of the built module, or None when the reference sources are not present (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return OUT if os.path.exists(OUT) else None
    srcs = [os.path.join(REF_SRC, s) for s in ("vision.cpp", "cpu/dcnv3_cpu.cpp", "cuda/dcnv3_cuda.cu", "cuda/dcnv3_im2col_cuda.cuh", "dcnv3.h")]
    deps = srcs + [os.path.join(HERE, "ref_ext", "shim_cuda.cu"), os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    from torch.utils import cpp_extension as ce
    import torch

    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{REF_SRC}"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-DTORCH_EXTENSION_NAME=DCNv3_ref", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DWITH_CUDA",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-std=c++17", "-O2"]
    objs = []
    procs = []
    for name, src in (("vision", srcs[0]), ("dcnv3_cpu", srcs[1])):
        o = os.path.join(OUT_DIR, name + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["g++", "-fPIC", "-w", *common, *inc, "-c", src, "-o", o]))
    o = os.path.join(OUT_DIR, "dcnv3_cuda.o")
    objs.append(o)
    procs.append(subprocess.Popen(
        ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-w", "--expt-relaxed-constexpr",
         *common, *inc, f'-DGP_REF_CUDA_TU="{srcs[2]}"', "-c", os.path.join(HERE, "ref_ext", "shim_cuda.cu"), "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference DCNv3 extension failed to compile")
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["g++", "-shared", "-o", OUT, *objs, f"-L{libdir}", "-L/usr/local/cuda/lib64", "-lc10", "-lc10_cuda",
                           "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart", f"-Wl,-rpath,{libdir}"])
    for o in objs:
        os.remove(o)
    return OUT


def load():
    """Imports the built module (torch must be imported first).  Raises if it was never built."""
    import importlib.util

    import torch  # noqa: F401

    if not os.path.exists(OUT):
        raise FileNotFoundError(f"{OUT} missing: run oracle/build_ref_ext.py where /root/reference exists")
    spec = importlib.util.spec_from_file_location("DCNv3_ref", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
