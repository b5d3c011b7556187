"""Generate the DCNv3-core golden vectors from the REFERENCE's own code.

Run once in the build container (``/root/reference`` exists only there):

    python tests/golden/make_golden.py

It imports ``network/ops_dcnv3/functions/dcnv3_func.py`` read-only from ``/root/reference`` (with a dummy
compiled ``DCNv3`` module and a ``DCNv3 1.1`` distribution stub -- the reference file imports both at
module scope, ``dcnv3_func.py:16-19``), evaluates the reference's ``dcnv3_core_pytorch`` (+ autograd) on
seeded inputs and writes ``tests/golden/dcnv3_core.npz``.  Inputs are stored next to the outputs so the
tests never depend on RNG reproducibility.  Nothing at test/bench time reads ``/root/reference``.

Cases mirror ``network/ops_dcnv3/test.py`` (fixture ``:19-40``: N=2, 8x8, M=4, D=16, 3x3, offset_scale 2,
pad 1, stride 1; backward channel sweep ``:255-264``) plus the configurations the reference's tests do
not reach: stride 2 through the flat-slice adapter, remove_center, dilation 2,
non-square images, a 5x5 kernel.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/network/ops_dcnv3/functions/dcnv3_func.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.modules.setdefault("DCNv3", types.ModuleType("DCNv3"))
    import pkg_resources

    real = pkg_resources.get_distribution

    def fake(name):
        if name == "DCNv3":
            return types.SimpleNamespace(version="1.1")
        return real(name)

    pkg_resources.get_distribution = fake
    spec = importlib.util.spec_from_file_location("ref_dcnv3_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def out_size(size, k, s, p, d):
    return (size + 2 * p - (d * (k - 1) + 1)) // s + 1


def make_inputs(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full_res_offsets=False):
    """dist 'T' = the reference test distribution (test.py:36-40); 'M' = model-like."""
    P = k * k - rc
    Ho, Wo = out_size(H, k, s, pad, dil), out_size(W, k, s, pad, dil)
    Hm, Wm = (H, W) if full_res_offsets else (Ho, Wo)
    if dist == "T":
        inp = torch.rand(N, H, W, G * gc, generator=gen) * 0.01
        off = torch.rand(N, Hm, Wm, G * P * 2, generator=gen) * 10
        m = torch.rand(N, Hm, Wm, G, P, generator=gen) + 1e-5
        m = m / m.sum(-1, keepdim=True)
    else:
        inp = torch.randn(N, H, W, G * gc, generator=gen)
        off = torch.randn(N, Hm, Wm, G * P * 2, generator=gen) * 1.5
        m = torch.softmax(torch.randn(N, Hm, Wm, G, P, generator=gen), -1)
    m = m.reshape(N, Hm, Wm, G * P)
    gout = torch.randn(N, Ho, Wo, G * gc, generator=gen)
    return inp, off, m, gout, Ho, Wo


CASES = [
    # name,            N, H,  W,  G, gc, k, s, pad, dil, scale, rc, dist, full_res
    ("test_py_fixture", 2, 8, 8, 4, 16, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc1", 2, 6, 6, 2, 1, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc16", 2, 6, 6, 2, 16, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc30", 2, 6, 6, 2, 30, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc32", 2, 6, 6, 2, 32, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc64", 1, 6, 6, 2, 64, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("bwd_gc71", 1, 6, 6, 2, 71, 3, 1, 1, 1, 2.0, 0, "T", False),
    ("model_like_s1", 1, 7, 6, 8, 32, 3, 1, 1, 1, 1.0, 0, "M", False),
    ("stride2_flat_slice", 4, 8, 8, 2, 16, 3, 2, 1, 1, 1.0, 0, "M", True),
    ("stride2_odd", 4, 6, 10, 2, 4, 3, 2, 1, 1, 1.0, 0, "T", True),
    ("remove_center", 2, 6, 6, 4, 8, 3, 1, 1, 1, 1.0, 1, "M", False),
    ("dilation2", 1, 9, 11, 2, 16, 3, 1, 2, 2, 1.0, 0, "M", False),
    ("kernel5", 1, 10, 10, 2, 8, 5, 1, 2, 1, 1.0, 0, "M", False),
    ("scale_non_pow2", 2, 8, 8, 2, 16, 3, 1, 1, 1, 0.7, 0, "T", False),
]


def main():
    ref = load_reference()
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.dcnv3 import flat_slice

    gen = torch.Generator().manual_seed(3)  # test.py:31
    blob = {}
    meta = []
    for (name, N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full) in CASES:
        inp, off, m, gout, Ho, Wo = make_inputs(gen, N, H, W, G, gc, k, s, pad, dil, rc, dist, full)
        res = {}
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            i_ = inp.to(dt).clone().requires_grad_(True)
            o_ = off.to(dt).clone().requires_grad_(True)
            m_ = m.to(dt).clone().requires_grad_(True)
            o_use, m_use = (flat_slice(o_, N, Ho, Wo), flat_slice(m_, N, Ho, Wo)) if full else (o_, m_)
            out = ref.dcnv3_core_pytorch(i_, o_use, m_use, k, k, s, s, pad, pad, dil, dil, G, gc, scale, rc)
            assert out.shape == (N, Ho, Wo, G * gc), out.shape
            (out * gout.to(dt)).sum().backward()
            res[tag] = (out.detach(), i_.grad, o_.grad, m_.grad)
        blob[f"{name}/input"] = inp.numpy()
        blob[f"{name}/offset"] = off.numpy()
        blob[f"{name}/mask"] = m.numpy()
        blob[f"{name}/grad_output"] = gout.numpy()
        # f64 results are the pin; of the f32 run only the output is kept (fixture size)
        out, gi, go, gm = res["f64"]
        blob[f"{name}/out_f64"] = out.numpy()
        blob[f"{name}/grad_input_f64"] = gi.numpy()
        blob[f"{name}/grad_offset_f64"] = go.numpy()
        blob[f"{name}/grad_mask_f64"] = gm.numpy()
        blob[f"{name}/out_f32"] = res["f32"][0].numpy()
        meta.append(f"{name}:{N},{H},{W},{G},{gc},{k},{s},{pad},{dil},{scale},{rc},{dist},{int(full)}")
        print(name, "ok", tuple(res["f64"][0].shape))
    blob["__cases__"] = np.array(meta)
    blob["__provenance__"] = np.array(
        ["reference dcnv3_core_pytorch (network/ops_dcnv3/functions/dcnv3_func.py:172-220), torch " + torch.__version__])
    path = os.path.join(HERE, "dcnv3_core.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
