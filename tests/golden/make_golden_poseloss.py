"""Generate ``tests/golden/poseloss.npz`` from the REFERENCE's own ``losses/pose_loss.py::PoseLoss`` (build container only;
``ref_loader`` supplies the leaf stubs).  Inputs: ``givepose_b200.loss.make_loss_inputs`` + ``oracle.pose_loss.make_predictions``
(seeded).  Stored per case: the six loss terms and the gradient of their sum w.r.t. every prediction tensor.

    python tests/golden/make_golden_poseloss.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402

CASES = {"sym3_B12": dict(B=12, seed=0, sym_every=3), "nosym_B5": dict(B=5, seed=1, sym_every=0), "allsym_B4": dict(B=4, seed=2, sym_every=1)}


def main():
    ref_loader.install_stubs(None)
    ref_loader.load_posenet_module()
    import losses.pose_loss as PL
    from givepose_b200.loss import make_loss_inputs   # plain torch, importable without the CUDA library? (needs the built .so)
    from oracle.pose_loss import make_predictions
    out = {}
    for name, c in CASES.items():
        data = make_loss_inputs(c["B"], c["seed"], c["sym_every"])
        pred = {k: v.clone().requires_grad_(True) for k, v in make_predictions(c["B"], c["seed"]).items()}
        loss = PL.PoseLoss()(pred, {k: v.clone() for k, v in data.items()})
        sum(loss.values()).backward()
        for k, v in loss.items():
            out[f"{name}/{k}"] = v.detach().numpy()
        for k, v in pred.items():
            out[f"{name}/grad_{k}"] = v.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "poseloss.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
