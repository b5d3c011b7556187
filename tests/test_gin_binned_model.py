"""CPU check of the binned grad_input ALGORITHM (window, counting sort by footprint cell, 2x2 destination blocks fed by
per-cell-row runs, per-pixel bounds check at flush, fallback scatter outside the window): the NumPy model of
givepose_b200/csrc/dcnv3_gin_binned.cuh against the C oracle's grad_input.  The CUDA kernel itself runs in the -m gpu tests."""
import numpy as np
import pytest
import torch

from gin_binned_model import grad_input_binned


@pytest.fixture(scope="module")
def O():
    from oracle import dcnv3
    dcnv3.build()
    return dcnv3


CASES = [  # N, H, W, G, gc, k, s, pad, dil, scale, rc, offset std / kind, full-res offsets
    (1, 20, 20, 2, 4, 3, 1, 1, 1, 1.0, 0, "T", False),
    (1, 20, 20, 2, 4, 3, 1, 1, 1, 1.0, 0, "M", False),
    (2, 16, 16, 1, 4, 3, 2, 1, 1, 1.0, 0, "M", True),      # stride 2, flat prefix of full-resolution offsets
    (1, 12, 12, 1, 4, 3, 1, 1, 1, 2.0, 1, "T", False),     # remove_center
    (1, 50, 50, 1, 4, 3, 1, 1, 1, 1.0, 0, "far", False),   # window overflow -> fallback scatter
    (1, 13, 11, 1, 4, 5, 1, 2, 1, 1.0, 0, "T", False),     # 5x5, ragged tiles
    (1, 13, 11, 1, 4, 3, 1, 0, 2, 0.7, 0, "M", False),     # no padding, dilation 2
]


@pytest.mark.parametrize("row_walk", [False, True], ids=["blocks(gin_binned)", "row_walk(bwd_fused)"])
@pytest.mark.parametrize("spec", CASES)
def test_binned_model_matches_oracle(O, spec, row_walk):
    N, H, W, G, gc, k, s, pad, d, scale, rc, dist, full = spec
    gen = torch.Generator().manual_seed(3)
    Ho, Wo = O.out_size(H, k, s, pad, d), O.out_size(W, k, s, pad, d)
    P = k * k - rc
    Hm, Wm = (H, W) if full else (Ho, Wo)
    inp = torch.randn(N, H, W, G * gc, generator=gen)
    if dist == "T":
        off = torch.rand(N, Hm, Wm, G * P * 2, generator=gen) * 10
    elif dist == "far":
        off = torch.randn(N, Hm, Wm, G * P * 2, generator=gen) * 30
    else:
        off = torch.randn(N, Hm, Wm, G * P * 2, generator=gen)
    m = torch.softmax(torch.randn(N, Hm, Wm, G, P, generator=gen), -1).reshape(N, Hm, Wm, G * P)
    gout = torch.randn(N, Ho, Wo, G * gc, generator=gen)
    args = (k, k, s, s, pad, pad, d, d, G, gc, scale)
    gi, _, _ = O.backward(inp, off, m, gout, *args, rc)
    hw, fl = O.index(off, N, H, W, k, k, s, s, pad, pad, d, d, G, scale, rc)
    st = {}
    mine = grad_input_binned(off.numpy(), m.numpy(), gout.numpy(), N, H, W, G, gc, k, k, s, s, pad, pad, d, d, scale, rc,
                             Ho, Wo, hw.numpy(), fl.numpy(), stats=st, row_walk=row_walk)
    assert np.abs(mine - gi.numpy()).max() / np.abs(gi.numpy()).max() < 2e-5
    assert st["flush_lines"] < st["corner_lines"]          # the point of the exercise: fewer reductions than corners
    if dist == "far":
        assert st.get("overflow", 0) > 0
