"""Loader for tests/golden/dcnv3_core.npz (written by tests/golden/make_golden.py from the reference)."""
import os
from dataclasses import dataclass

import numpy as np
import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dcnv3_core.npz")


@dataclass
class Case:
    name: str
    N: int
    H: int
    W: int
    G: int
    gc: int
    k: int
    s: int
    pad: int
    dil: int
    scale: float
    rc: int
    dist: str
    full_res: bool
    arrays: dict

    def t(self, key, dtype=None):
        x = torch.from_numpy(self.arrays[key])
        return x if dtype is None else x.to(dtype)

    @property
    def args(self):
        """(kh, kw, sh, sw, ph, pw, dh, dw, group, group_channels, offset_scale) -- reference arg order."""
        return (self.k, self.k, self.s, self.s, self.pad, self.pad, self.dil, self.dil, self.G, self.gc, self.scale)

    @property
    def out_hw(self):
        f = lambda n: (n + 2 * self.pad - (self.dil * (self.k - 1) + 1)) // self.s + 1
        return f(self.H), f(self.W)


def load_cases():
    z = np.load(_PATH)
    cases = []
    for line in z["__cases__"]:
        name, rest = str(line).split(":")
        N, H, W, G, gc, k, s, pad, dil, scale, rc, dist, full = rest.split(",")
        arrays = {key.split("/", 1)[1]: z[key] for key in z.files if key.startswith(name + "/")}
        cases.append(Case(name, int(N), int(H), int(W), int(G), int(gc), int(k), int(s), int(pad), int(dil),
                          float(scale), int(rc), dist, bool(int(full)), arrays))
    return cases


CASES = load_cases()
CASE_IDS = [c.name for c in CASES]
