"""Loader for tests/golden/*.npz (written by oracle/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.cfg = json.loads(str(z["cfg"]))
        self.sd, self.inp, self.out = {}, {}, {}
        for k in z.files:
            if k == "cfg":
                continue
            grp, key = k.split("/", 1)
            v = torch.from_numpy(z[k])
            {"sd": self.sd, "in": self.inp, "out": self.out}[grp][key] = v


def max_abs(a, b):
    return (a.double() - b.double()).abs().max().item()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
