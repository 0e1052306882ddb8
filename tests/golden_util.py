"""Load tests/golden/*.npz (written by oracle/gen_golden.py from the reference's
own in-tree kernels) back into torch tensors."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_BF16_KEYS = ("q", "k_buf", "v_buf", "k_new", "v_new")


def golden_names(kind):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if name.startswith(kind):
            out.append(name)
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {}
    for k in z.files:
        a = z[k]
        if k in _BF16_KEYS:
            d[k] = torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16)
        elif a.ndim == 0:
            d[k] = a.item()
        else:
            d[k] = torch.from_numpy(a)
    return d
