"""Shared test helpers: golden fixture loading and oracle glue."""
import glob
import os
import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("sasrec_"):-4] for p in glob.glob(os.path.join(GOLDEN, "sasrec_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"sasrec_{name}.npz"))
    g = {k: z[k] for k in z.files}
    B, L, H, nh, nl, I = [int(v) for v in g["cfg"]]
    g["dims"] = dict(B=B, L=L, H=H, nh=nh, nl=nl, I=I)
    return g


def sd_from(g, prefix="sd0/", dtype=torch.float32, requires_grad=False):
    sd = {}
    for k, v in g.items():
        if k.startswith(prefix):
            t = torch.from_numpy(np.array(v)).to(dtype)
            if requires_grad:
                t.requires_grad_(True)
            sd[k[len(prefix):]] = t
    return sd


def ids(g):
    return [torch.from_numpy(g[k]).long() for k in ("seq", "dec", "pos", "neg")]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
