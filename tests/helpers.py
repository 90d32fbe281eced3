"""Shared helpers for the test-suite (test infrastructure)."""
import json
import os

import numpy as np
import torch

from oracle import vaesne_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_params(g, dtype=torch.float32):
    shapes = json.loads(str(g["shapes"]))
    return O.random_params(shapes, int(g["seed"]), dtype=dtype)


def golden_x(g, prefix, dtype=torch.float32):
    out = []
    for i in range(4):
        t = torch.from_numpy(g[f"{prefix}.{i}"])
        out.append(t.to(dtype) if t.is_floating_point() else t)
    return tuple(out)


def golden_grads(g):
    return {k[5:]: torch.from_numpy(g[k]) for k in g if k.startswith("grad.")}


def rel_err(a, b):
    """max |a-b| / max |b|  (relative to the tensor's scale; robust to zero entries)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.abs().max().item()
    if den == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / den


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def mm_config(g):
    fam = str(g["family"])
    vaes = [O.VAEConfig("photometry", 4, 4, prior=fam, likelihood=fam, posterior=fam),
            O.VAEConfig("spectra", 4, 4, prior=fam, likelihood=fam, posterior=fam)]
    cfg = O.MMVAEConfig(vaes, beta=float(g["beta"]), prior=fam)
    cfg.apply_scaling()
    return cfg
