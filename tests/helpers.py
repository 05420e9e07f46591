"""Shared helpers for the parity tests (fixture loading, error measures)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    feats = [int(k) for k in d["features"]]
    params = []
    for row, ints in zip(d["params"], d["params_is_int"]):
        p = []
        for v, is_int in zip(row, ints):
            if np.isnan(v):
                continue
            p.append(int(v) if is_int else float(v))
        params.append(p)
    sfd = {"features": feats, "params": params, "n_features": len(feats)}
    return d, sfd


def relerr(a, b):
    """max |a-b| / max |b|  (normwise, per tensor)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def col_relerr(a, b, dim_f):
    """worst normwise relative error over feature columns."""
    a = torch.as_tensor(a, dtype=torch.float64).movedim(dim_f, 0).flatten(1)
    b = torch.as_tensor(b, dtype=torch.float64).movedim(dim_f, 0).flatten(1)
    scale = b.abs().max(dim=1).values.clamp_min(1e-300)
    return float(((a - b).abs().max(dim=1).values / scale).max())


SF_CASES = ["sf_cluster", "sf_cu111_json", "sf_cu111_auto", "sf_fcc108", "sf_lj_trimer"]
