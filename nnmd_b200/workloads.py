"""Synthetic workloads named in BASELINE.json (shared by bench.py, __graft_entry__.smoke and the tests).

`bench_table()` is the literal 96-row parameter table of the 1M-atom config: 32 G2 rows given in
g2_function order [rc, eta, rs] with eta = 4.0 and rs = linspace(0.5, 5.5, 32), followed by the 64 G4 rows that
`calculate_params(6.0, 0, 32, 0, 64, 0)` emits (eta = 0.127921, zeta_k = 1 + k*30/62, lambda = +-1).
"""
from __future__ import annotations

import numpy as np
import torch

from .features.auto_params import calculate_params

CU_LATTICE = 3.615
_FCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])


def bench_table(rc: float = 6.0, n_g2: int = 32, n_g4: int = 64) -> dict:
    radial = [[rc, 4.0, float(rs)] for rs in np.linspace(0.5, rc - 0.5, n_g2)]
    angular = calculate_params(rc, 0, 32, 0, n_g4, 0)[32:]
    angular = [[float(r[0]), float(r[1]), float(r[2]), float(r[3])] for r in angular]
    return {"features": [2] * n_g2 + [4] * n_g4, "params": radial + angular, "n_features": n_g2 + n_g4}


def fcc_cell(ncell: int, a: float = CU_LATTICE, sigma: float = 0.05, seed: int = 0, dtype=torch.float32):
    """Jittered fcc block of ncell^3 conventional cells: positions (4*ncell^3, 3) on the CPU, cell (3,3), pbc (3,).

    The jitter comes from a CPU torch.Generator seeded with `seed`, drawn in float64 and rounded once.
    """
    grid = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    pos = torch.from_numpy(((grid + _FCC_BASIS[None]).reshape(-1, 3) * a))
    gen = torch.Generator().manual_seed(seed)
    pos = pos + sigma * torch.randn(pos.shape, generator=gen, dtype=torch.float64)
    cell = torch.eye(3, dtype=torch.float64) * (ncell * a)
    return pos.to(dtype), cell.to(dtype), torch.ones(3, dtype=dtype)


BENCH_HIDDEN = [64, 64]


def bench_net(nf: int, device, dtype=torch.float32):
    """The atomic network of the bench workload: nf-64-64-1, the reference's initialisation (bpnn.py:149-157), seed 0."""
    from .nn import AtomicNN
    torch.manual_seed(0)
    net = AtomicNN(nf, list(BENCH_HIDDEN))
    for layer in net.model:
        torch.nn.init.xavier_uniform_(layer.weight)
        torch.nn.init.constant_(layer.bias, 0)
    return net.to(device=device, dtype=dtype)
