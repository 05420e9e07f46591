"""Synthetic workloads named in BASELINE.json (shared by bench.py, __graft_entry__.smoke and the tests).

`li_frames` / `li_table` and `binary_frames` / `binary_table` are the training workloads of configs[2] and configs[4];
`lj_targets` gives them analytic Lennard-Jones energies and forces so that the loss is meaningful.
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


# ------------------------------------------------------------------------------------------- training workloads
LI_LATTICE = 3.51


def li_table(rc: float = 3.54) -> dict:
    """The Li sample's own set (samples/li/input/symm_li_27.json): 12 G2 (eta list, rs = 0) + 2 G4 (zeta 3, lambda -1 / +1)."""
    etas = [0.001, 0.01, 0.03, 0.05, 0.07, 0.1, 0.2, 0.4, 0.5, 0.7, 0.9, 1.0]
    return {"features": [2] * 12 + [4] * 2, "n_features": 14,
            "params": [[rc, e, 0.0] for e in etas] + [[rc, 0.01, 3, -1], [rc, 0.02, 3, 1]]}


def li_frames(n_frames: int, seed: int = 0, sigma: float = 0.05):
    """BASELINE.json configs[2]: bcc Li, 27 atoms per frame (samples/li: Li_crystal_27), jittered; positions (M,27,3) on the
    CPU, cell, pbc.  The block sits inside the cell so that wrap-only PBC does not tear atoms off."""
    a = LI_LATTICE
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    grid = np.stack(np.meshgrid(*[np.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    lat = ((grid + base[None]).reshape(-1, 3) * a)[:27] + 0.25 * a
    gen = torch.Generator().manual_seed(seed)
    pos = torch.from_numpy(lat)[None] + sigma * torch.randn(n_frames, 27, 3, generator=gen, dtype=torch.float64)
    return pos.float(), torch.eye(3) * 3 * a, torch.ones(3)


def binary_table(rc: float = 6.0) -> dict:
    """configs[4]: per species 16 G2 + 16 G5 (rc = 6), the G5 block as `calculate_params` emits it."""
    ang = calculate_params(rc, 0, 2, 0, 0, 16)[2:]
    return {"features": [2] * 16 + [5] * 16, "n_features": 32,
            "params": [[rc, 2.0, float(rs)] for rs in np.linspace(0.5, 5.5, 16)] + [[float(v) for v in r] for r in ang]}


def binary_frames(n_frames: int, seed: int = 0, sigma: float = 0.05):
    """BASELINE.json configs[4]: bcc-ordered binary cell, 32 A + 32 B atoms (4x4x2 cells, A on corners, B on centres)."""
    a = 3.2
    grid = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    A = grid * a + 0.25 * a
    B = (grid + 0.5) * a + 0.25 * a
    gen = torch.Generator().manual_seed(seed)
    pa = torch.from_numpy(A)[None] + sigma * torch.randn(n_frames, 32, 3, generator=gen, dtype=torch.float64)
    pb = torch.from_numpy(B)[None] + sigma * torch.randn(n_frames, 32, 3, generator=gen, dtype=torch.float64)
    cell = torch.diag(torch.tensor([4 * a, 4 * a, 2 * a]))
    return pa.float(), pb.float(), cell.float(), torch.ones(3)


def lj_targets(pos: torch.Tensor, sigma: float, eps: float = 0.05):
    """Analytic Lennard-Jones energy (M,1) and forces (M,N,3) of every frame taken as a free cluster (the reference never
    adds periodic images), then scaled like nnmd/nn/dataset.py:85-87: E -> (E - Emin) / (Emax - Emin),
    F -> F / (Emax - Emin).  pos (M,N,3) on any device; float64 inside."""
    p = pos.double()
    d = p[:, :, None, :] - p[:, None, :, :]
    n = p.shape[1]
    r2 = (d ** 2).sum(-1) + torch.eye(n, dtype=p.dtype, device=p.device)
    s6 = (sigma ** 2 / r2) ** 3 * (1.0 - torch.eye(n, dtype=p.dtype, device=p.device))
    E = (2.0 * eps * (s6 ** 2 - s6)).sum(dim=(1, 2))
    F = ((24.0 * eps * (2.0 * s6 ** 2 - s6) / r2).unsqueeze(-1) * d).sum(dim=2)
    span = E.max() - E.min()
    return ((E - E.min()) / span).unsqueeze(1).float(), (F / span).float()
