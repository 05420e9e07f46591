"""Dense helpers of the reference's `nnmd.features` surface (nnmd/features/symm_funcs.py).

These take a dense distance matrix and exist for API compatibility with the plotting samples
(samples/symm_funcs_plots/plot.py:75,92); they are thin torch code and are NOT the product path.
`calculate_sf` (the hot path) never calls them: it goes through the CUDA kernels in csrc/.
Semantics follow SURVEY.md section 0 (wrap-only PBC, closed-triangle mask, |cos| angular term).
"""
from __future__ import annotations

import math
from enum import Enum

import torch


class SymmetryFunction(Enum):
    """Implemented symmetry-function kinds (reference: symm_funcs.py:6-15)."""

    G1 = 1
    G2 = 2
    G4 = 4
    G5 = 5


def _inside(distances: torch.Tensor, cutoff: float) -> torch.Tensor:
    return (distances > 0) & (distances < cutoff)


def map_positions_pbc(positions: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """Wrap-only periodic boundaries (reference: symm_funcs.py:53-82)."""
    if torch.det(cell) == 0:
        return positions
    frac = positions @ torch.linalg.inv(cell)
    frac = frac - torch.floor(frac) * pbc.to(frac.dtype)
    return frac @ cell


def calculate_distances(positions: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """Dense (B,N,N) distances of the wrapped positions, no minimum image (reference: symm_funcs.py:85-104)."""
    p = map_positions_pbc(positions, cell, pbc)
    return torch.linalg.vector_norm(p.unsqueeze(-3) - p.unsqueeze(-2), dim=-1)


def f_cutoff(r: torch.Tensor, cutoff: float) -> torch.Tensor:
    """Cosine cutoff (reference: symm_funcs.py:107-120)."""
    inside = r < cutoff
    return torch.where(inside, 0.5 * torch.cos(r * (math.pi / cutoff)) + 0.5, torch.zeros_like(r))


def g1_function(r: torch.Tensor, cutoff: float) -> torch.Tensor:
    """G1 (reference: symm_funcs.py:123-140)."""
    return (f_cutoff(r, cutoff) * _inside(r, cutoff)).sum(dim=-1)


def g2_function(r: torch.Tensor, cutoff: float, eta: float, rs: float) -> torch.Tensor:
    """G2 (reference: symm_funcs.py:143-162); argument ORDER is the reference's."""
    gauss = torch.exp(-eta * (r - rs) ** 2)
    return (gauss * f_cutoff(r, cutoff) * _inside(r, cutoff)).sum(dim=-1)


def _angular(distances, cutoff, eta, zeta, lambd, with_jk_in_exponent):
    ok = _inside(distances, cutoff)
    fc = f_cutoff(distances, cutoff) * ok
    d_ij, d_ik, d_jk = distances.unsqueeze(-1), distances.unsqueeze(-2), distances.unsqueeze(-3)
    closed = ok.unsqueeze(-1) & ok.unsqueeze(-2) & ok.unsqueeze(-3)
    two_ab = 2.0 * d_ij * d_ik
    cos_t = torch.where(two_ab >= 1e-3, (d_ij**2 + d_ik**2 - d_jk**2) / two_ab.clamp(min=1e-8),
                        torch.zeros_like(two_ab * d_jk))
    absc = cos_t.abs()
    radial = d_ij**2 + d_ik**2 + (d_jk**2 if with_jk_in_exponent else 0.0)
    powered = torch.pow((1.0 + lambd * absc).clamp(min=0.0), zeta)
    term = torch.where(absc > 1e-3, powered, torch.zeros_like(powered))
    term = term * torch.exp(-eta * radial) * fc.unsqueeze(-1) * fc.unsqueeze(-2) * fc.unsqueeze(-3)
    term = torch.where(closed, term, torch.zeros_like(term))
    return 2.0 ** (1.0 - zeta) * term.sum(dim=(-2, -1))


def g4_function(distances: torch.Tensor, cutoff: float, eta: float, zeta: float, lambd: float) -> torch.Tensor:
    """G4 with the reference's closed-triangle mask and |cos| factor (reference: symm_funcs.py:165-221)."""
    return _angular(distances, cutoff, eta, zeta, lambd, True)


def g5_function(distances: torch.Tensor, cutoff: float, eta: float, zeta: float, lambd: float) -> torch.Tensor:
    """G5; keeps fc(r_jk) and the r_jk < rc mask like the reference (symm_funcs.py:224-277)."""
    return _angular(distances, cutoff, eta, zeta, lambd, False)
