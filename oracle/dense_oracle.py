"""Dense CPU oracle (O2) for the nnmd descriptor -> energy -> force path.

TEST INFRASTRUCTURE ONLY.  Nothing in `nnmd_b200/` imports this file; it is the
checker used by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.

It restates, with dense O(N^2)/O(N^3) tensors and `torch.autograd`, the algorithm
of the reference (ded-otshelnik/nnmd) exactly as implemented, including its
non-standard semantics:

  * wrap-only periodic boundaries, no minimum image   (symm_funcs.py:53-82)
  * pair set  {0 < d_ij < rc}                          (symm_funcs.py:18-31)
  * triplets need d_ij, d_ik AND d_jk inside (0, rc)   (symm_funcs.py:34-50)
  * angular factor (1 + lambda*|cos|)^zeta, zero when |cos| <= 1e-3, cos forced
    to zero when 2*d_ij*d_ik < 1e-3                    (symm_funcs.py:193-211)
  * G5 keeps fc(d_jk)                                  (symm_funcs.py:246-260)
  * `dg` = d(sum_i G_i)/d r_a, NOT a Jacobian          (calculate_sf.py:85-90)
  * g is L2-normalised over the feature axis, dg is not (calculate_sf.py:100-104)
  * force = -sum_f dE/dg_hat[a,f] * dg[a,f,:]          (bpnn.py:347-354)

Parity pinning: the reference has no golden vectors of its own,
so this oracle is pinned against outputs of the reference itself, executed in
the build container by tests/golden/make_golden.py and committed as fixtures
(tests/test_oracle_golden.py).

`pow_mode`:
  "exact"      (1+lambda|c|)^zeta in the working dtype            (oracle O2)
  "complex64"  round the base to fp32 and take the power in complex64, as the
               reference does even for fp64 inputs (symm_funcs.py:207-209);
               used only to pin this file against the reference at ~1e-12.
"""
from __future__ import annotations

import math

import torch

G1, G2, G4, G5 = 1, 2, 4, 5


def wrap_positions(pos: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """symm_funcs.py:53-82: fractional coords, subtract floor on periodic axes, back."""
    if float(torch.det(cell)) == 0.0:
        return pos
    inv = torch.linalg.inv(cell)
    frac = pos @ inv
    frac = frac - torch.floor(frac) * pbc.to(frac.dtype)
    return frac @ cell


def distance_matrix(pos: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """symm_funcs.py:85-104: d[b,i,j] = |p_j - p_i| on wrapped positions (plain differences)."""
    p = wrap_positions(pos, cell, pbc)
    delta = p[:, None, :, :] - p[:, :, None, :]
    return torch.linalg.vector_norm(delta, dim=-1)


def pair_set(pos: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, rc: float):
    """Sorted (frame, i, j) triples with 0 < d_ij < rc  (symm_funcs.py:18-31)."""
    d = distance_matrix(pos, cell, pbc)
    return torch.nonzero((d < rc) & (d > 0))


def cutoff_fn(d: torch.Tensor, rc: float) -> torch.Tensor:
    """symm_funcs.py:107-120."""
    return torch.where(d < rc, 0.5 * (torch.cos(math.pi * d / rc) + 1.0), torch.zeros_like(d))


def _radial(d, kind, prm):
    rc = prm[0]
    inside = (d < rc) & (d > 0)
    val = cutoff_fn(d, rc)
    if kind == G2:
        # positional binding (rc, eta, rs) exactly as the reference splats it (Q10)
        eta, rs = prm[1], prm[2]
        val = torch.exp(-eta * (d - rs) ** 2) * val
    return torch.where(inside, val, torch.zeros_like(val)).sum(-1)


def _angular(d, kind, prm, pow_mode):
    rc, eta, zeta, lam = prm
    inside = (d < rc) & (d > 0)
    fc = cutoff_fn(d, rc)
    dij = d[:, :, :, None]
    dik = d[:, :, None, :]
    djk = d[:, None, :, :]
    tri_ok = inside[:, :, :, None] & inside[:, :, None, :] & inside[:, None, :, :]
    denom = 2.0 * dij * dik
    cosv = torch.where(denom.abs() >= 1e-3,
                       (dij ** 2 + dik ** 2 - djk ** 2) / denom.clamp(min=1e-8),
                       torch.zeros_like(denom * djk))
    radial2 = dij ** 2 + dik ** 2
    if kind == G4:
        radial2 = radial2 + djk ** 2
    damp = 2.0 ** (1 - zeta) * torch.exp(-eta * radial2)
    base = 1 + lam * cosv.abs()
    if pow_mode == "complex64":
        powered = torch.pow(base.to(torch.complex64), zeta).real
    elif pow_mode == "exact":
        # guard the 0**0 / negative-rounding corner exactly like a real pow would
        powered = torch.pow(base.clamp(min=0.0), zeta)
    else:
        raise ValueError(pow_mode)
    ang = torch.where(cosv.abs() > 1e-3, powered.to(d.dtype), torch.zeros_like(cosv))
    term = ang * damp * (fc[:, :, :, None] * fc[:, :, None, :] * fc[:, None, :, :])
    return torch.where(tri_ok, term, torch.zeros_like(term)).sum(dim=(-2, -1))


def raw_descriptors(pos, cell, pbc, features, params, pow_mode="exact"):
    """Un-normalised G (B,N,F) and the reference's summed gradient dg (B,N,F,3)."""
    pos = pos.detach().clone().requires_grad_(True)
    d = distance_matrix(pos, cell, pbc)
    cols, grads = [], []
    for kind, prm in zip(features, params):
        kind = int(kind)
        if kind in (G1, G2):
            col = _radial(d, kind, prm)
        elif kind in (G4, G5):
            col = _angular(d, kind, prm, pow_mode)
        else:
            raise ValueError(f"Unknown symmetry function number: {kind}")
        (gr,) = torch.autograd.grad(col.sum(), pos, retain_graph=True)
        cols.append(col.detach())
        grads.append(gr.detach())
    G = torch.stack(cols, dim=-1)
    dG = torch.stack(grads, dim=-2)  # (B,N,F,3)
    return G, dG


def descriptors(pos, cell, pbc, sfd: dict, pow_mode="exact", chunk=2):
    """calculate_sf semantics (calculate_sf.py:23-136): (g_hat, dg)."""
    gs, dgs = [], []
    for s in range(0, pos.shape[0], chunk):
        G, dG = raw_descriptors(pos[s:s + chunk], cell, pbc, sfd["features"], sfd["params"], pow_mode)
        gs.append(G)
        dgs.append(dG)
    G = torch.cat(gs)
    dG = torch.cat(dgs)
    ghat = G / torch.linalg.vector_norm(G, dim=-1, keepdim=True)
    if torch.isnan(ghat).any():
        raise ValueError("NaN values in the symmetry functions")
    if torch.isnan(dG).any():
        raise ValueError("NaN values in the gradients of the symmetry functions")
    return ghat, dG


# ----------------------------------------------------------------------------
# atomic MLP, energy / force contraction, loss  (atomic_nn.py:27-48, bpnn.py:196-215,342-364)
# ----------------------------------------------------------------------------

def mlp_forward(x: torch.Tensor, weights, biases) -> torch.Tensor:
    """Linear -> sigmoid -> ... -> Linear (no activation on the last layer)."""
    h = x
    last = len(weights) - 1
    for li, (W, b) in enumerate(zip(weights, biases)):
        h = h @ W.T + b
        if li != last:
            h = torch.sigmoid(h)
    return h


def energy_force(ghat, dG, weights, biases, create_graph=False):
    """e (…,N,1), dE/dg_hat, f = -einsum(dE, dG)  (bpnn.py:342-354)."""
    x = ghat if ghat.requires_grad else ghat.detach().clone().requires_grad_(True)
    e = mlp_forward(x, weights, biases)
    (dE,) = torch.autograd.grad(e.sum(), x, create_graph=create_graph)
    f = -(dE.unsqueeze(-1) * dG).sum(dim=-2)
    return e, dE, f


def bp_loss(e_atoms, f, e_ref, f_ref, e_coeff, f_coeff):
    """bpnn.py:196-215 with criterion = MSELoss; e_atoms (B,N,1) summed over atoms."""
    e_tot = e_atoms.sum(dim=1)
    le = e_coeff * torch.mean((e_tot - e_ref) ** 2)
    lf = f_coeff / 3.0 * torch.mean((f - f_ref) ** 2)
    return le + lf, le, lf
