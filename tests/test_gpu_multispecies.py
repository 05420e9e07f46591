"""Multi-species descriptors with cross-species neighbours -- an EXTENSION with no reference oracle: the
reference isolates species (nnmd/nn/dataset.py:59-67).  The checker is a dense torch restatement of the
species-weighted functions written here (same masks and thresholds as symm_funcs.py:18-50, 193-211) whose summed gradient
comes from torch.autograd, plus central finite differences of sum_i G_i,f."""
import math

import numpy as np
import pytest
import torch

from helpers import col_relerr, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dense(pos, tags, feats, params, W):
    """G (N, F) of ONE frame, differentiable w.r.t. pos.  W: (F, S, S) weights as nnmd_b200 builds them."""
    N = pos.shape[0]
    diff = pos[None, :, :] - pos[:, None, :]
    d = torch.sqrt((diff ** 2).sum(-1) + torch.eye(N, dtype=pos.dtype))  # eye keeps sqrt differentiable on the diagonal
    off = ~torch.eye(N, dtype=torch.bool)
    cols = []
    for f, (kind, prm) in enumerate(zip(feats, params)):
        rc = prm[0]
        ok = off & (d < rc)
        fc = torch.where(ok, 0.5 * (torch.cos(math.pi * d / rc) + 1.0), torch.zeros_like(d))
        if kind in (1, 2):
            eta, rs = (prm[1], prm[2]) if kind == 2 else (0.0, 0.0)
            w = W[f, 0][tags]                                     # weight of neighbour j
            cols.append((torch.exp(-eta * (d - rs) ** 2) * fc * w[None, :]).sum(dim=1))
            continue
        eta, zeta, lam = prm[1], prm[2], prm[3]
        dij, dik, djk = d[:, :, None], d[:, None, :], d[None, :, :]
        closed = ok[:, :, None] & ok[:, None, :] & ok[None, :, :]
        two = 2.0 * dij * dik
        cos = torch.where(two >= 1e-3, (dij ** 2 + dik ** 2 - djk ** 2) / two.clamp(min=1e-8), torch.zeros_like(two * djk))
        ac = cos.abs()
        P = torch.where(ac > 1e-3, (1.0 + lam * ac).clamp(min=0.0) ** zeta, torch.zeros_like(ac))
        rad = dij ** 2 + dik ** 2 + (djk ** 2 if kind == 4 else 0.0)
        pw = W[f][tags][:, tags]                                   # (N, N): pair weight of neighbours (j, k)
        term = P * torch.exp(-eta * rad) * fc[:, :, None] * fc[:, None, :] * fc[None, :, :] * pw[None, :, :]
        term = torch.where(closed, term, torch.zeros_like(term))
        cols.append(2.0 ** (1.0 - zeta) * term.sum(dim=(1, 2)))
    return torch.stack(cols, dim=1)


def _system(seed=0, n_a=9, n_b=6, box=5.5):
    rng = np.random.default_rng(seed)
    pa = rng.uniform(0.3, box - 0.3, (2, n_a, 3))
    pb = rng.uniform(0.3, box - 0.3, (2, n_b, 3))
    cell = np.eye(3) * box
    sfd = {"features": [2, 2, 2, 1, 4, 4, 5, 4, 4],
           "params": [[3.5, 0.8, 1.0], [3.5, 0.8, 1.0], [3.0, 0.3, 0.5], [3.2], [3.5, 0.1, 2.0, 1.0], [3.5, 0.1, 2.0, 1.0],
                      [3.5, 0.05, 1.5, -1.0], [3.0, 0.2, 3.0, 0.5], [3.5, 0.1, 2.0, 1.0]],
           "neighbors": [None, "B", ("A", "B"), {"A": 2.0, "B": -0.5}, None, ("A", "B"), "A", {"A": 1.5, "B": 0.7}, ("B", "B")]}
    return pa, pb, cell, sfd


def test_species_weight_table_selectors():
    from nnmd_b200.features.multispecies import species_weight_table
    _, _, _, sfd = _system()
    W = species_weight_table(sfd, ["A", "B"])
    assert W.shape == (9, 2, 2)
    assert W[0, 0].tolist() == [1, 1] and W[1, 0].tolist() == [0, 1] and W[2, 0].tolist() == [1, 1] and W[3, 0].tolist() == [2.0, -0.5]
    assert W[4].tolist() == [[1, 1], [1, 1]] and W[5].tolist() == [[0, 1], [1, 0]] and W[6].tolist() == [[1, 0], [0, 0]]
    assert np.allclose(W[7], [[2.25, 1.05], [1.05, 0.49]]) and W[8].tolist() == [[0, 0], [0, 1]]
    with pytest.raises(ValueError):
        species_weight_table({"features": [2], "params": [[3.0, 1, 0]], "neighbors": ["C"]}, ["A", "B"])


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 2e-5)])
def test_cross_species_descriptors_against_dense_autograd(dtype, tol):
    from nnmd_b200.features import calculate_sf_species
    from nnmd_b200.features.multispecies import species_weight_table
    pa, pb, cell, sfd = _system()
    W = torch.tensor(species_weight_table(sfd, ["A", "B"]))
    cart = {"A": torch.tensor(pa, dtype=dtype, device=DEV), "B": torch.tensor(pb, dtype=dtype, device=DEV)}
    g, dg = calculate_sf_species(cart, torch.tensor(cell, dtype=dtype), torch.ones(3, dtype=dtype), sfd)
    assert g["A"].shape == (2, 9, 9) and dg["B"].shape == (2, 6, 9, 3)
    tags = torch.tensor([0] * 9 + [1] * 6)
    for b in range(2):
        pos = torch.tensor(np.concatenate([pa[b], pb[b]]), dtype=torch.float64, requires_grad=True)
        G = _dense(pos, tags, sfd["features"], sfd["params"], W)
        dG = torch.stack([torch.autograd.grad(G[:, f].sum(), pos, retain_graph=True)[0] for f in range(G.shape[1])], dim=1)
        ghat = (G / G.norm(dim=1, keepdim=True)).detach()
        got_g = torch.cat([g["A"][b], g["B"][b]]).cpu()
        got_dg = torch.cat([dg["A"][b], dg["B"][b]]).cpu()
        assert relerr(got_g, ghat) < tol, relerr(got_g, ghat)
        assert col_relerr(got_dg, dG, 1) < tol, col_relerr(got_dg, dG, 1)
        # the element-resolved columns really differ from the all-neighbour ones
        assert float((G[:, 0] - G[:, 1]).abs().max()) > 1e-3 and float((G[:, 4] - G[:, 5]).abs().max()) > 1e-6


def test_summed_gradient_is_the_finite_difference_of_the_summed_functions():
    """dg[a, f, :] = d(sum_i G_i,f) / d r_a for the species-weighted functions, by central differences in fp64."""
    from nnmd_b200.features.multispecies import raw_descriptors_species
    pa, pb, cell, sfd = _system(seed=3, n_a=6, n_b=5, box=4.5)
    cellt, pbc = torch.tensor(cell), torch.ones(3, dtype=torch.float64)

    def total(pa_, pb_):
        out = raw_descriptors_species({"A": torch.tensor(pa_[:1], device=DEV), "B": torch.tensor(pb_[:1], device=DEV)}, cellt, pbc, sfd)
        return out[4].sum(dim=0).cpu(), out[5].cpu()

    _, dG = total(pa, pb)
    h = 1e-5
    for (arr, a, off) in ((pa, 2, 0), (pb, 4, 6)):
        for c in range(3):
            plus, minus = arr.copy(), arr.copy()
            plus[0, a, c] += h
            minus[0, a, c] -= h
            sp = total(plus, pb)[0] if arr is pa else total(pa, plus)[0]
            sm = total(minus, pb)[0] if arr is pa else total(pa, minus)[0]
            fd = (sp - sm) / (2 * h)
            assert float((fd - dG[off + a, :, c]).abs().max()) < 1e-6 * max(1.0, float(dG.abs().max())), (a, c)


def test_one_species_with_unit_weights_reproduces_the_single_species_path():
    from nnmd_b200.features import calculate_sf, calculate_sf_species
    pa, _, cell, sfd = _system()
    sfd1 = {"features": sfd["features"], "params": sfd["params"]}
    cart = torch.tensor(pa, dtype=torch.float64, device=DEV)
    cellt, pbc = torch.tensor(cell), torch.ones(3, dtype=torch.float64)
    g0, dg0 = calculate_sf(cart, cellt, pbc, sfd1)
    g1, dg1 = calculate_sf_species({"X": cart}, cellt, pbc, sfd1)
    assert relerr(g1["X"].cpu(), g0.cpu()) < 1e-13 and relerr(dg1["X"].cpu(), dg0.cpu()) < 1e-12


def test_cross_species_energy_and_forces_use_each_species_network():
    """Type-sorted rows, one network per species, forces back in the caller's per-species order."""
    from nnmd_b200.engine import energy_forces_species
    from nnmd_b200.features import calculate_sf_species
    from nnmd_b200.nn import AtomicNN
    pa, pb, cell, sfd = _system(seed=5)
    dt = torch.float64
    cart = {"B": torch.tensor(pb, dtype=dt, device=DEV), "A": torch.tensor(pa, dtype=dt, device=DEV)}   # caller's order: B first
    cellt, pbc = torch.tensor(cell, dtype=dt), torch.ones(3, dtype=dt)
    torch.manual_seed(1)
    nets = {"A": AtomicNN(9, [8, 4]).to(DEV, dt), "B": AtomicNN(9, [6]).to(DEV, dt)}
    E, e_at, F = energy_forces_species(cart, cellt, pbc, sfd, nets)
    g, dg = calculate_sf_species(cart, cellt, pbc, sfd)
    tot = torch.zeros(2, dtype=dt, device=DEV)
    for s in ("A", "B"):
        x = g[s].clone().requires_grad_(True)
        e = nets[s](x)
        dE = torch.autograd.grad(e.sum(), x)[0]
        f = -torch.einsum("ijk,ijkl->ijl", dE, dg[s])
        assert relerr(e_at[s].cpu(), e.squeeze(-1).detach().cpu()) < 1e-10 and relerr(F[s].cpu(), f.cpu()) < 1e-10
        tot += e.sum(dim=(1, 2)).detach()
    assert relerr(E.cpu(), tot.cpu()) < 1e-12 and F["B"].shape == (2, 6, 3) and F["A"].shape == (2, 9, 3)
