"""The algebra behind the moment form of K3e (nnmd_b200/csrc/sf_edge_m.cuh), checked on the CPU with torch.autograd in float64.

For the edge (a, j) of length x and a closing third atom k (y = d_ak, z = d_jk) the kernel never differentiates the radial
product: every centre term factorises as E_X = e_X(x) W_X, so that with b = 1 + lambda |c_X| and q = b^(zeta - 1)
    d/dx sum_X b_X^zeta E_X  =  zeta * sum_X q_X A_X  +  e_a'(x) (q_a b_a W_a + q_j b_j W_j)  +  e_k'(x) q_k b_k W_k ,
    A_X = lambda sgn(c_X) E_X dc_X/dx
(G4: e_X = psi for all centres; G5 as the reference implements it, SURVEY Q3: e_a = e_j = psi, e_k = fc).  The same W
moments give the values: b_a^zeta E_a = e_a(x) q_a b_a W_a.  This is test infrastructure: it restates formulas, not kernels.
"""
import math

import pytest
import torch

RC, ETA = 6.0, 0.127921


def fc(r):
    return 0.5 * (torch.cos(math.pi * r / RC) + 1.0)


def psi(r):
    return torch.exp(-ETA * r * r) * fc(r)


def cosines(x, y, z):
    return ((x * x + y * y - z * z) / (2 * x * y),   # at a: sides x, y
            (x * x + z * z - y * y) / (2 * x * z),   # at j: sides x, z
            (y * y + z * z - x * x) / (2 * y * z))   # at k: sides y, z


def centre_terms(kind, x, y, z):
    """E_X of the three centres and their factorisation (e_X(x), W_X)."""
    if kind == 4:
        e = [psi(x)] * 3
        w = [psi(y) * psi(z)] * 3
    else:
        e = [psi(x), psi(x), fc(x)]
        w = [psi(y) * fc(z), psi(z) * fc(y), psi(y) * psi(z)]
    return e, w


@pytest.mark.parametrize("kind", [4, 5])
@pytest.mark.parametrize("lam", [1.0, -1.0])
@pytest.mark.parametrize("zeta", [1.0, 1.483871, 7.290323, 16.0])
def test_moment_form_equals_the_derivative_of_the_triangle_sum(kind, lam, zeta):
    gen = torch.Generator().manual_seed(int(zeta * 1000) + kind + (0 if lam > 0 else 7))
    checked = 0
    while checked < 50:
        p = torch.rand(3, 3, generator=gen, dtype=torch.float64) * 5.0
        x0, y0, z0 = (p[1] - p[0]).norm(), (p[2] - p[0]).norm(), (p[2] - p[1]).norm()
        if not all(0.5 < float(d) < RC for d in (x0, y0, z0)):
            continue
        if min(abs(float(c)) for c in cosines(x0, y0, z0)) < 1e-2:
            continue  # stay off the |cos| = 1e-3 discontinuity of the reference
        x = x0.clone().requires_grad_(True)
        e, w = centre_terms(kind, x, y0, z0)
        c = cosines(x, y0, z0)
        total = sum((1.0 + lam * ci.abs()) ** zeta * ei * wi for ci, ei, wi in zip(c, e, w))
        (want,) = torch.autograd.grad(total, x)
        # moment form, everything a function of plain numbers
        with torch.no_grad():
            xv = x.detach()
            c = cosines(xv, y0, z0)
            dcx = (1.0 / y0 - c[0] / xv, 1.0 / z0 - c[1] / xv, -xv / (y0 * z0))
            e, w = centre_terms(kind, xv, y0, z0)
        de = []
        for X in range(3):  # e_X'(x): the edge's own radial factor, differentiated on its own
            xr = xv.clone().requires_grad_(True)
            de.append(torch.autograd.grad(centre_terms(kind, xr, y0, z0)[0][X], xr)[0])
        b = [1.0 + lam * ci.abs() for ci in c]
        q = [bi ** (zeta - 1.0) for bi in b]
        A = [lam * torch.sign(ci) * ei * wi * di for ci, ei, wi, di in zip(c, e, w, dcx)]
        SA = sum(qi * Ai for qi, Ai in zip(q, A))
        MW = [qi * bi * wi for qi, bi, wi in zip(q, b, w)]
        phi = zeta * SA + de[0] * (MW[0] + MW[1]) + de[2] * MW[2]
        assert de[0] == de[1]
        assert abs(float(phi - want)) <= 1e-11 * max(1.0, abs(float(want)))
        # the values come out of the same moments
        for X in range(2):
            val = b[X] ** zeta * e[X] * w[X]
            assert abs(float(e[X] * MW[X] - val)) <= 1e-13 * max(1.0, abs(float(val)))
        checked += 1
