"""Auto-parameter scheme of the reference (nnmd/features/auto_params.py:4-59), host-side numpy.

Output rows: G1 [rc]; G2 [rc, r_s, eta_rad]  (note the order: `calculate_sf` binds rows POSITIONALLY to
g2_function(r, cutoff, eta, rs), so with these rows eta := r_s and rs := eta_rad -- SURVEY Q10);
G3 [rc, kappa] (emitted but not implemented downstream); G4/G5 [rc, eta_ang, zeta, lambda].
All values rounded to 6 decimals like the reference.
"""
from __future__ import annotations

import numpy as np


def _zeta_ladder(n_func: int) -> np.ndarray:
    n_levels = int(np.ceil(n_func / 2 - 1)) + 1
    levels = [1 if k == 0 else 1 + k * 30 / (n_func - 2) for k in range(n_levels)]
    return np.round(levels, 6)


def calculate_params(r_cutoff, N_g1, N_g2, N_g3, N_g4, N_g5):
    """Parameter rows for N_g1 G1, N_g2 G2, N_g3 G3, N_g4 G4 and N_g5 G5 functions sharing one cutoff."""
    shifts = np.round(np.linspace(0, r_cutoff, N_g2), 6)
    spacing = r_cutoff / (N_g2 - 1)  # N_g2 == 1 -> ZeroDivisionError, as in the reference
    eta_radial = np.round(5 * np.log(10) / (4 * spacing**2), 6)
    eta_angular = np.round(2 * np.log(10) / (r_cutoff**2), 6)
    zeta4, zeta5 = _zeta_ladder(N_g4), _zeta_ladder(N_g5)

    rows = [[r_cutoff] for _ in range(N_g1)]
    rows += [[r_cutoff, shifts[k], eta_radial] for k in range(N_g2)]
    rows += [[r_cutoff, 1] for _ in range(N_g3)]
    rows += [[r_cutoff, eta_angular, zeta4[k // 2], (-1) ** k * 1] for k in range(N_g4)]
    rows += [[r_cutoff, eta_angular, zeta5[k // 2], (-1) ** k * 1] for k in range(N_g5)]
    return rows
