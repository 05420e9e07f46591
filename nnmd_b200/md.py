"""Callers of the hot path for molecular dynamics ("next" rows 1 and 4).

`NNMDCalc`       -- the reference's ASE calculator (nnmd/md/calculator.py:10-77) as a thin adaptor over `BPNN.predict`.
                    Species are matched by CHEMICAL SYMBOL (the reference compares symbols with atomic numbers,
                    calculator.py:34 vs :42, and selects nothing).  Works with any object that has `.positions`,
                    `.cell`, `.pbc` and `get_chemical_symbols()`; when `ase` is installed `as_ase_calculator()` wraps it
                    in a real `ase.calculators.calculator.Calculator`.
`VelocityVerlet` -- device-resident integrator: positions, velocities and forces stay on the GPU; one call of
                    `engine.energy_forces` per step, no host round trip (the reference's hand-rolled loop,
                    nnmd/md/verlet.py:97-218, is O(N^2) Python and calls `predict` with a stale signature).
The forces are the reference-semantic ones.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import energy_forces


class NNMDCalc:
    implemented_properties = ["energy", "forces"]

    def __init__(self, model, atoms=None, properties=("energy", "forces"), symm_funcs_data: dict | None = None):
        self.model = model
        self.atoms = atoms
        self.implemented_properties = list(properties)
        self.symm_funcs_data = symm_funcs_data
        self.results = {}

    def _prediction(self, atoms):
        symbols = np.asarray(atoms.get_chemical_symbols())
        positions = {}
        for spec in self.model.species:
            mask = symbols == spec
            positions[spec] = torch.tensor(np.asarray(atoms.positions)[mask], dtype=self.model.dtype, device=self.model.device)
        cell = torch.tensor(np.asarray(getattr(atoms.cell, "array", atoms.cell)), dtype=self.model.dtype, device=self.model.device)
        pbc = torch.tensor(np.asarray(atoms.pbc, dtype=float), dtype=self.model.dtype)
        e, f = self.model.predict(positions, cell, self.symm_funcs_data, pbc=pbc)
        # scatter per-species forces back to the caller's atom order
        forces = np.zeros((len(symbols), 3))
        off = 0
        f = f.reshape(-1, 3).detach().cpu().numpy()
        for spec in self.model.species:
            mask = symbols == spec
            n = int(mask.sum())
            forces[mask] = f[off:off + n]
            off += n
        return float(e.sum().item()), forces

    def calculate(self, atoms=None, properties=None, system_changes=None):
        if atoms is not None:
            self.atoms = atoms
        energy, forces = self._prediction(self.atoms)
        self.results = {"energy": energy, "forces": forces}
        return self.results

    def as_ase_calculator(self):
        """A real ASE Calculator delegating to this object (needs `ase`)."""
        from ase.calculators.calculator import Calculator, all_changes

        outer = self

        class _Calc(Calculator):
            implemented_properties = list(outer.implemented_properties)

            def calculate(self, atoms=None, properties=None, system_changes=all_changes):
                Calculator.calculate(self, atoms, properties, system_changes)
                self.results = dict(outer.calculate(self.atoms))

        return _Calc()


class VelocityVerlet:
    """x += v dt + a dt^2/2 ; v += (a_old + a_new) dt/2, single species, everything on the GPU."""

    def __init__(self, positions: torch.Tensor, velocities: torch.Tensor, mass: float, cell: torch.Tensor, pbc: torch.Tensor,
                 symm_funcs_data: dict, net, dt: float, mode: str = "materialize"):
        if not positions.is_cuda:
            raise ValueError("VelocityVerlet keeps its state on the GPU: pass CUDA tensors")
        self.x = positions.detach().clone()
        self.v = velocities.detach().to(self.x).clone()
        self.mass, self.dt = float(mass), float(dt)
        # cell and pbc are kernel ARGUMENTS of K0: host copies avoid a device read-back in every step
        self.cell, self.pbc = cell.detach().to("cpu", self.x.dtype), pbc.detach().to("cpu", self.x.dtype)
        self.sfd, self.net, self.mode = symm_funcs_data, net, mode
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.x.device)  # sticky OR of every step's NaN flag word
        self.energy, self.f = self._forces()
        self.steps = 0
        self.check()

    def _forces(self):
        res = energy_forces(self.x[None], self.cell, self.pbc, self.sfd, self.net, mode=self.mode, check=False)
        self.flag |= res.flag  # device-side: no host sync inside the loop
        return res.energy[0], res.forces[0]

    def check(self) -> None:
        """Raise the reference's ValueError (calculate_sf.py:106-109) if any step so far produced NaN descriptors, e.g. an
        atom that lost all its neighbours.  One host sync; `step` calls it once at the end of its n steps."""
        from .features.calculate_sf import raise_on_flag
        raise_on_flag(self.flag)

    def step(self, n: int = 1, check: bool = True):
        dt, m = self.dt, self.mass
        for _ in range(n):
            self.x += self.v * dt + self.f * (0.5 * dt * dt / m)
            f_old = self.f
            self.energy, self.f = self._forces()
            self.v += (f_old + self.f) * (0.5 * dt / m)
            self.steps += 1
        if check:
            self.check()
        return self.energy
