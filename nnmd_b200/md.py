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
    """x += v dt + a dt^2/2 ; v += (a_old + a_new) dt/2, single species, everything on the GPU, NO host round trip per step.

    The neighbour list lives in a `NeighborWorkspace`: capacity-bound, built without a host synchronisation, and with
    skin > 0 reused until some atom has moved more than skin / 2 (decided on the device).  After two eager steps the whole
    step -- integrator update, K0 wrap, skin check, (conditional) K1, K2/K3/K4, K5, K6, K4c -- is captured in ONE CUDA graph
    and replayed; `step(n)` synchronises once, at its end, to read the sticky status word (NaN descriptors raise the
    reference's ValueError; a neighbour list that outgrew its buffers makes the n steps run again from a snapshot with
    larger buffers).  Replaces the reference's hand-rolled loop (nnmd/md/verlet.py:97-218: O(N^2) Python, one `predict`
    call with a stale signature per step) and ASE's host integrator round trip (samples/cu111/md_simulation.py:46-52)."""

    def __init__(self, positions: torch.Tensor, velocities: torch.Tensor, mass: float, cell: torch.Tensor, pbc: torch.Tensor,
                 symm_funcs_data: dict, net, dt: float, mode: str = "materialize", skin: float = 0.0, graph: bool = True):
        if not positions.is_cuda:
            raise ValueError("VelocityVerlet keeps its state on the GPU: pass CUDA tensors")
        from .features.calculate_sf import get_plan
        from .neighbor import NeighborWorkspace
        self.x = positions.detach().clone()
        self.v = velocities.detach().to(self.x).clone()
        self.mass, self.dt = float(mass), float(dt)
        # cell and pbc are kernel ARGUMENTS of K0: host copies avoid a device read-back in every step
        self.cell, self.pbc = cell.detach().to("cpu", self.x.dtype), pbc.detach().to("cpu", self.x.dtype)
        self.sfd, self.net, self.mode = symm_funcs_data, net, mode
        plan = get_plan(symm_funcs_data, self.x.dtype, self.x.device)
        self.ws = NeighborWorkspace(1, self.x.shape[0], plan.rc_max, self.x.dtype, self.x.device, skin=skin)
        self.flag = self.ws.status                      # sticky OR of every step's flag word (device)
        self.use_graph = bool(graph)
        self._graph, self._eager_steps = None, 0
        e, f = self._forces()
        self.energy, self.f = e.clone(), f.clone()      # static tensors: the captured step writes into them
        self.steps = 0
        self.check()

    def _forces(self):
        res = energy_forces(self.x[None], self.cell, self.pbc, self.sfd, self.net, mode=self.mode, check=False, workspace=self.ws)
        return res.energy[0], res.forces[0]

    def _one_step(self):
        dt, m = self.dt, self.mass
        self.x += self.v * dt + self.f * (0.5 * dt * dt / m)
        e, f_new = self._forces()
        self.v += (self.f + f_new) * (0.5 * dt / m)
        self.f.copy_(f_new)
        self.energy.copy_(e)

    def check(self) -> bool:
        """One host sync.  Raises the reference's ValueError (calculate_sf.py:106-109) if any step so far produced NaN
        descriptors, e.g. an atom that lost all its neighbours; returns False when the neighbour list outgrew its buffers
        (they have been regrown: the caller repeats the steps)."""
        from . import _lib
        from .features.calculate_sf import raise_on_flag
        if not self.ws.validate():
            self._graph, self._eager_steps = None, 0    # buffers were reallocated: the captured addresses are stale
            return False
        bits = int(self.flag.item()) & ~_lib.FLAG_OVERFLOW
        if bits:
            raise_on_flag(torch.tensor([bits]))
        return True

    def _advance(self, n: int) -> None:
        for _ in range(n):
            if self.use_graph and self._graph is None and self._eager_steps >= 2:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._one_step()
                self._graph = g                          # a capture does not execute: replay below
            if self._graph is not None:
                self._graph.replay()
            else:
                self._one_step()
                self._eager_steps += 1
            self.steps += 1

    def step(self, n: int = 1, check: bool = True):
        snap = (self.x.clone(), self.v.clone(), self.f.clone(), self.energy.clone(), self.steps) if check else None
        self._advance(n)
        if check and not self.check():                   # list overflow somewhere in these n steps: redo them
            self.x.copy_(snap[0]); self.v.copy_(snap[1]); self.f.copy_(snap[2]); self.energy.copy_(snap[3])
            self.steps = snap[4]
            self._advance(n)
            if not self.check():
                raise RuntimeError("VelocityVerlet: the neighbour list overflowed twice in a row")
        return self.energy

    def list_builds(self) -> int:
        """Number of neighbour-list builds so far (one host sync): steps - builds were served by the skin."""
        return int(self.ws.n_builds.item())
