"""Positions -> (descriptors) -> per-atom energy -> force, entirely on the GPU kernels (K0..K6).

This is the evaluation pipeline behind `BPNN.predict` and `bench.py`.  Force semantics are the reference's
(nnmd/nn/bpnn.py:347-354):  F[a] = - sum_f (dE/dg_hat[a,f]) * dG[a,f,:]  with dG the un-normalised
summed gradient of calculate_sf.

Two data flows over the same kernels:
  materialize : K1 -> K2/K3 (+K4a: dG written once to HBM) -> K5 -> K6 -> K4c (contract against dG)
  fused       : K1 -> K2/K3 (values only) -> K5 -> K6 -> K4b (re-walks the triangles, contracts on the fly;
                no (N,F,3) tensor ever exists -- for structures whose dG would not fit)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib
from .features.calculate_sf import SfPlan, edge_scratch, get_plan, normalize, raise_on_flag, raw_descriptors
from .neighbor import NeighborList, NeighborWorkspace, build_from_wrapped, wrap_positions


@dataclass
class EvalResult:
    energy: torch.Tensor          # (n_frames,) total energy per frame
    e_atoms: torch.Tensor         # (n_frames, n_centres)
    forces: torch.Tensor          # (n_frames, n_centres, 3)
    g: torch.Tensor               # (n_frames, n_centres, F) normalised descriptors
    nlist: NeighborList
    flag: torch.Tensor


def force_contract(dE: torch.Tensor, dG: torch.Tensor) -> torch.Tensor:
    """K4c: F[a,:] = - sum_f dE[a,f] dG[a,f,:]   (einsum "ijk,ijkl->ijl" of bpnn.py:353)."""
    lib = _lib.load()
    nf = dE.shape[-1]
    dE2 = dE.detach().reshape(-1, nf).contiguous()
    dG2 = dG.detach().reshape(-1, nf, 3).contiguous()
    out = torch.empty((dE2.shape[0], 3), dtype=dE.dtype, device=dE.device)
    with torch.cuda.device(dE.device):
        _lib.check(lib.nnmd_force_contract(_lib.dtype_code(dE.dtype), _lib.ptr(dE2), _lib.ptr(dG2), dE2.shape[0], nf,
                                           _lib.ptr(out), _lib.stream_ptr()))
    return out.view(*dE.shape[:-1], 3)


def virial(nl: NeighborList, forces: torch.Tensor, origin=None) -> torch.Tensor:
    """Per-frame virial W = - sum_a (p_a - o) (x) F_a, shape (n_frames, 3, 3), on the wrapped positions of `nl`.

    Extension (the reference has none).  Because the reference never adds periodic images every frame is a
    finite cluster, so this is the usual pair virial whenever the forces sum to zero; o defaults to the frame centroid."""
    import ctypes
    lib = _lib.load()
    f2 = forces.detach().reshape(-1, 3).contiguous()
    out = torch.empty((nl.n_frames, 3, 3), dtype=f2.dtype, device=f2.device)
    ws = torch.empty(max(int(lib.nnmd_virial_workspace_bytes(nl.n_frames)), 1), dtype=torch.uint8, device=f2.device)
    org = None if origin is None else (ctypes.c_double * 3)(*[float(v) for v in origin])
    with torch.cuda.device(f2.device):
        _lib.check(lib.nnmd_virial(_lib.dtype_code(f2.dtype), _lib.ptr(nl.posw), _lib.ptr(f2), nl.n_frames, nl.n_atoms, nl.n_centres,
                                   org, _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr()))
    return out


def fused_force(plan: SfPlan, nl: NeighborList, w: torch.Tensor, edge: bool = True, with_virial: bool = False, workspace=None):
    """K4b: forces from dE/dg_hat without a materialised dG (edge=False: vertex-centric angular kernel).
    with_virial=True also returns the per-frame virial (n_frames, 3, 3) of those forces from the same call."""
    lib = _lib.load()
    w2 = w.detach().reshape(-1, plan.nf).contiguous()
    out = torch.empty((nl.n_rows, 3), dtype=w.dtype, device=w.device)
    vir = torch.empty((nl.n_frames, 3, 3), dtype=w.dtype, device=w.device) if with_virial else None
    ews = edge_scratch(plan, nl, workspace) if edge else None
    with torch.cuda.device(w.device):
        _lib.check(lib.nnmd_sf_force_flag(plan.handle, _lib.ptr(nl.posw), nl.n_frames, nl.n_atoms, nl.n_centres,
                                          _lib.ptr(nl.offsets), _lib.ptr(nl.idx), nl.max_row, _lib.ptr(nl.edge_offsets),
                                          nl.n_edges, _lib.ptr(ews), 0 if ews is None else ews.numel(), _lib.ptr(w2),
                                          _lib.ptr(out), _lib.ptr(vir), _lib.ptr(nl.status), _lib.stream_ptr()))
    return (out, vir) if with_virial else out


def energy_forces(cart: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, net,
                  mode: str = "materialize", mlp_precision: str = "auto", n_centres: int | None = None,
                  check: bool = True, workspace: NeighborWorkspace | None = None) -> EvalResult:
    """Evaluate one species: cart (B,N,3) on the GPU -> energies and reference-semantic forces.

    workspace: a NeighborWorkspace for repeated evaluations of the same system (MD, the sharded bench step): the neighbour
    build then runs without a host synchronisation into static buffers (and reuses the list inside its skin); with
    check=False the whole call only enqueues work -- poll `workspace.validate()` / the returned flag when convenient."""
    _lib.require_cuda()
    if cart.dim() == 2:
        cart = cart.unsqueeze(0)
    B, N, _ = cart.shape
    plan = get_plan(symm_funcs_data, cart.dtype, cart.device)
    posw = wrap_positions(cart, cell, pbc)
    if workspace is not None:
        if (workspace.n_frames, workspace.n_atoms, workspace.n_centres) != (B, N, N if n_centres is None else int(n_centres)):
            raise ValueError("energy_forces: the NeighborWorkspace was made for another system size")
        nl = workspace.build(posw)
    else:
        nl = build_from_wrapped(posw, B, N, plan.rc_max, n_centres)
    nc = nl.n_centres
    if mode == "materialize":
        G, dG, flag = raw_descriptors(plan, nl, with_grad=True, workspace=workspace)
    elif mode == "fused":
        G, dG, flag = raw_descriptors(plan, nl, with_grad=False, workspace=workspace)
    else:
        raise ValueError(f"unknown mode {mode!r}")
    if mode == "materialize" and mlp_precision in ("auto", "tf32") and net.fused_shape(G.dtype):
        # K5 + K6 + K4c in ONE pass over G / dG: g_hat and dE/dg_hat never travel through HBM
        e, f = net.energy_force_fused(G, dG, flag, out_g=G)
        g = G
    else:
        g = normalize(G, flag, out=G)
        e, dE = net.energy_and_grad(g, need_grad=True, precision=mlp_precision)
        if mode == "materialize":
            f = force_contract(dE, dG)
        else:
            f = fused_force(plan, nl, dE, workspace=workspace)
    if check:
        raise_on_flag(flag)
    e_atoms = e.view(B, nc)
    return EvalResult(e_atoms.sum(dim=1), e_atoms, f.view(B, nc, 3), g.view(B, nc, plan.nf), nl, flag)


class GraphedEnergyForces:
    """`energy_forces` on a FIXED position buffer, captured once in a CUDA graph and replayed: for callers that evaluate the
    same system over and over with new coordinates written into the same tensor (the sharded bench step, MD drivers).
    The neighbour list lives in a NeighborWorkspace (capacity-bound, no host sync), every intermediate is static, and a
    replay is one graph launch instead of ~45 kernel launches -- what matters once a step is down to a few milliseconds.
    `__call__` returns the EvalResult whose tensors the replay overwrites; poll `ok()` when convenient (one host sync): it
    raises the reference's ValueError on NaN descriptors and returns False after a buffer overflow (re-captured: call again)."""

    def __init__(self, pos_buffer: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, net,
                 n_centres: int | None = None, mode: str = "materialize", mlp_precision: str = "auto", skin: float = 0.0):
        if pos_buffer.dim() != 2 or not pos_buffer.is_cuda:
            raise ValueError("GraphedEnergyForces needs a CUDA position buffer of shape (n_atoms, 3)")
        self.pos, self.args = pos_buffer, (cell.detach().to("cpu", pos_buffer.dtype), pbc.detach().to("cpu", pos_buffer.dtype), symm_funcs_data, net)
        self.kw = dict(mode=mode, mlp_precision=mlp_precision, n_centres=n_centres, check=False)
        plan = get_plan(symm_funcs_data, pos_buffer.dtype, pos_buffer.device)
        self.ws = NeighborWorkspace(1, pos_buffer.shape[0], plan.rc_max, pos_buffer.dtype, pos_buffer.device, n_centres=n_centres, skin=skin)
        self.graph = None
        self._capture()

    def _eval(self):
        return energy_forces(self.pos[None], *self.args, workspace=self.ws, **self.kw)

    def _capture(self):
        for _ in range(2):  # the first call sizes the capacities (one sync), the second runs the steady-state sequence
            self.result = self._eval()
        torch.cuda.synchronize()
        if not self.ws.validate():
            self.result = self._eval()
            torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.result = self._eval()
        self.graph = g

    def __call__(self) -> EvalResult:
        self.graph.replay()
        return self.result

    def ok(self) -> bool:
        if not self.ws.validate():      # buffers regrown: the captured addresses are stale
            self.graph = None
            self._capture()
            return False
        raise_on_flag(self.result.flag)
        return True


def energy_forces_species(cartesians: dict, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, nets: dict,
                          mlp_precision: str = "auto", check: bool = True):
    """Multi-species evaluation with CROSS-species neighbours (extension; see features/multispecies.py):
    {species: (B, N_s, 3)} and one atomic network per species -> (energy (B,), {species: per-atom energies (B, N_s)},
    {species: forces (B, N_s, 3)}), atoms of each species in the caller's order.  The force is the reference's local
    contraction with the network of the atom's own species."""
    from .features.multispecies import raw_descriptors_species
    species, counts, plan, nl, G, dG, flag = raw_descriptors_species(cartesians, cell, pbc, symm_funcs_data)
    g = normalize(G, flag, out=G)
    B, N, F = nl.n_frames, nl.n_atoms, plan.nf
    g3, dG4 = g.view(B, N, F), dG.view(B, N, F, 3)
    e_at, forces, off = {}, {}, 0
    energy = torch.zeros(B, dtype=g.dtype, device=g.device)
    for s, n in zip(species, counts):       # rows are type-sorted: one contiguous block of atoms per species
        gs = g3[:, off:off + n].contiguous()
        e, dE = nets[s].energy_and_grad(gs, need_grad=True, precision=mlp_precision)
        forces[s] = force_contract(dE, dG4[:, off:off + n].contiguous()).view(B, n, 3)
        e_at[s] = e.view(B, n)
        energy = energy + e_at[s].sum(dim=1)
        off += n
    if check:
        raise_on_flag(flag)
    return energy, e_at, forces
