"""Multi-species symmetry functions with CROSS-species neighbours -- an extension.

The reference evaluates every species in isolation (`calculate_sf` is called once per species with only that species'
coordinates, nnmd/nn/dataset.py:59-67; atoms of species A never see species B: SURVEY Q9), so there is no reference
oracle for this module: tests/test_gpu_multispecies.py checks it against a dense restatement and finite differences.

Here all atoms of a frame are neighbours of each other.  Every function of the table may carry a neighbour selector:

    symm_funcs_data = {"features": [2, 2, 4, ...], "params": [[rc, eta, rs], ...],
                       "neighbors": [None, "Li", ("Cu", "Li"), {"Cu": 29.0, "Li": 3.0}, ...]}

    None                 every neighbour counts with weight 1 (the plain function over all atoms)
    "Li"                 radial: neighbours of species Li only;  angular: both neighbours Li      (element-resolved ACSF)
    ("Cu", "Li")         angular: one neighbour Cu and one Li (unordered);  radial: neighbours Cu or Li
    {"Cu": 29, "Li": 3}  weighted functions: radial weight w[s_j], angular weight w[s_j] * w[s_k]

The SAME table serves every central species (each species has its own atomic network on top, as in the reference).
`dg` keeps the reference's meaning: dg[a, f, :] = d(sum over ALL atoms i of G_i,f) / d r_a, and the force is the
reference's local contraction  F_a = - sum_f (dE/dg_hat[a, f]) dg[a, f, :]  with the network of a's species.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch

from .. import _lib
from ..neighbor import build_from_wrapped, wrap_positions
from .calculate_sf import _N_PARAMS, normalize, raise_on_flag, raw_descriptors
from .symm_funcs import SymmetryFunction


def species_weight_table(symm_funcs_data: dict, species: list) -> np.ndarray:
    """(nf, S, S) float64: block f = symmetric pair weights of function f; radial functions use row 0 as wr[s_j]."""
    feats = [int(k) for k in symm_funcs_data["features"]]
    sel = symm_funcs_data.get("neighbors") or [None] * len(feats)
    if len(sel) != len(feats):
        raise ValueError("symm_funcs_data['neighbors'] must have one entry per function")
    S = len(species)
    pos = {s: i for i, s in enumerate(species)}
    out = np.zeros((len(feats), S, S))
    for f, (kind, s) in enumerate(zip(feats, sel)):
        radial = kind in (1, 2)
        if s is None:
            w = np.ones(S)
            pw = np.ones((S, S))
        elif isinstance(s, dict):
            w = np.array([float(s.get(name, 0.0)) for name in species])
            pw = np.outer(w, w)
        else:
            names = (s,) if isinstance(s, str) else tuple(s)
            if any(n not in pos for n in names) or not 1 <= len(names) <= 2:
                raise ValueError(f"function {f}: neighbour selector {s!r} does not name one or two of the species {species}")
            w = np.zeros(S)
            w[[pos[n] for n in names]] = 1.0
            a, b = pos[names[0]], pos[names[-1]]
            pw = np.zeros((S, S))
            pw[a, b] = pw[b, a] = 1.0
        if radial:
            out[f, 0, :] = w
        else:
            out[f] = pw
    return out


class SpeciesPlan:
    """Device-resident multi-species parameter table (wraps nnmd_sf_plan_create_species)."""

    def __init__(self, symm_funcs_data: dict, species: list, dtype: torch.dtype, device: torch.device):
        feats, params = symm_funcs_data["features"], symm_funcs_data["params"]
        kinds = []
        table = np.zeros((len(feats), 4), dtype=np.float64)
        for f, (kind, row) in enumerate(zip(feats, params)):
            kind = SymmetryFunction(int(kind)).value
            row = list(row)
            if len(row) != _N_PARAMS[kind]:
                raise TypeError(f"G{kind} takes {_N_PARAMS[kind]} parameters, got {len(row)}")
            table[f, :len(row)] = [float(v) for v in row]
            kinds.append(kind)
        self.kinds = np.asarray(kinds, dtype=np.int32)
        self.table = table
        self.weights = np.ascontiguousarray(species_weight_table(symm_funcs_data, species))
        self.species = list(species)
        self.nf, self.dtype, self.device = len(kinds), dtype, device
        lib = _lib.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.nnmd_sf_plan_create_species(_lib.dtype_code(dtype), self.nf, self.kinds.ctypes.data_as(ctypes.c_void_p),
                                                       self.table.ctypes.data_as(ctypes.c_void_p), len(species),
                                                       self.weights.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle)))
        self.handle = handle
        self.rc_max = float(lib.nnmd_sf_plan_rc_max(handle))
        self._finalizer = weakref.finalize(self, lib.nnmd_sf_plan_destroy, handle)


def _concat(cartesians: dict):
    species = list(cartesians.keys())
    carts = []
    for s in species:
        c = cartesians[s]
        c = c.unsqueeze(0) if c.dim() == 2 else c
        if c.dim() != 3 or c.shape[-1] != 3:
            raise ValueError(f"cartesians[{s!r}] must have shape (n_batch, n_atoms, 3), got {tuple(cartesians[s].shape)}")
        carts.append(c)
    if len({c.shape[0] for c in carts}) != 1 or len({c.dtype for c in carts}) != 1:
        raise ValueError("every species needs the same number of frames and the same dtype")
    counts = [c.shape[1] for c in carts]
    return species, counts, torch.cat([c.detach().cuda() for c in carts], dim=1)


def raw_descriptors_species(cartesians: dict, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, with_grad: bool = True):
    """Shared front end: atoms of all species concatenated in dict order (rows are type-sorted), K0 -> species tags -> K1 ->
    K2/K3 with the species-weighted plan.  Returns (species, counts, plan, nl, G, dG, flag)."""
    _lib.require_cuda()
    lib = _lib.load()
    species, counts, cart = _concat(cartesians)
    B, N, _ = cart.shape
    plan = SpeciesPlan(symm_funcs_data, species, cart.dtype, cart.device)
    posw = wrap_positions(cart, cell, pbc)
    tags = torch.repeat_interleave(torch.arange(len(species), dtype=torch.int32), torch.tensor(counts)).to(cart.device)
    with torch.cuda.device(cart.device):
        _lib.check(lib.nnmd_set_species(_lib.dtype_code(cart.dtype), _lib.ptr(posw), _lib.ptr(tags), B, N, _lib.stream_ptr()))
    nl = build_from_wrapped(posw, B, N, plan.rc_max)
    G, dG, flag = raw_descriptors(plan, nl, with_grad=with_grad)
    return species, counts, plan, nl, G, dG, flag


def calculate_sf_species(cartesians: dict, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, **kwargs):
    """Cross-species counterpart of `calculate_sf`: {species: (B, N_s, 3)} -> ({species: g (B, N_s, F)}, {species: dg (B, N_s, F, 3)}),
    g L2-normalised over the feature axis, dg un-normalised summed gradients, atoms of each species in the caller's order."""
    torch.manual_seed(42)  # the same RNG side effect as calculate_sf (calculate_sf.py:115)
    species, counts, plan, nl, G, dG, flag = raw_descriptors_species(cartesians, cell, pbc, symm_funcs_data)
    g = normalize(G, flag, out=G)
    raise_on_flag(flag)
    B = nl.n_frames
    g, dG = g.view(B, nl.n_atoms, plan.nf), dG.view(B, nl.n_atoms, plan.nf, 3)
    out_g, out_dg, off = {}, {}, 0
    for s, n in zip(species, counts):
        out_g[s], out_dg[s] = g[:, off:off + n].contiguous(), dG[:, off:off + n].contiguous()
        off += n
    return out_g, out_dg
