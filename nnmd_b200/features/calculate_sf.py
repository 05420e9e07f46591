"""`calculate_sf`: the reference's descriptor driver (nnmd/features/calculate_sf.py:23-136) on the B200 path.

Same signature, same outputs: g (n_batch, n_atoms, F) L2-normalised over F, dg (n_batch, n_atoms, F, 3) =
d(sum_i G_i,f)/d r_a un-normalised, both detached, same dtype and device as `cartesians`.
All frames go through ONE pass of K0 (wrap) -> K1 (cell list) -> K2/K3 (radial/angular + analytic gradient)
-> K5 (normalise + NaN guard); the reference's 2-frame chunk loop, per-function autograd passes and dense
(B,N,N,N) temporaries do not exist here.  One host sync for the pair total, one for the NaN flag.
"""
from __future__ import annotations

import ctypes
import warnings
import weakref

import numpy as np
import torch

from .. import _lib
from ..neighbor import NeighborList, build_from_wrapped, wrap_positions
from .symm_funcs import SymmetryFunction

_N_PARAMS = {1: 1, 2: 3, 4: 4, 5: 4}


class SfPlan:
    """Device-resident parameter table of one species (wraps nnmd_sf_plan_*)."""

    def __init__(self, features, params, dtype: torch.dtype, device: torch.device):
        kinds = []
        table = np.zeros((len(features), 4), dtype=np.float64)
        for f, (kind, row) in enumerate(zip(features, params)):
            kind = SymmetryFunction(int(kind)).value  # ValueError "3 is not a valid SymmetryFunction" like the enum
            row = list(row)
            if len(row) != _N_PARAMS[kind]:
                # the reference splats the row into g{k}_function(distances, *row): wrong arity is a TypeError
                raise TypeError(f"G{kind} takes {_N_PARAMS[kind]} parameters, got {len(row)}")
            table[f, :len(row)] = [float(v) for v in row]
            kinds.append(kind)
        if not kinds:
            raise ValueError("symm_funcs_data has no features")
        self.kinds = np.asarray(kinds, dtype=np.int32)
        self.table = table
        self.nf = len(kinds)
        self.dtype = dtype
        self.device = device
        lib = _lib.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.nnmd_sf_plan_create(_lib.dtype_code(dtype), self.nf, self.kinds.ctypes.data_as(ctypes.c_void_p),
                                               self.table.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle)))
        self.handle = handle
        self.rc_max = float(lib.nnmd_sf_plan_rc_max(handle))
        self._finalizer = weakref.finalize(self, lib.nnmd_sf_plan_destroy, handle)


_PLAN_CACHE: dict = {}
_PLAN_CACHE_MAX = 64


def get_plan(symm_funcs_data: dict, dtype: torch.dtype, device: torch.device) -> SfPlan:
    feats = tuple(int(k) for k in symm_funcs_data["features"])
    prm = tuple(tuple(float(v) for v in row) for row in symm_funcs_data["params"])
    dev = torch.device(device)
    key = (feats, prm, dtype, dev.index if dev.index is not None else torch.cuda.current_device())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = SfPlan(symm_funcs_data["features"], symm_funcs_data["params"], dtype, dev)
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:  # parameter scans: drop the oldest table (callers holding it keep it alive)
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        _PLAN_CACHE[key] = plan
    return plan


def edge_scratch(plan: SfPlan, nl: NeighborList, workspace=None):
    """Scratch for the edge-centric angular kernel (Phi/Gamma per owned edge), or None when the plan does not use it
    or the buffer does not fit: the C side then runs the vertex-centric kernel (no scratch, ~2x the arithmetic).
    workspace: a NeighborWorkspace that keeps the scratch between steps (static buffers for CUDA-graph replay)."""
    if nl.edge_offsets is None or nl.n_edges == 0:
        return None
    need = int(_lib.load().nnmd_sf_edge_scratch_bytes(plan.handle, nl.n_edges))
    if need == 0:
        return None
    try:
        if workspace is not None:
            return workspace.edge_scratch(need)
        return torch.empty(need, dtype=torch.uint8, device=nl.posw.device)
    except torch.cuda.OutOfMemoryError:
        warnings.warn(f"nnmd_b200: the {need / 2**30:.1f} GiB edge scratch of the edge-centric angular kernel does not fit; "
                      "falling back to the vertex-centric kernel (same results, about twice the arithmetic)", RuntimeWarning)
        return None


def raw_descriptors(plan: SfPlan, nl: NeighborList, with_grad: bool = True, edge: bool = True, workspace=None):
    """K2+K3(+K4a): un-normalised G (n_rows,F) and dG (n_rows,F,3) for the rows of `nl`, plus the device flag word.
    edge=False withholds the edge scratch, which makes the C side take the vertex-centric angular kernel.
    A capacity-bound list (NeighborWorkspace) brings its own sticky status word, which is used as the flag: the kernels
    return at once when it carries FLAG_OVERFLOW."""
    lib = _lib.load()
    dev = nl.posw.device
    n_rows = nl.n_rows
    G = torch.empty((n_rows, plan.nf), dtype=plan.dtype, device=dev)
    dG = torch.empty((n_rows, plan.nf, 3), dtype=plan.dtype, device=dev) if with_grad else None
    flag = nl.status if nl.status is not None else torch.zeros(1, dtype=torch.int32, device=dev)
    ews = edge_scratch(plan, nl, workspace) if edge else None
    with torch.cuda.device(dev):
        _lib.check(lib.nnmd_sf_eval(plan.handle, _lib.ptr(nl.posw), nl.n_frames, nl.n_atoms, nl.n_centres,
                                    _lib.ptr(nl.offsets), _lib.ptr(nl.idx), nl.max_row, _lib.ptr(nl.edge_offsets),
                                    nl.n_edges, _lib.ptr(ews), 0 if ews is None else ews.numel(), _lib.ptr(G),
                                    _lib.ptr(dG), _lib.ptr(flag), _lib.stream_ptr()))
    return G, dG, flag


def normalize(G: torch.Tensor, flag: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """K5: g / ||g||_2 over the feature axis; sets FLAG_NAN_G on 0/0 (calculate_sf.py:100-107)."""
    lib = _lib.load()
    out = torch.empty_like(G) if out is None else out
    with torch.cuda.device(G.device):
        _lib.check(lib.nnmd_sf_normalize(_lib.dtype_code(G.dtype), _lib.ptr(G), G.shape[0], G.shape[1], _lib.ptr(out),
                                         _lib.ptr(flag), _lib.stream_ptr()))
    return out


def raise_on_flag(flag: torch.Tensor) -> None:
    bits = int(flag.item())
    if bits & _lib.FLAG_OVERFLOW:
        raise _lib.NnmdError("the neighbour list outgrew its NeighborWorkspace (call workspace.validate() and repeat the step)")
    if bits & _lib.FLAG_NAN_G:
        raise ValueError("NaN values in the symmetry functions")
    if bits & _lib.FLAG_NAN_DG:
        raise ValueError("NaN values in the gradients of the symmetry functions")


def calculate_sf(cartesians: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, symm_funcs_data: dict, **kwargs):
    """Symmetry functions and their summed gradients for a batch of frames (reference: calculate_sf.py:23-136).

    Args:
        cartesians: (n_batch, n_atoms, 3) float32/float64.  A CPU tensor is moved to the current CUDA device and
            the results are returned on the CPU; the computation itself always runs on the GPU.
        cell: (3, 3) cell vectors; zero determinant = no cell.
        pbc: (3,) periodic flags (float or bool mask).
        symm_funcs_data: {"features": [1|2|4|5, ...], "params": [[...], ...]} (input_parser.py:103-107).
        disable_tqdm: accepted for compatibility (there is no per-chunk loop to report on).
    Returns:
        (g, dg): (n_batch, n_atoms, F) normalised descriptors, (n_batch, n_atoms, F, 3) gradients.
    Raises:
        ValueError("NaN values in the symmetry functions") for an atom without neighbours (0/0), like the reference.
    """
    _lib.require_cuda()
    # the reference reseeds the global RNG on every call (calculate_sf.py:16-20,115); later random splits and
    # weight initialisation of the samples depend on it, so the side effect is kept
    torch.manual_seed(42)
    if cartesians.dim() != 3 or cartesians.shape[-1] != 3:
        raise ValueError(f"cartesians must have shape (n_batch, n_atoms, 3), got {tuple(cartesians.shape)}")
    out_device = cartesians.device
    cart = cartesians.detach()
    if not cart.is_cuda:
        cart = cart.cuda()
    plan = get_plan(symm_funcs_data, cart.dtype, cart.device)
    B, N, _ = cart.shape
    posw = wrap_positions(cart, cell, pbc)
    nl = build_from_wrapped(posw, B, N, plan.rc_max)
    G, dG, flag = raw_descriptors(plan, nl, with_grad=True)
    g = normalize(G, flag, out=G)
    raise_on_flag(flag)
    g = g.view(B, N, plan.nf)
    dG = dG.view(B, N, plan.nf, 3)
    if out_device != g.device:
        g, dG = g.to(out_device), dG.to(out_device)
    return g, dG
