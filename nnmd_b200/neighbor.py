"""K0 + K1 host side: wrap-only PBC and the cell-list neighbour list (csrc/nlist.cu).

Replaces the dense `calculate_distances` + `_pairwise_filter` of the reference
(nnmd/features/symm_funcs.py:18-31, :53-104).  The pair set is `{(i,j): 0 < |p_j - p_i| < rc}` on the WRAPPED
positions with plain differences (no minimum image, no periodic images: SURVEY Q1/Q2).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class NeighborList:
    """CSR neighbour list over centre atoms; rows ascending; indices are global (frame*n_atoms + j)."""

    posw: torch.Tensor      # (n_frames*n_atoms, 4) wrapped positions, xyz + pad
    offsets: torch.Tensor   # int64 (n_rows+1)
    idx: torch.Tensor       # int32 (total)
    n_frames: int
    n_atoms: int
    n_centres: int
    rc: float
    total: int
    max_row: int
    edge_offsets: torch.Tensor | None = None   # int64 (n_rows+1): prefix sums of per-row neighbours with a larger index
    n_edges: int = 0

    @property
    def n_rows(self) -> int:
        return self.n_frames * self.n_centres

    def counts(self) -> torch.Tensor:
        return (self.offsets[1:] - self.offsets[:-1]).view(self.n_frames, self.n_centres)

    def pairs(self) -> torch.Tensor:
        """(total, 3) int64 rows (frame, i, j), lexicographically sorted -- the reference's `nonzero(mask)`."""
        cnt = (self.offsets[1:] - self.offsets[:-1])
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.idx.device), cnt)
        frame = rows // self.n_centres
        i = rows - frame * self.n_centres
        j = self.idx.to(torch.int64) - frame * self.n_atoms
        return torch.stack([frame, i, j], dim=1)


def _cell_host(cell: torch.Tensor, pbc: torch.Tensor, dtype: torch.dtype):
    """Cell, its inverse (computed in the WORKING dtype, like symm_funcs.py:74) and the pbc mask as host doubles."""
    c = cell.detach().to("cpu", dtype)
    has_cell = bool(torch.det(c) != 0)  # symm_funcs.py:71: zero determinant -> positions pass through
    if has_cell:
        inv = torch.linalg.inv(c)
    else:
        inv = torch.zeros(3, 3, dtype=dtype)
    p = pbc.detach().to("cpu", dtype)
    arr = lambda t: (ctypes.c_double * t.numel())(*[float(v) for v in t.reshape(-1).tolist()])
    return arr(c), arr(inv), arr(p), has_cell


def wrap_positions(cart: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor) -> torch.Tensor:
    """(..., 3) positions -> (n_total, 4) wrapped positions on the GPU (K0)."""
    _lib.require_cuda()
    lib = _lib.load()
    if not cart.is_cuda:
        raise _lib.NnmdError("wrap_positions needs a CUDA tensor")
    code = _lib.dtype_code(cart.dtype)
    pos = cart.detach().reshape(-1, 3).contiguous()
    n_total = pos.shape[0]
    posw = torch.empty((n_total, 4), dtype=cart.dtype, device=cart.device)
    c, inv, p, has_cell = _cell_host(cell, pbc, cart.dtype)
    with torch.cuda.device(cart.device):
        _lib.check(lib.nnmd_wrap_positions(code, _lib.ptr(pos), n_total, c, inv, p, int(has_cell), _lib.ptr(posw),
                                           _lib.stream_ptr()))
    return posw


def build_from_wrapped(posw: torch.Tensor, n_frames: int, n_atoms: int, rc: float, n_centres: int | None = None) -> NeighborList:
    """K1 on already wrapped (n_total,4) positions.  One host sync (the pair total sizes the index array)."""
    lib = _lib.load()
    code = _lib.dtype_code(posw.dtype)
    n_centres = n_atoms if n_centres is None else int(n_centres)
    n_rows = n_frames * n_centres
    dev = posw.device
    with torch.cuda.device(dev):
        ws_bytes = int(lib.nnmd_nlist_workspace_bytes(n_frames, n_atoms))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        offsets = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        edge_offsets = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        stats = torch.empty(3, dtype=torch.int64, device=dev)
        st = _lib.stream_ptr()
        _lib.check(lib.nnmd_nlist_count(code, _lib.ptr(posw), n_frames, n_atoms, n_centres, float(rc), _lib.ptr(ws),
                                        _lib.ptr(offsets), _lib.ptr(edge_offsets), _lib.ptr(stats), st))
        total, max_row, n_edges = (int(v) for v in stats.tolist())  # the one sync of the build
        idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.nnmd_nlist_fill(code, _lib.ptr(posw), n_frames, n_atoms, n_centres, float(rc), _lib.ptr(ws),
                                       _lib.ptr(offsets), max_row, _lib.ptr(idx), st))
    return NeighborList(posw, offsets, idx[:total] if total > 0 else idx[:0], n_frames, n_atoms, n_centres, float(rc),
                        total, max_row, edge_offsets, n_edges)


def build_neighbor_list(cart: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, rc: float,
                        n_centres: int | None = None) -> NeighborList:
    """cart (B,N,3) or (N,3) on the GPU -> neighbour list inside `rc` (K0 + K1)."""
    if cart.dim() == 2:
        cart = cart.unsqueeze(0)
    if cart.dim() != 3 or cart.shape[-1] != 3:
        raise ValueError(f"cartesians must have shape (n_batch, n_atoms, 3), got {tuple(cart.shape)}")
    posw = wrap_positions(cart, cell, pbc)
    return build_from_wrapped(posw, cart.shape[0], cart.shape[1], rc, n_centres)
