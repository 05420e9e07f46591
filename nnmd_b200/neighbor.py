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
    # capacity-bound lists (NeighborWorkspace): `total`, `max_row`, `n_edges` above are CAPACITIES, the true values sit in
    # `stats` on the device, and `status` is the sticky flag word every later kernel of the step reads and ORs into
    status: torch.Tensor | None = None         # int32 (1,) device: NNMD_FLAG_* bits
    stats: torch.Tensor | None = None          # int64 (6,) device: total pairs, longest row, owned edges (latest | first overflow)

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
        j = self.idx[:rows.numel()].to(torch.int64) - frame * self.n_atoms  # capacity-bound lists: idx is longer than the list
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


class NeighborWorkspace:
    """Persistent buffers of a capacity-bound neighbour build: K1 with NO host synchronisation (nnmd_nlist_build).

    The first `build` sizes the buffers from an exact count (one sync, like `build_from_wrapped`) times `margin`; every
    later `build` only enqueues kernels.  A structure that outgrows a capacity raises NNMD_FLAG_OVERFLOW in `status`
    instead of writing out of bounds, the K2..K4 kernels of that step then do nothing, and `validate()` -- called whenever
    the caller synchronises anyway, e.g. when it reads the NaN flag -- regrows the buffers and asks for the step again.
    skin > 0: the list is built with cutoff rc + skin and REUSED until an atom has moved more than skin / 2 (decided on the
    device by nnmd_skin_check, no sync); the descriptor kernels re-test every neighbour against their own cutoffs, so the
    results do not depend on the skin.  All buffers are static, so a whole MD step can be captured in a CUDA graph.
    """

    def __init__(self, n_frames: int, n_atoms: int, rc: float, dtype: torch.dtype, device, n_centres: int | None = None,
                 skin: float = 0.0, margin: float = 1.25):
        self.lib = _lib.load()
        self.n_frames, self.n_atoms = int(n_frames), int(n_atoms)
        self.n_centres = self.n_atoms if n_centres is None else int(n_centres)
        self.rc, self.skin, self.margin = float(rc), float(skin), float(margin)
        self.dtype, self.device = dtype, torch.device(device)
        self.code = _lib.dtype_code(dtype)
        dev = self.device
        n_rows = self.n_frames * self.n_centres
        with torch.cuda.device(dev):
            self.ws = torch.empty(max(int(self.lib.nnmd_nlist_workspace_bytes(self.n_frames, self.n_atoms)), 1), dtype=torch.uint8, device=dev)
        self.offsets = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        self.edge_offsets = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
        self.stats = torch.zeros(6, dtype=torch.int64, device=dev)   # latest build | first overflowing build
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.go = torch.ones(1, dtype=torch.int32, device=dev)
        self.acc = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_builds = torch.zeros(1, dtype=torch.int64, device=dev)
        self.pos_ref = torch.full((self.n_frames * self.n_atoms, 4), float("nan"), dtype=dtype, device=dev) if self.skin > 0 else None
        self.idx = None
        self.idx_cap = self.max_row_cap = self.edge_cap = 0
        self._scratch = {}

    # ---- capacities
    def _size_from(self, total: int, max_row: int, n_edges: int) -> None:
        m = self.margin
        self.idx_cap = max(int(total * m) + 1024, 1)
        # rows get 10 % head room only: the longest row sizes the per-warp shared memory of K2/K3 (a 25 % margin on the
        # 80-entry rows of bulk Cu would cost the angular kernel one of its four resident blocks per SM)
        self.max_row_cap = (max(max_row + max(4, (max_row + 9) // 10), 8) + 3) & ~3
        self.edge_cap = max(int(n_edges * m) + 512, 1)
        self.idx = torch.empty(self.idx_cap, dtype=torch.int32, device=self.device)
        n_rows = self.n_frames * self.n_centres
        nbytes = int(self.lib.nnmd_nlist_row_scratch_bytes(n_rows, self.max_row_cap))
        # padded scratch rows of the single-pass build (skipped when it would be unreasonably large: many short frames)
        self.row_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self.device) if 0 < nbytes <= (8 << 30) else None
        self._scratch = {}

    def _enqueue(self, posw: torch.Tensor, go) -> None:
        with torch.cuda.device(self.device):
            _lib.check(self.lib.nnmd_nlist_build(self.code, _lib.ptr(posw), self.n_frames, self.n_atoms, self.n_centres,
                                                 self.rc + self.skin, _lib.ptr(self.ws), _lib.ptr(self.offsets), _lib.ptr(self.edge_offsets),
                                                 _lib.ptr(self.idx), self.idx_cap, self.max_row_cap, self.edge_cap,
                                                 _lib.ptr(self.row_scratch), _lib.ptr(self.stats),
                                                 _lib.ptr(self.status), _lib.ptr(go), _lib.stream_ptr()))

    def build(self, posw: torch.Tensor) -> NeighborList:
        """Neighbour list of the wrapped positions (n_total, 4).  Sync-free after the first call."""
        if posw.dtype != self.dtype or posw.shape[0] != self.n_frames * self.n_atoms:
            raise ValueError("NeighborWorkspace.build: positions do not match the workspace (dtype / atom count)")
        if self.idx is None:  # first build: exact sizes (one sync), then capacities with margin
            nl = build_from_wrapped(posw, self.n_frames, self.n_atoms, self.rc + self.skin, self.n_centres)
            self._size_from(nl.total, nl.max_row, nl.n_edges)
            self._enqueue(posw, None)
            if self.pos_ref is not None:
                self.pos_ref.copy_(posw)
            self.n_builds += 1
        elif self.skin > 0:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.nnmd_skin_check(self.code, _lib.ptr(posw), _lib.ptr(self.pos_ref), posw.shape[0], self.skin,
                                                    _lib.ptr(self.acc), _lib.ptr(self.go), _lib.ptr(self.n_builds), _lib.stream_ptr()))
            self._enqueue(posw, self.go)
        else:
            self._enqueue(posw, None)
        return NeighborList(posw, self.offsets, self.idx, self.n_frames, self.n_atoms, self.n_centres, self.rc + self.skin,
                            self.idx_cap, self.max_row_cap, self.edge_offsets, self.edge_cap, self.status, self.stats)

    def edge_scratch(self, nbytes: int) -> torch.Tensor:
        """Static scratch of the edge-centric angular kernel for the current capacities (reused across steps)."""
        buf = self._scratch.get(nbytes)
        if buf is None:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._scratch = {nbytes: buf}
        return buf

    def validate(self) -> bool:
        """One host sync: True when every build since the last call fit.  On overflow the buffers are regrown from the
        recorded demand, the flag is cleared and False is returned: the caller repeats the step(s)."""
        bits = int(self.status.item())
        if not bits & _lib.FLAG_OVERFLOW:
            return True
        total, max_row, n_edges = (int(v) for v in self.stats[3:].tolist())  # demand of the first build that did not fit
        self._size_from(max(total, self.idx_cap), max(max_row, self.max_row_cap), max(n_edges, self.edge_cap))
        self.status.zero_()  # NaN bits raised by the steps that ran on the incomplete list mean nothing; the repeat re-raises real ones
        if self.pos_ref is not None:
            self.pos_ref.fill_(float("nan"))  # force the next skin check to rebuild
        return False

    def true_sizes(self):
        """(total pairs, longest row, owned edges) of the last build: one host sync."""
        return tuple(int(v) for v in self.stats[:3].tolist())
