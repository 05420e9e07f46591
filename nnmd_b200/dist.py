"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests).

Two ways the path shards:
  * spatial  -- one large structure is cut into x-slabs; every rank owns the atoms of its slab and imports a HALO
    of width rc from the slabs next to it.  Because every contributing triplet is a closed triangle
    (SURVEY 7.2), G_a, dG[a] and the reference-semantic force of an owned atom depend only on atoms within rc
    of it: ONE exchange of positions per step, no reverse (force) communication.  Wrap-only PBC
    means no halo crosses a periodic face.
  * frames   -- training sets shard the frame axis; the only collective is the all-reduce of the flat MLP
    gradient buffer (`allreduce_gradients`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as td

from .features.symm_funcs import map_positions_pbc


def _world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


class PeerWindow:
    """A symmetric device window mapped into every rank of the job (CUDA IPC over NVLink; csrc/peer.cu): the collectives of
    the path run on it as ordinary kernels, so they are captured into the CUDA graph of the step they belong to.

    Collective constructor: every rank must create its windows in the same order.  `PeerWindow.create` returns None (after
    saying so once) where peer windows cannot be had -- CPU tensors / gloo tests, one rank, ranks on different hosts, or an
    IPC mapping the platform refuses -- and the callers fall back to torch.distributed (NCCL)."""

    _warned = False

    def __init__(self, handle, lib, cap_bytes: int):
        self.handle, self._lib, self.cap_bytes = handle, lib, cap_bytes

    @classmethod
    def create(cls, cap_bytes: int, device) -> "PeerWindow | None":
        import ctypes
        import os
        from . import _lib
        rank, world = _world()
        device = torch.device(device)
        if world == 1 or device.type != "cuda" or os.environ.get("NNMD_NO_PEER", "0") == "1":
            return None
        lib = _lib.load()
        handle, blob, err = ctypes.c_void_p(), b"", None
        try:
            with torch.cuda.device(device):
                _lib.check(lib.nnmd_peer_create(rank, world, int(cap_bytes), ctypes.byref(handle)))
                buf = ctypes.create_string_buffer(lib.nnmd_peer_handle_bytes())
                _lib.check(lib.nnmd_peer_export(handle, buf))
                blob = buf.raw
        except _lib.NnmdError as exc:
            err = str(exc)
        # one all-gather carries the handles AND the verdicts: either every rank connects or none does
        gathered = [None] * world
        td.all_gather_object(gathered, (blob, err, os.uname().nodename))
        ok = all(g[1] is None for g in gathered) and len({g[2] for g in gathered}) == 1
        if ok:
            try:
                with torch.cuda.device(device):
                    _lib.check(lib.nnmd_peer_connect(handle, b"".join(g[0] for g in gathered)))
            except _lib.NnmdError as exc:
                err = str(exc)
        verdicts = [None] * world
        td.all_gather_object(verdicts, err)
        if not ok or any(v is not None for v in verdicts):
            if handle:
                lib.nnmd_peer_destroy(handle)
            if rank == 0 and not cls._warned:
                cls._warned = True
                why = next((v for v in verdicts if v), None) or next((g[1] for g in gathered if g[1]), "ranks on different hosts")
                import sys
                print(f"nnmd_b200: peer windows unavailable ({why}); using NCCL for the exchanges", file=sys.stderr, flush=True)
            return None
        return cls(handle, lib, int(lib.nnmd_peer_capacity(handle)))

    def status(self) -> int:
        """0, or 1 when a peer failed to arrive in some call since creation (that call's result is garbage).  Synchronises."""
        import ctypes
        from . import _lib
        v = ctypes.c_int(0)
        _lib.check(self._lib.nnmd_peer_status(self.handle, ctypes.byref(v)))
        return int(v.value)

    def check(self) -> None:
        if self.status():
            from . import _lib
            raise _lib.NnmdError("a peer did not arrive at a peer-window exchange within the spin limit")

    def allreduce_weighted(self, buf: torch.Tensor) -> None:
        from . import _lib
        with torch.cuda.device(buf.device):
            _lib.check(self._lib.nnmd_peer_allreduce_weighted_f32(self.handle, _lib.ptr(buf), buf.numel(), _lib.stream_ptr()))

    def allreduce_step(self, buf: torch.Tensor, weight: float, loss_in=None, loss_acc=None) -> None:
        from . import _lib
        with torch.cuda.device(buf.device):
            _lib.check(self._lib.nnmd_peer_allreduce_step_f32(self.handle, _lib.ptr(buf), buf.numel(), float(weight), _lib.ptr(loss_in),
                                                              _lib.ptr(loss_acc), _lib.stream_ptr()))

    def allreduce_sum(self, buf: torch.Tensor) -> None:
        from . import _lib
        with torch.cuda.device(buf.device):
            _lib.check(self._lib.nnmd_peer_allreduce_sum_f32(self.handle, _lib.ptr(buf), buf.numel(), _lib.stream_ptr()))

    def halo_exchange(self, own: torch.Tensor, send_idx, dst_peer, dst_slot, n_halo: int, local: torch.Tensor) -> None:
        from . import _lib
        width = own.shape[1] * own.element_size() // 4
        with torch.cuda.device(own.device):
            _lib.check(self._lib.nnmd_peer_halo_exchange(self.handle, _lib.ptr(own), own.shape[0], _lib.ptr(send_idx), _lib.ptr(dst_peer),
                                                         _lib.ptr(dst_slot), 0 if send_idx is None else send_idx.numel(), int(n_halo),
                                                         width, _lib.ptr(local), _lib.stream_ptr()))

    def __del__(self):
        try:
            if self.handle:
                self._lib.nnmd_peer_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_SMALL_WINDOW: dict = {}


def allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    """In-place sum of a small float32 CUDA tensor over the ranks (e.g. the total energy of a sharded structure): one kernel
    over a peer window where available (first call is collective: it creates the window), NCCL / gloo otherwise."""
    rank, world = _world()
    if world == 1:
        return t
    if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() <= 4096:
        key = t.device.index
        if key not in _SMALL_WINDOW:
            _SMALL_WINDOW[key] = PeerWindow.create(4 * (4096 + 64), t.device)
        if _SMALL_WINDOW[key] is not None:
            _SMALL_WINDOW[key].allreduce_sum(t)
            return t
    td.all_reduce(t)
    return t


@dataclass
class SlabShard:
    """Ownership and halo plan of one rank for a decomposition into x-slabs or, with `grid`, into px x py x pz boxes."""

    rank: int
    world: int
    rc: float
    n_global: int
    owned_idx: torch.Tensor                     # (n_owned,) global atom indices, ascending
    owned_pos: torch.Tensor                     # (n_owned, 3) ORIGINAL (unwrapped) positions
    send_idx: list = field(default_factory=list)  # per peer: indices into `owned_*` to send (None: nothing)
    edges: torch.Tensor | None = None           # (world+1,) slab boundaries on the wrapped x axis
    recv_counts: list | None = None             # atoms each peer sends here; learned once per send plan (no per-step sync)
    grid: tuple = (1, 1, 1)                     # boxes per axis (x-slabs: (world, 1, 1))
    edges3: list | None = None                  # per axis: (grid[ax]+1,) boundaries

    @property
    def n_owned(self) -> int:
        return int(self.owned_idx.numel())

    @staticmethod
    def slab_edges(x_wrapped: torch.Tensor, world: int, rc: float | None = None) -> torch.Tensor:
        """Slab boundaries on the wrapped x coordinate: edges[0] = -inf, edges[world] = +inf.

        rc=None: equal atom counts.  rc given: equal WORK.  The reference never adds periodic images, so the
        structure ends in free surfaces at both ends of the x axis and an atom closer than rc to one of them sees only the
        fraction f(h) = 1 - (1 - h/rc)^2 (2 + h/rc) / 4 of its cutoff sphere (h = distance to the face); its angular work
        (owned edges x closing third atoms) scales like f^2.  Boundaries are placed at equal cumulative f^2, which evens out
        the ~6 % by which equal-count end slabs are cheaper than interior ones.
        """
        xs = torch.sort(x_wrapped.double()).values
        n = xs.numel()
        if rc is None or n == 0:
            inner = [float(xs[(k * n) // world]) for k in range(1, world)]
        else:
            t = (torch.minimum(xs - xs[0], xs[-1] - xs) / float(rc)).clamp(max=1.0)
            f = 1.0 - (1.0 - t) ** 2 * (2.0 + t) / 4.0
            cw = torch.cumsum(f * f, 0)
            targets = cw[-1] * torch.arange(1, world, dtype=torch.float64) / world
            pos = torch.searchsorted(cw, targets).clamp(max=n - 1)
            inner = [float(xs[int(p)]) for p in pos]
        return torch.tensor([-float("inf")] + inner + [float("inf")], dtype=torch.float64)

    @staticmethod
    def grid_for(world: int, lengths=(1.0, 1.0, 1.0)) -> tuple:
        """(px, py, pz) with px*py*pz = world that minimises the total area of the internal boundaries of a box with the
        given side lengths: the duplicated owner<->halo edge work and the halo volume both scale with that area
        (8 ranks on a cube: 2x2x2 has 3 L^2 of boundary, 8 x-slabs have 7 L^2)."""
        lx, ly, lz = (float(v) for v in lengths)
        best, best_area = (world, 1, 1), None
        for px in range(1, world + 1):
            if world % px:
                continue
            for py in range(1, world // px + 1):
                if (world // px) % py:
                    continue
                pz = world // (px * py)
                area = (px - 1) * ly * lz + (py - 1) * lx * lz + (pz - 1) * lx * ly
                if best_area is None or area < best_area - 1e-12:
                    best, best_area = (px, py, pz), area
        return best

    @classmethod
    def from_global(cls, pos: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, rc: float, rank: int, world: int,
                    balance: str = "work", grid=None):
        """Domain decomposition of a structure every rank can see (synthetic benchmarks, file-based inputs).
        balance: "work" (default, see `slab_edges`) or "count" (equal atom counts).
        grid: None = `world` x-slabs; (px, py, pz) = boxes, rank = (ix*py + iy)*pz + iz, the boundaries of every axis placed
        by `slab_edges` on that axis' wrapped coordinate; "auto" = `grid_for(world, extent of the structure)`.

        Ownership is decided on the WRAPPED coordinates (the ones the kernels measure distances in);
        the positions that travel are the original ones, so every rank's K0 reproduces the single-GPU wrap bit for bit.
        """
        pos = pos.detach().cpu()
        pw = map_positions_pbc(pos, cell.detach().cpu().to(pos.dtype), pbc.detach().cpu().to(pos.dtype)).double()
        if balance not in ("work", "count"):
            raise ValueError("balance must be 'work' or 'count'")
        if grid is None:
            grid = (world, 1, 1)
        elif isinstance(grid, str):
            if grid != "auto":
                raise ValueError("grid must be None, 'auto' or (px, py, pz)")
            ext = (pw.max(0).values - pw.min(0).values) if pw.shape[0] else torch.ones(3, dtype=torch.float64)
            grid = cls.grid_for(world, [float(v) for v in ext])
        grid = tuple(int(v) for v in grid)
        if len(grid) != 3 or grid[0] * grid[1] * grid[2] != world:
            raise ValueError(f"grid {grid} does not have {world} boxes")
        edges = [cls.slab_edges(pw[:, ax], grid[ax], rc if balance == "work" else None) for ax in range(3)]
        coords = lambda r: (r // (grid[1] * grid[2]), (r // grid[2]) % grid[1], r % grid[2])
        inside = lambda r, margin: torch.stack([(pw[:, ax] >= edges[ax][c] - margin) & (pw[:, ax] < edges[ax][c + 1] + margin)
                                               for ax, c in enumerate(coords(r))]).all(0)
        mine = torch.nonzero(inside(rank, 0.0)).flatten()
        send = []
        for peer in range(world):
            if peer == rank:
                send.append(None)
                continue
            sel = torch.nonzero(inside(peer, rc)[mine]).flatten()
            send.append(sel if sel.numel() else None)
        shard = cls(rank, world, float(rc), pos.shape[0], mine, pos[mine].contiguous(), send, edges[0])
        shard.grid, shard.edges3 = grid, edges
        return shard

    def to(self, device) -> "SlabShard":
        """Move the owned positions and the send plan to `device` (done once; the per-step exchange then stays on it)."""
        device = torch.device(device)
        self.owned_pos = self.owned_pos.to(device)
        self.send_idx = [None if s is None else s.to(device) for s in self.send_idx]
        return self

    def set_owned_positions(self, pos: torch.Tensor) -> None:
        """New positions of the owned atoms (same order), e.g. after an MD step or an H2D copy."""
        self.owned_pos = pos

    def exchange_halo(self, device=None) -> torch.Tensor:
        """Positions this rank computes on: owned atoms first, then the halo atoms received from the peers.

        world == 1: just the owned positions on `device`.  Otherwise one batched send/recv round (NCCL on GPUs; gloo when
        the tensors live on the CPU) straight into a persistent local buffer: the peers' counts are learned once per send
        plan (one all-gather), after which a step issues no host synchronisation and no concatenation.
        """
        device = torch.device(device) if device is not None else self.owned_pos.device
        own = self.owned_pos.to(device)
        if self.world == 1:
            return own
        if self.recv_counts is None:  # the send plan is fixed between re-decompositions: learn the peers' counts once
            counts = torch.tensor([0 if s is None else int(s.numel()) for s in self.send_idx], dtype=torch.int64, device=device)
            all_counts = [torch.empty_like(counts) for _ in range(self.world)]
            td.all_gather(all_counts, counts)
            self.recv_counts = [int(all_counts[peer][self.rank]) for peer in range(self.world)]
            self._local = None
            self._window = None
            if device.type == "cuda" and own.dtype in (torch.float32, torch.float64):
                # peer-memory route: every sender writes its rows straight into the receiver's halo (same order as the
                # send/recv route: blocks by ascending sender rank), three small kernels per step and no NCCL call
                cm = torch.stack(all_counts).cpu()                     # cm[q][p] = rows q sends to p
                row_bytes = own.shape[1] * own.element_size()
                self._window = PeerWindow.create(max(int(cm.sum(0).max()) * row_bytes, 256), device)
                if self._window is not None:
                    idx, peers, slots = [], [], []
                    for peer in range(self.world):
                        if peer == self.rank or self.send_idx[peer] is None:
                            continue
                        k = int(self.send_idx[peer].numel())
                        base = int(cm[:self.rank, peer].sum())          # my block starts after the lower-ranked senders'
                        idx.append(self.send_idx[peer].to(device=device, dtype=torch.int64))
                        peers.append(torch.full((k,), peer, dtype=torch.int32, device=device))
                        slots.append(torch.arange(base, base + k, dtype=torch.int64, device=device))
                    self._push = (torch.cat(idx), torch.cat(peers), torch.cat(slots)) if idx else (None, None, None)
        n_own = own.shape[0]
        local = getattr(self, "_local", None)
        if local is None or local.device != device or local.dtype != own.dtype or local.shape[0] != n_own + sum(self.recv_counts):
            local = self._local = torch.empty((n_own + sum(self.recv_counts), 3), dtype=own.dtype, device=device)
        if getattr(self, "_window", None) is not None:
            self._window.halo_exchange(own.contiguous(), *self._push, sum(self.recv_counts), local)
            return local
        local[:n_own].copy_(own)
        ops, keep, off = [], [], n_own
        for peer in range(self.world):
            if peer == self.rank:
                continue
            if self.send_idx[peer] is not None:
                buf = own.index_select(0, self.send_idx[peer].to(device))
                keep.append(buf)
                ops.append(td.P2POp(td.isend, buf, peer))
            if self.recv_counts[peer] > 0:
                ops.append(td.P2POp(td.irecv, local[off:off + self.recv_counts[peer]], peer))
                off += self.recv_counts[peer]
        if ops:
            for req in td.batch_isend_irecv(ops):
                req.wait()
        return local


def shard_frames(n_frames: int, rank: int | None = None, world: int | None = None) -> range:
    """Contiguous block of frame indices owned by `rank` (training / descriptor-cache sharding)."""
    if rank is None or world is None:
        rank, world = _world()
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return range(lo, hi)


class FlatGradients:
    """All parameter gradients of the atomic networks as views of ONE contiguous buffer, so that the per-step
    collective is a single in-place all-reduce with no flatten / unflatten copies (the buffer is <= ~30k floats for the
    shipped nets: the exchange is latency-bound).  One extra slot behind the gradients carries the number of frames of
    the rank's batch through the same all-reduce: the global-batch gradient is sum_r b_r grad_r / sum_r b_r, which is
    exact for ragged shards and lets a rank that ran out of batches take part with b_r = 0."""

    def __init__(self, params):
        self.params = [p for p in params]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.buf = torch.zeros(n + 1, dtype=ref.dtype, device=ref.device)
        self.flat = self.buf[:n]
        self.count = self.buf[n:]  # (1,) frames behind the gradient: local before, global after `allreduce`
        self.window, self._window_tried = None, False  # PeerWindow of the buffer's size, made at the first multi-rank all-reduce
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()
        for p, g in zip(self.params, self.views()):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g  # an optimizer.zero_grad(set_to_none=True) elsewhere dropped the view

    def views(self):
        off = 0
        for p in self.params:
            yield self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def world_size(self) -> int:
        return _world()[1]

    def allreduce(self, n_frames: int | None = None, loss_in=None, loss_acc=None) -> bool:
        """Global-batch gradient from per-rank batch-mean gradients.  n_frames = frames of THIS rank's batch (0 when the
        rank has none left this step; None = equal batches everywhere, plain average).  No-op on one rank.
        loss_in / loss_acc (float32 (3,) on the device): on the peer-window route the kernel also adds this rank's share of the
        global-batch loss, loss_in * b_r / sum_r b_r, to loss_acc; returns True when it did (else the caller does)."""
        rank, world = _world()
        if world == 1:
            return False
        if n_frames is None:
            td.all_reduce(self.flat, op=td.ReduceOp.SUM)
            self.flat /= world
            return False
        if self.window is None and not self._window_tried:
            self._window_tried = True
            if self.buf.dtype == torch.float32:
                self.window = PeerWindow.create(4 * (self.buf.numel() + 64), self.buf.device)
        if self.window is not None:  # one kernel over NVLink peer memory, capturable into the step's CUDA graph
            self.window.allreduce_step(self.buf, float(n_frames), loss_in, loss_acc)
            return loss_acc is not None
        self.count.fill_(float(n_frames))
        self.flat *= float(n_frames)
        td.all_reduce(self.buf, op=td.ReduceOp.SUM)
        self.flat /= self.count
        return False

    def graph_capturable(self) -> bool:
        """True when `allreduce` is plain kernels on the current stream (one rank, or a connected peer window)."""
        return _world()[1] == 1 or self.window is not None


def allreduce_gradients(params, average: bool = True) -> None:
    """One all-reduce of the flat gradient buffer of all atomic networks (<= ~30k floats for the shipped nets:
    latency-bound, so everything goes in a single message)."""
    rank, world = _world()
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    td.all_reduce(flat, op=td.ReduceOp.SUM)
    if average:
        flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
