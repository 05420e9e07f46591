"""world_size-2 gloo tests of the multi-GPU host logic (spatial slab sharding with halo exchange, frame sharding,
gradient all-reduce).  CPU only: the kernels are not involved, the checker is the O(N) oracle's pair set."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from nnmd_b200 import dist
    from nnmd_b200.workloads import fcc_cell
    pos, cell, pbc = fcc_cell(6, dtype=torch.float64)
    pos = pos - 3.0  # some atoms start outside the cell: ownership must follow the WRAPPED coordinate
    shard = dist.SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=rank, world=world)
    local = shard.exchange_halo("cpu")
    # gradient all-reduce
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(rank + 1))
    dist.allreduce_gradients([p])
    q1, q2 = torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4))
    fg = dist.FlatGradients([q1, q2])
    fg.zero()
    q1.grad += float(rank + 1)
    q2.grad += 10.0 * (rank + 1)
    fg.allreduce()
    assert q1.grad.data_ptr() == fg.flat.data_ptr() and torch.allclose(q1.grad, torch.full((2, 3), 1.5))
    assert torch.allclose(q2.grad, torch.full((4,), 15.0))
    # ragged shards: rank 0 holds a batch of 3 frames, rank 1 a batch of 1 -> global-batch gradient (3*1 + 1*2)/4; and a
    # step in which rank 1 has no batch left (weight 0) must complete on both ranks with rank 0's gradient alone
    fg.zero()
    q1.grad += float(rank + 1)
    fg.allreduce(3 if rank == 0 else 1)
    assert torch.allclose(q1.grad, torch.full((2, 3), 1.25)) and float(fg.count) == 4.0
    fg.zero()
    if rank == 0:
        q2.grad += 7.0
    fg.allreduce(2 if rank == 0 else 0)
    assert torch.allclose(q2.grad, torch.full((4,), 7.0)) and float(fg.count) == 2.0
    # per-epoch step count: 127 vs 128 frames in batches of 101 is 2 steps everywhere (a rank with an extra
    # batch used to hang in the all-reduce)
    from nnmd_b200.nn import BPNN
    bp = BPNN(use_cuda=False)
    n_local = -(-int((127 + rank) * 0.8) // 101)
    assert bp._epoch_steps(n_local) == 2

    class _Loader:  # the answer is remembered per loader object: the second epoch issues no collective (and no host sync)
        pass
    ld = _Loader()
    assert bp._epoch_steps(n_local, ld) == 2 and bp._steps_cache[ld] == (n_local, 2)
    calls = []
    orig = td.all_reduce
    td.all_reduce = lambda *a, **k: calls.append(1) or orig(*a, **k)
    try:
        assert bp._epoch_steps(n_local, ld) == 2 and not calls
        assert bp._epoch_steps(n_local + 1, ld) == 3 and len(calls) == 1   # a changed length asks again (on every rank)
    finally:
        td.all_reduce = orig
    frames = list(dist.shard_frames(11))
    # dataset build on pre-sharded frames: the target scaling uses the GLOBAL extrema (one MIN and one MAX all-reduce)
    from nnmd_b200.nn.dataset import cache_name, scale_targets
    e_all = torch.tensor([[3.0], [-1.0], [7.0], [2.0], [5.0]])
    mine = dist.shard_frames(5)
    e_s, f_s, emin, emax = scale_targets(e_all[mine.start:mine.stop], torch.ones(len(mine), 2, 3), presharded=True)
    assert float(emin) == -1.0 and float(emax) == 7.0
    assert torch.allclose(e_s, (e_all[mine.start:mine.stop] + 1.0) / 8.0) and torch.allclose(f_s, torch.full_like(f_s, 0.125))
    assert cache_name("dg", "Cu", rank, world) == f"dg_Cu.rank{rank}of2.pt" and cache_name("g", "Cu", 0, 1) == "g_Cu.pt"
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), owned=shard.owned_idx.numpy(), local=local.numpy(),
             grad=p.grad.numpy(), frames=np.asarray(frames))
    td.destroy_process_group()


def test_slab_shard_halo_and_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from nnmd_b200.features.symm_funcs import map_positions_pbc
    from nnmd_b200.workloads import fcc_cell
    from oracle import nl_oracle as O3
    pos, cell, pbc = fcc_cell(6, dtype=torch.float64)
    pos = pos - 3.0
    n = pos.shape[0]
    pairs = O3.pairs(pos, cell, pbc, 6.0)
    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    owned = [set(p["owned"].tolist()) for p in parts]
    assert owned[0].isdisjoint(owned[1]) and len(owned[0] | owned[1]) == n
    assert abs(len(owned[0]) - len(owned[1])) <= n // 50  # two end slabs: equal work is (nearly) equal count
    # equal-count and equal-work boundaries: 4 slabs of a cell with free surfaces -> the end slabs get more atoms
    from nnmd_b200.dist import SlabShard
    xw4 = map_positions_pbc(pos, cell, pbc)[:, 0]
    ec = SlabShard.slab_edges(xw4, 4)
    ew = SlabShard.slab_edges(xw4, 4, rc=6.0)
    cnt = lambda e: [int(((xw4 >= e[k]) & (xw4 < e[k + 1])).sum()) for k in range(4)]
    cc, cwk = cnt(ec), cnt(ew)
    assert sum(cc) == sum(cwk) == n and max(cc) - min(cc) <= n // 20
    assert cwk[0] > cwk[1] and cwk[3] > cwk[2]
    for r, p in enumerate(parts):
        local = torch.from_numpy(p["local"])
        n_own = len(p["owned"])
        # owned atoms come first, in ascending global order, with their ORIGINAL positions
        assert torch.equal(local[:n_own], pos[torch.from_numpy(p["owned"])])
        # every neighbour (inside rc, reference pair set) of an owned atom is present locally
        have = {tuple(np.round(v, 9)) for v in p["local"]}
        need = {int(j) for i, j in pairs if int(i) in owned[r]}
        for j in need:
            assert tuple(np.round(pos[j].numpy(), 9)) in have
        # the halo is a strict subset of the other rank's atoms (no duplicates of owned atoms)
        assert local.shape[0] < n and local.shape[0] - n_own > 0
        assert np.allclose(p["grad"], 1.5)
    assert sorted(parts[0]["frames"].tolist() + parts[1]["frames"].tolist()) == list(range(11))
    # slabs are cut on the wrapped coordinate
    xw = map_positions_pbc(pos, cell, pbc)[:, 0]
    assert xw[list(owned[0])].max() <= xw[list(owned[1])].min()


def test_box_decomposition_partitions_and_covers_every_neighbour():
    """grid="auto": 8 ranks on a cube become 2x2x2 boxes (3 L^2 of internal boundary instead of the 7 L^2 of x-slabs); the
    boxes partition the atoms, and the rows the peers send to a rank contain every neighbour (reference pair set) of its owned
    atoms.  The halo volume is what the duplicated owner<->halo edge work scales with: it must shrink against x-slabs."""
    from nnmd_b200.dist import SlabShard
    from nnmd_b200.workloads import fcc_cell
    from oracle import nl_oracle as O3
    assert SlabShard.grid_for(8) == (2, 2, 2) and SlabShard.grid_for(2, (1.0, 1.0, 3.0)) == (1, 1, 2)
    assert SlabShard.grid_for(8, (100.0, 1.0, 1.0)) == (8, 1, 1)
    pos, cell, pbc = fcc_cell(8, dtype=torch.float64)
    pos = pos - 2.0
    n = pos.shape[0]
    pairs = O3.pairs(pos, cell, pbc, 6.0)
    world = 8
    boxes = [SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=r, world=world, grid="auto") for r in range(world)]
    slabs = [SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=r, world=world) for r in range(world)]
    assert boxes[0].grid == (2, 2, 2) and slabs[0].grid == (8, 1, 1)
    owned = torch.cat([b.owned_idx for b in boxes])
    assert owned.numel() == n and owned.unique().numel() == n
    counts = [b.n_owned for b in boxes]
    assert max(counts) - min(counts) <= n // 20  # the eight octants are images of one another
    halo = lambda shards: sum(0 if s is None else int(s.numel()) for sh in shards for s in sh.send_idx)
    assert halo(boxes) < 0.75 * halo(slabs)
    nbrs = {}
    for i, j in pairs:
        nbrs.setdefault(int(i), set()).add(int(j))
    for r, b in enumerate(boxes):
        have = set(b.owned_idx.tolist())
        for q, peer in enumerate(boxes):
            if q != r and peer.send_idx[r] is not None:
                have |= set(peer.owned_idx[peer.send_idx[r]].tolist())
        for i in b.owned_idx.tolist():
            assert nbrs.get(i, set()) <= have
