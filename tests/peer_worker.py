"""Worker of tests/test_gpu_peer.py: run under torchrun with >= 2 GPUs; every rank checks the peer-window collectives
(csrc/peer.cu) against torch.distributed (NCCL) on the same inputs and exits non-zero on a mismatch."""
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    from nnmd_b200 import dist

    # ---- 1. weighted all-reduce: eager calls with changing data (both parities, many steps), then graph replays
    n = 7001
    win = dist.PeerWindow.create(4 * (n + 64), dev)
    assert win is not None, "peer windows unavailable on this box"
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    def reference(buf):
        w = buf[-1].clone()
        x = buf.clone()
        x[:-1] *= w
        td.all_reduce(x)
        out = x.clone()
        out[:-1] /= x[-1]
        return out

    for it in range(40):
        buf = torch.randn(n, device=dev, generator=gen)
        buf[-1] = float((rank + it) % 3)  # ragged weights, some ranks with weight 0
        if it % 7 == 3:
            buf[-1] = 1.0 + rank
        want = reference(buf) if float(buf[-1]) >= 0 else None
        tot = want[-1].item()
        got = buf.clone()
        win.allreduce_weighted(got)
        torch.cuda.synchronize()
        if tot > 0:
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (it, float((got - want).abs().max()))
        # identical on every rank, bit for bit
        gathered = [torch.empty_like(got) for _ in range(world)]
        td.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], g) for g in gathered), "ranks disagree"
    static = torch.zeros(n, device=dev)
    src = torch.randn(n, device=dev, generator=gen)
    src[-1] = 2.0 + rank
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        static.copy_(src)
        win.allreduce_weighted(static)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        static.copy_(src)
        win.allreduce_weighted(static)
    want = reference(src)
    for _ in range(25):
        g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(static, want, rtol=1e-5, atol=1e-6)
    # plain sum
    e = torch.full((3,), float(rank + 1), device=dev)
    dist.allreduce_sum(e)
    torch.cuda.synchronize()
    assert torch.equal(e, torch.full((3,), float(world * (world + 1) // 2), device=dev))
    win.check()

    # ---- 2. halo exchange of the slab decomposition: window route == send/recv route, over several steps with moving atoms
    from nnmd_b200.workloads import fcc_cell
    pos, cell, pbc = fcc_cell(8, dtype=torch.float32)
    a = dist.SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=rank, world=world).to(dev)
    b = dist.SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=rank, world=world).to(dev)
    os.environ["NNMD_NO_PEER"] = "1"
    lb = b.exchange_halo(dev).clone()
    del os.environ["NNMD_NO_PEER"]
    la = a.exchange_halo(dev)
    assert a._window is not None and b._window is None
    torch.cuda.synchronize()
    assert la.shape == lb.shape and torch.equal(la, lb)
    for step in range(6):
        moved = a.owned_pos + 0.01 * (step + 1)
        a.set_owned_positions(moved)
        b.set_owned_positions(moved)
        la, lb = a.exchange_halo(dev), b.exchange_halo(dev)
        torch.cuda.synchronize()
        assert torch.equal(la, lb), step
    a._window.check()

    # ---- 3. data-parallel training: one-graph step over the window == NCCL route (losses of three epochs)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_train
    losses = {}
    for route in ("peer", "nccl"):
        if route == "nccl":
            os.environ["NNMD_NO_PEER"] = "1"
        torch.manual_seed(0)
        ds, meta = bench_train.build_dataset("li", 1200, dev, rank, world)
        net = bench_train.make_bpnn(meta, 100, world)
        from nnmd_b200.nn.bpnn import batch_loader
        loader = batch_loader(ds, 100, shuffle=False)
        losses[route] = [float(net._train_loop(ep, loader)[2]) for ep in range(3)]
        if route == "peer":
            assert net._flat_grads.window is not None
            assert all(len(p.graphs) == 1 for p in net._plans.values() if p.graphs is not None)
            net._flat_grads.window.check()
        os.environ.pop("NNMD_NO_PEER", None)
    for x, y in zip(losses["peer"], losses["nccl"]):
        assert abs(x - y) <= 1e-4 * abs(y) + 1e-7, losses
    if rank == 0:
        print("peer_worker: ok", losses, flush=True)
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
