"""Latency of the gradient all-reduce: peer-window kernel (eager and graph-replayed) against NCCL.  Run under torchrun."""
import os
import sys
import torch
import torch.distributed as td
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize(); td.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    td.init_process_group("nccl", device_id=dev)
    from nnmd_b200 import dist
    out = {}
    for n in (6338, 30000):
        win = dist.PeerWindow.create(4 * (n + 64), dev)
        buf = torch.randn(n, device=dev); buf[-1] = 1.0
        out[f"peer_eager_us_{n}"] = timed(lambda: win.allreduce_weighted(buf))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                win.allreduce_weighted(buf)
        out[f"peer_graph10_us_per_call_{n}"] = timed(g.replay, 300) / 10
        out[f"nccl_us_{n}"] = timed(lambda: td.all_reduce(buf))
    if rank == 0:
        import json
        print(json.dumps({"world": world, **out}))
    td.destroy_process_group()


if __name__ == "__main__":
    main()
