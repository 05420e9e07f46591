#!/usr/bin/env python
"""Training-step throughput (frames/s, atoms/s) of BPNN._train_loop on a synthetic bulk-Li-like set
(BASELINE.json configs[2]: bcc Li, 27 atoms per frame, 12 G2 + 2 G4, rc = 3.54 A, MLP 14-64-64-1, batch 500).

    python tools/bench_train.py [--frames 20000] [--steps 40]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py   (frames sharded)

Frames are sharded across ranks (nnmd_b200.dist.shard_frames), descriptors are computed per rank by the CUDA path,
the only collective of a step is the flat gradient all-reduce.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402


def li_frames(n_frames, seed=0):
    a = 3.51
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    grid = np.stack(np.meshgrid(*[np.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    lat = ((grid + base[None]).reshape(-1, 3) * a)[:27] + 0.25 * a  # inside the cell: wrap-only PBC must not tear atoms off
    gen = torch.Generator().manual_seed(seed)
    pos = torch.from_numpy(lat)[None] + 0.05 * torch.randn(n_frames, 27, 3, generator=gen, dtype=torch.float64)
    return pos.float(), torch.eye(3) * 3 * a, torch.ones(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    from nnmd_b200 import dist
    from nnmd_b200.features import calculate_sf
    from nnmd_b200.nn import BPNN, TrainAtomicDataset
    from nnmd_b200.nn.bpnn import batch_loader
    rc = 3.54
    sfd = {"features": [2] * 12 + [4] * 2,
           "params": [[rc, 4.0, float(rs)] for rs in np.linspace(0.0, 3.0, 12)] + [[rc, 0.1, 1, 1], [rc, 0.1, 1, -1]]}
    pos, cell, pbc = li_frames(args.frames)
    mine = dist.shard_frames(args.frames, rank, world)
    pos = pos[mine.start:mine.stop].to(dev)
    t0 = time.perf_counter()
    g, dg = calculate_sf(pos, cell.to(dev), pbc.to(dev), sfd)
    torch.cuda.synchronize()
    t_desc = time.perf_counter() - t0
    torch.manual_seed(1)
    E = torch.rand(pos.shape[0], 1, device=dev)
    F = torch.randn(pos.shape[0], 27, 3, device=dev) * 0.05
    ds = TrainAtomicDataset({"Li": (g.requires_grad_(True), dg)}, E, F, {"Li": sfd})
    torch.manual_seed(0)
    net = BPNN(dtype=torch.float32)
    net.config(dict(atom_species=["Li"], input_sizes=[14], hidden_sizes=[[64, 64]], learning_rate=0.01, l2_regularization=1e-4,
                    e_loss_coeff=1.0, f_loss_coeff=1.0, load_models=False, path="", optimizer="adamw",
                    batch_size=args.batch, epochs=1))
    if world > 1:  # identical weights everywhere
        for p in net.atomic_nets[0].parameters():
            td.broadcast(p.data, 0)
    per_rank_batch = args.batch  # weak scaling: every GPU keeps the reference's batch of 500 frames
    loader = batch_loader(ds, per_rank_batch, shuffle=True)

    def run(n_steps):
        done = 0
        while done < n_steps:
            class _Cap:
                def __init__(self, it, n): self.it, self.n = it, n
                def __len__(self): return self.n
                def __iter__(self):
                    for k, b in enumerate(self.it):
                        if k >= self.n:
                            break
                        yield b
            n = min(len(loader), n_steps - done)
            net._train_loop(0, _Cap(loader, n))
            done += n
    run(args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.steps)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    if rank == 0:
        step_ms = float(ms.item()) / args.steps
        frames_per_step = per_rank_batch * world
        print(json.dumps({"metric": "training step throughput (synthetic Li bcc, 27 atoms, 14 descriptors, MLP 14-64-64-1)",
                          "n_gpus": world, "global_batch_frames": frames_per_step, "ms_per_step": step_ms,
                          "frames_per_s": frames_per_step / (step_ms * 1e-3), "atoms_per_s": frames_per_step * 27 / (step_ms * 1e-3),
                          "descriptor_build_s_per_rank": t_desc, "frames_per_rank": pos.shape[0], "scaling": "weak (batch per GPU fixed)"}))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
