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


def binary_frames(n_frames, seed=0):
    """BASELINE.json configs[4]: bcc-ordered binary cell, 32 A + 32 B atoms (4x4x2 cells, A on corners, B on centres)."""
    a = 3.2
    grid = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    A = grid * a + 0.25 * a
    B = (grid + 0.5) * a + 0.25 * a
    gen = torch.Generator().manual_seed(seed)
    pa = torch.from_numpy(A)[None] + 0.05 * torch.randn(n_frames, 32, 3, generator=gen, dtype=torch.float64)
    pb = torch.from_numpy(B)[None] + 0.05 * torch.randn(n_frames, 32, 3, generator=gen, dtype=torch.float64)
    cell = torch.diag(torch.tensor([4 * a, 4 * a, 2 * a]))
    return pa.float(), pb.float(), cell.float(), torch.ones(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="li", choices=["li", "bin"])
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
    from nnmd_b200.nn.bpnn import FrameBatches, batch_loader
    from nnmd_b200.features import calculate_params
    mine = dist.shard_frames(args.frames, rank, world)
    t0 = time.perf_counter()
    if args.config == "li":
        rc = 3.54
        sfd = {"features": [2] * 12 + [4] * 2,
               "params": [[rc, 4.0, float(rs)] for rs in np.linspace(0.0, 3.0, 12)] + [[rc, 0.1, 1, 1], [rc, 0.1, 1, -1]]}
        pos, cell, pbc = li_frames(args.frames)
        pos = pos[mine.start:mine.stop].to(dev)
        g, dg = calculate_sf(pos, cell.to(dev), pbc.to(dev), sfd)
        sf_data, params = {"Li": (g.requires_grad_(True), dg)}, {"Li": sfd}
        species, n_atoms, n_desc, hidden = ["Li"], 27, [14], [[64, 64]]
        label = "synthetic Li bcc, 27 atoms, 12 G2 + 2 G4, MLP 14-64-64-1"
    else:
        rc = 6.0
        ang = calculate_params(rc, 0, 2, 0, 0, 16)[2:]
        sfd = {"features": [2] * 16 + [5] * 16,
               "params": [[rc, 2.0, float(rs)] for rs in np.linspace(0.5, 5.5, 16)] + [[float(v) for v in r] for r in ang]}
        pa, pb, cell, pbc = binary_frames(args.frames)
        sf_data, params = {}, {}
        for name, p in (("A", pa), ("B", pb)):
            p = p[mine.start:mine.stop].to(dev)
            g, dg = calculate_sf(p, cell.to(dev), pbc.to(dev), sfd)
            sf_data[name], params[name] = (g.requires_grad_(True), dg), sfd
        pos = pa[mine.start:mine.stop]
        species, n_atoms, n_desc, hidden = ["A", "B"], 64, [32, 32], [[64, 64], [64, 64]]
        label = "synthetic binary bcc, 32 A + 32 B atoms, per species 16 G2 + 16 G5 (rc = 6), MLPs 32-64-64-1"
    torch.cuda.synchronize()
    t_desc = time.perf_counter() - t0
    n_local = pos.shape[0]
    torch.manual_seed(1)
    E = torch.rand(n_local, 1, device=dev)
    F = torch.randn(n_local, n_atoms, 3, device=dev) * 0.05
    ds = TrainAtomicDataset(sf_data, E, F, params)
    torch.manual_seed(0)
    net = BPNN(dtype=torch.float32)
    net.config(dict(atom_species=species, input_sizes=n_desc, hidden_sizes=hidden, learning_rate=0.01, l2_regularization=1e-4,
                    e_loss_coeff=1.0, f_loss_coeff=1.0, load_models=False, path="", optimizer="adamw",
                    batch_size=args.batch, epochs=1))
    if world > 1:  # identical weights everywhere
        for an in net.atomic_nets:
            for p in an.parameters():
                td.broadcast(p.data, 0)
    per_rank_batch = args.batch  # weak scaling: every GPU keeps the reference's batch of 500 frames
    loader = batch_loader(ds, per_rank_batch, shuffle=True)

    def run(n_steps):
        done = 0
        while done < n_steps:
            n = min(len(loader), n_steps - done)
            part = loader if n == len(loader) else FrameBatches(ds, torch.arange(n * per_rank_batch), per_rank_batch, True)
            net._train_loop(0, part)
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
        print(json.dumps({"metric": "training step throughput (" + label + ")",
                          "n_gpus": world, "global_batch_frames": frames_per_step, "ms_per_step": step_ms,
                          "frames_per_s": frames_per_step / (step_ms * 1e-3), "atoms_per_s": frames_per_step * n_atoms / (step_ms * 1e-3),
                          "descriptor_build_s_per_rank": t_desc, "frames_per_rank": n_local, "scaling": "weak (batch per GPU fixed)"}))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
