#!/usr/bin/env python
"""Training-step throughput (frames/s, atoms/s) of BPNN._train_loop on the two training workloads of BASELINE.json:

  li   configs[2]: synthetic bulk-Li-like set, bcc Li, 27 atoms per frame, the Li sample's own 12 G2 + 2 G4 set (rc = 3.54 A),
       MLP 14-64-64-1, AdamW, batch 500 (samples/li/input/input.yaml)
  bin  configs[4]: synthetic binary set, 32 A + 32 B atoms per frame, per species 16 G2 + 16 G5 (rc = 6 A), MLPs 32-64-64-1

    python tools/bench_train.py [--config li|bin] [--frames 20000] [--steps 40] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py   (frames sharded)

Targets are analytic Lennard-Jones energies and forces, scaled like nnmd/nn/dataset.py:85-87.  Frames are sharded across
ranks (nnmd_b200.dist.shard_frames), descriptors are computed per rank by the CUDA path, the only collective of a step is
the flat gradient all-reduce.  weak: every GPU keeps the reference's batch of 500 frames (global batch 500 N); strong: the
global batch stays 500 frames (500 / N per GPU) -- the optimisation the reference runs.  `train_bench` is what bench.py
calls for its `train` block; run as a script it prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

LABELS = {"li": "synthetic Li bcc, 27 atoms, 12 G2 + 2 G4 (rc 3.54), MLP 14-64-64-1, AdamW, LJ targets",
          "bin": "synthetic binary bcc, 32 A + 32 B atoms, per species 16 G2 + 16 G5 (rc 6), MLPs 32-64-64-1, AdamW, LJ targets"}


def build_dataset(config: str, frames: int, dev, rank: int, world: int):
    """This rank's shard of the synthetic training set: descriptors from the CUDA path, LJ targets with GLOBAL scaling."""
    from nnmd_b200 import dist
    from nnmd_b200.features import calculate_sf
    from nnmd_b200.nn import TrainAtomicDataset
    from nnmd_b200.workloads import binary_frames, binary_table, li_frames, li_table, lj_targets
    mine = dist.shard_frames(frames, rank, world)
    t0 = time.perf_counter()
    if config == "li":
        sfd = li_table()
        pos, cell, pbc = li_frames(frames)
        E, F = lj_targets(pos.to(dev), sigma=2.7)
        g, dg = calculate_sf(pos[mine.start:mine.stop].to(dev), cell.to(dev), pbc.to(dev), sfd)
        sf_data, params = {"Li": (g.requires_grad_(True), dg)}, {"Li": sfd}
        meta = dict(species=["Li"], n_atoms=27, n_desc=[14], hidden=[[64, 64]])
    else:
        sfd = binary_table()
        pa, pb, cell, pbc = binary_frames(frames)
        E, F = lj_targets(torch.cat([pa, pb], dim=1).to(dev), sigma=2.45)
        sf_data, params = {}, {}
        for name, p in (("A", pa), ("B", pb)):
            g, dg = calculate_sf(p[mine.start:mine.stop].to(dev), cell.to(dev), pbc.to(dev), sfd)
            sf_data[name], params[name] = (g.requires_grad_(True), dg), sfd
        meta = dict(species=["A", "B"], n_atoms=64, n_desc=[32, 32], hidden=[[64, 64], [64, 64]])
    torch.cuda.synchronize()
    meta["descriptor_build_s"] = time.perf_counter() - t0
    ds = TrainAtomicDataset(sf_data, E[mine.start:mine.stop].contiguous(), F[mine.start:mine.stop].contiguous(), params)
    return ds, meta


def make_bpnn(meta, batch: int, world: int):
    from nnmd_b200.nn import BPNN
    torch.manual_seed(0)
    net = BPNN(dtype=torch.float32)
    net.config(dict(atom_species=meta["species"], input_sizes=meta["n_desc"], hidden_sizes=meta["hidden"], learning_rate=0.01,
                    l2_regularization=1e-4, e_loss_coeff=1.0, f_loss_coeff=0.01, load_models=False, path="", optimizer="adamw",
                    batch_size=batch, epochs=1))
    if world > 1:  # identical weights everywhere
        for an in net.atomic_nets:
            for p in an.parameters():
                td.broadcast(p.data, 0)
    return net


def train_bench(config: str, dev, rank: int, world: int, frames: int = 20000, batch: int = 500, steps: int = 40,
                warmup: int = 5, scaling: str = "weak", dataset=None):
    """Time `steps` optimisation steps of BPNN._train_loop (device events, max over ranks).  Returns a dict (all ranks)."""
    from nnmd_b200.nn.bpnn import FrameBatches, batch_loader
    ds, meta = dataset if dataset is not None else build_dataset(config, frames, dev, rank, world)
    per_rank = batch if scaling == "weak" else max(batch // world, 1)
    net = make_bpnn(meta, per_rank, world)
    loader = batch_loader(ds, per_rank, shuffle=True)
    n_local = len(ds)

    def run(n_steps):
        done, last = 0, None
        while done < n_steps:
            n = min(len(loader), n_steps - done)
            part = loader if n == len(loader) else FrameBatches(ds, torch.arange(n * per_rank), per_rank, True)
            last = net._train_loop(0, part)
            done += n
        return last

    first = run(max(warmup, 3) + len(loader))  # at least one whole epoch: every graph of the step (single, chunk of 8) is captured before the clock starts
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    last = run(steps)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    step_ms = float(ms.item()) / steps
    fg = getattr(net, "_flat_grads", None)
    if fg is not None and fg.window is not None:
        fg.window.check()  # raises if a peer ever failed to arrive at an all-reduce (the timing would be of garbage)
    global_batch = per_rank * world
    return {"config": LABELS[config], "n_gpus": world, "scaling": scaling, "global_batch_frames": global_batch,
            "ms_per_step": step_ms, "frames_per_s": global_batch / (step_ms * 1e-3),
            "atoms_per_s": global_batch * meta["n_atoms"] / (step_ms * 1e-3), "steps": steps,
            "descriptor_build_s_per_rank": meta["descriptor_build_s"], "frames_per_rank": n_local,
            "loss_first_epoch": float(first[2]), "loss_last_epoch": float(last[2])}


def reference_train_baseline(config: str, ds, meta, batch: int = 500, max_frames: int = 4000):
    """The reference's own `_train_loop` (nnmd/nn/bpnn.py:308-377: MLP forward, autograd dE/dg with create_graph, einsum,
    loss, double backward, AdamW) on the host cores, on the first `max_frames` frames of the SAME descriptors and targets.
    Uses the unmodified reference staged in oracle/_ref when present (kind "reference"), else a plain-torch restatement of
    the same step (kind "port").  Returns the cpu_baseline dict of the train block."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = min(len(ds), max_frames) // batch * batch
    if n == 0:
        n = min(len(ds), batch)
    sf = {s: (g.detach()[:n].cpu().clone().requires_grad_(True), dg[:n].cpu()) for s, (g, dg) in ds.sf_data.items()}
    E, F = ds.energies[:n].cpu(), ds.forces[:n].cpu()
    cfg = dict(atom_species=meta["species"], input_sizes=meta["n_desc"], hidden_sizes=meta["hidden"], learning_rate=0.01,
               l2_regularization=1e-4, e_loss_coeff=1.0, f_loss_coeff=0.01, load_models=False, path="", optimizer="adamw",
               batch_size=batch, epochs=1)
    kind = "port"
    if len(meta["species"]) == 1:  # the reference's training step raises for two species
        try:
            from oracle.build_ref import available, load_reference
            if available():
                load_reference()
                from nnmd.nn import BPNN as RefBPNN, TrainAtomicDataset as RefDataset
                net = RefBPNN(dtype=torch.float32, use_cuda=False)
                net.config(cfg)
                loader = torch.utils.data.DataLoader(RefDataset(sf, E, F, ds.symm_func_params), batch_size=batch, shuffle=True)
                net._train_loop(0, torch.utils.data.DataLoader(RefDataset(sf, E, F, ds.symm_func_params), batch_size=batch))  # warm-up epoch
                t0 = time.perf_counter()
                net._train_loop(0, loader)
                dt = time.perf_counter() - t0
                kind = "reference"
        except Exception as exc:  # fall through to the restatement, loudly
            print(f"bench_train: reference _train_loop unavailable ({exc!r}); timing the torch restatement", file=sys.stderr)
    if kind == "port":
        nets = [torch.nn.Sequential(*[m for a, b in zip([f] + h, h + [1]) for m in (torch.nn.Linear(a, b), torch.nn.Sigmoid())][:-1])
                for f, h in zip(meta["n_desc"], meta["hidden"])]
        opt = torch.optim.AdamW([p for m in nets for p in m.parameters()], lr=0.01, weight_decay=1e-4, amsgrad=True)

        def epoch():
            for lo in range(0, n, batch):
                opt.zero_grad(set_to_none=True)
                es, fs = [], []
                for m, s in zip(nets, meta["species"]):
                    g = sf[s][0][lo:lo + batch]
                    e = m(g)
                    dE = torch.autograd.grad(e.sum(), g, create_graph=True)[0]
                    es.append(e)
                    fs.append(-torch.einsum("ijk,ijkl->ijl", dE, sf[s][1][lo:lo + batch]))
                e_all, f_all = torch.cat(es, dim=1), torch.cat(fs, dim=1)
                loss = torch.mean((e_all.sum(dim=1) - E[lo:lo + batch]) ** 2) + 0.01 / 3 * torch.mean((f_all - F[lo:lo + batch]) ** 2)
                loss.backward()
                opt.step()
        epoch()
        t0 = time.perf_counter()
        epoch()
        dt = time.perf_counter() - t0
    n_steps = max(n // batch, 1)
    return {"value": n / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind, "ms_per_step": dt / n_steps * 1e3,
            "sample": f"{n_steps} steps of batch {batch} on the first {n} frames of the same descriptors and targets, fp32, "
                      + ("unmodified reference BPNN._train_loop from oracle/_ref" if kind == "reference"
                         else "plain-torch restatement of bpnn.py:342-366 (the reference's step raises for two species, SURVEY Q9)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="li", choices=["li", "bin"])
    ap.add_argument("--frames", type=int, default=20000)
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    dataset = build_dataset(args.config, args.frames, dev, rank, world)
    out = train_bench(args.config, dev, rank, world, args.frames, args.batch, args.steps, args.warmup, args.scaling, dataset)
    if args.cpu_baseline and rank == 0 and world == 1:
        out["cpu_baseline"] = reference_train_baseline(args.config, *dataset, batch=args.batch)
    if rank == 0:
        out["metric"] = "training step throughput"
        print(json.dumps(out))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
