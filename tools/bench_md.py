"""MD-loop throughput of nnmd_b200.md.VelocityVerlet on the bench cell: steps/s with the whole step replayed from
one CUDA graph, with and without a Verlet skin.  One GPU.  `python tools/bench_md.py [--cells 63] [--steps 20] [--skin 0.3]`"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=63)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--skin", type=float, default=0.3)
    ap.add_argument("--dt", type=float, default=0.05)      # the random-weight network is no physical potential: small steps keep the lattice
    args = ap.parse_args()
    from nnmd_b200.md import VelocityVerlet
    from nnmd_b200.workloads import bench_net, bench_table, fcc_cell
    dev = torch.device("cuda")
    sfd = bench_table()
    pos, cell, pbc = fcc_cell(args.cells)
    # a quarter lattice constant off the cell faces: under the reference's wrap-only PBC an atom that crosses a face jumps to the
    # other side of the structure (no periodic images), which rightly forces a rebuild; lattice planes ON the faces would do so
    # in every step
    pos = pos + 0.25 * 3.615
    net = bench_net(len(sfd["features"]), dev)
    out = {"atoms": int(pos.shape[0]), "steps": args.steps}
    for skin in (0.0, args.skin):
        torch.manual_seed(0)
        v = 1e-3 * torch.randn_like(pos)
        md = VelocityVerlet(pos.to(dev), v.to(dev), 63.5, cell, pbc, sfd, net, dt=args.dt, skin=skin)
        md.step(4)  # two eager steps, the capture, one replay
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0 = md.list_builds()
        a.record()
        md.step(args.steps)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        out[f"skin_{skin}"] = {"ms_per_step": ms, "atom_steps_per_s": pos.shape[0] / (ms * 1e-3),
                               "list_builds": args.steps if skin == 0.0 else md.list_builds() - b0, "energy": float(md.energy)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
