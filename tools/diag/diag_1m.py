import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nnmd_b200.features.calculate_sf import get_plan, raw_descriptors
from nnmd_b200.neighbor import build_neighbor_list
from nnmd_b200.workloads import bench_table, fcc_cell
from oracle import nl_oracle as O3
DEV = "cuda"
sfd = bench_table()
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 63
lo = float(sys.argv[2]) if len(sys.argv) > 2 else 96.0
pos, cell, pbc = fcc_cell(ncell, dtype=torch.float32)
nl = build_neighbor_list(pos[None].to(DEV), cell.to(DEV), pbc.to(DEV), 6.0)
plan = get_plan(sfd, torch.float32, torch.device(DEV))
for edge in (True, False):
    G, dG, flag = raw_descriptors(plan, nl, edge=edge)
    p = nl.posw[:, :3].double().cpu()
    size, rc = 24.0, 6.0
    sel = ((p >= lo - rc) & (p < lo + size + rc)).all(dim=1)
    idx = torch.nonzero(sel).flatten()
    sub = p[idx]
    inner = ((sub >= lo) & (sub < lo + size)).all(dim=1)
    z3 = torch.zeros(3, 3, dtype=torch.float64)
    G3, dG3, cnt3 = O3.raw_descriptors(sub[None], z3, torch.zeros(3, dtype=torch.float64), sfd)
    ii = idx[inner].to(DEV)
    a = G[ii].cpu().double(); b = G3[0][inner]
    err = (a - b).abs().max(dim=0).values / b.abs().max(dim=0).values
    print("edge", edge, "n inner", int(inner.sum()), "G col err: radial max", float(err[:32].max()), "angular max", float(err[32:].max()))
    print(" per col", [f"{float(v):.1e}" for v in err])
    da = dG[ii].cpu().double(); db = dG3[0][inner]
    derr = (da - db).abs().amax(dim=(0, 2)) / db.abs().amax(dim=(0, 2))
    print(" dG col err: radial max", float(derr[:32].max()), "angular max", float(derr[32:].max()))
    worst = (a - b).abs().max(dim=1).values.argmax()
    print(" worst atom", int(worst), "pos", sub[inner][worst].tolist(), "count", int(cnt3[0][inner][worst]))
