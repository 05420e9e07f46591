import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from helpers import col_relerr
from nnmd_b200 import _lib
from nnmd_b200.features.calculate_sf import get_plan, raw_descriptors
from nnmd_b200.neighbor import build_neighbor_list
from nnmd_b200.workloads import bench_table, fcc_cell
DEV = "cuda"
for ncell, rc in ((7, 6.0), (5, 12.0), (5, 9.0)):
    for kind in (4, 5):
        sfd = bench_table(rc=rc)
        sfd = dict(sfd, features=[f if f == 2 else kind for f in sfd["features"]])
        pos, cell, pbc = fcc_cell(ncell, dtype=torch.float32)
        nl64 = build_neighbor_list(pos[None].double().to(DEV), cell.double().to(DEV), pbc.double().to(DEV), rc)
        p64 = get_plan(sfd, torch.float64, torch.device(DEV))
        G64, dG64, _ = raw_descriptors(p64, nl64, with_grad=True)
        plan = get_plan(sfd, torch.float32, torch.device(DEV))
        nl = build_neighbor_list(pos[None].to(DEV), cell.to(DEV), pbc.to(DEV), rc)
        for form in (0, 1):
            _lib.set_option("edge_form", form)
            for wg in (True, False):
                G, dG, flag = raw_descriptors(plan, nl, with_grad=wg)
                torch.cuda.synchronize()
                eg = col_relerr(G, G64, 1)
                ed = col_relerr(dG, dG64, 1) if wg else -1
                print(f"cells {ncell} rc {rc} kind {kind} form {form} grad {wg}: max_row {nl.max_row} G err {eg:.2e} dG err {ed:.2e}")
