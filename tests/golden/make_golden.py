"""Generate the golden fixtures in this directory by EXECUTING THE UNMODIFIED REFERENCE
(/root/reference, ded-otshelnik/nnmd) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own for this path, so these
files are what pins the oracle (oracle/) and, through it, the CUDA path.  Every fixture
stores its inputs next to the reference's outputs; tests never import the reference.

torch build used: see the `torch_version` field stored in every file.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

nnmd = import_reference()
from nnmd.features import calculate_sf, calculate_params  # noqa: E402
from nnmd.features.symm_funcs import calculate_distances, _pairwise_filter  # noqa: E402
from nnmd.nn import BPNN, AtomicNN, TrainAtomicDataset  # noqa: E402
from nnmd.io import gpaw_parser  # noqa: E402
from nnmd.io.input_parser import _symmetry_functions_parser  # noqa: E402

META = {"torch_version": torch.__version__, "reference": "ded-otshelnik/nnmd 0.2.0-dev (/root/reference)"}


def ref_sf(cart, cell, pbc, sfd, dtype):
    c = torch.tensor(cart, dtype=dtype)
    g, dg = calculate_sf(c, torch.tensor(cell, dtype=dtype), torch.tensor(pbc, dtype=dtype), sfd, disable_tqdm=True)
    torch.autograd.set_detect_anomaly(False)  # the reference switches it on globally (calculate_sf.py:42)
    return g.numpy(), dg.numpy()


def ref_pairs(cart, cell, pbc, rc, dtype):
    d = calculate_distances(torch.tensor(cart, dtype=dtype), torch.tensor(cell, dtype=dtype),
                            torch.tensor(pbc, dtype=dtype))
    m = _pairwise_filter(d, rc)
    dd = d[(d > 0)]
    margin = float((dd - rc).abs().min())
    return torch.nonzero(m).numpy().astype(np.int32), margin


def pack(sfd):
    feats = np.asarray(sfd["features"], dtype=np.int32)
    prm = np.full((len(feats), 4), np.nan)
    is_int = np.zeros((len(feats), 4), dtype=bool)
    for i, p in enumerate(sfd["params"]):
        prm[i, :len(p)] = [float(x) for x in p]
        is_int[i, :len(p)] = [isinstance(x, (int, np.integer)) for x in p]
    return feats, prm, is_int


def save_sf_case(name, cart, cell, pbc, sfd, rc_pairs):
    feats, prm, is_int = pack(sfd)
    out = dict(cart=cart, cell=cell, pbc=pbc, features=feats, params=prm, params_is_int=is_int, rc_pairs=rc_pairs)
    for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
        g, dg = ref_sf(cart, cell, pbc, sfd, dt)
        out[f"g_{tag}"] = g
        out[f"dg_{tag}"] = dg
        pr, margin = ref_pairs(cart, cell, pbc, rc_pairs, dt)
        out[f"pairs_{tag}"] = pr
        out[f"pairs_margin_{tag}"] = margin
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out, **META)
    print(name, "g", out["g_f64"].shape, "pairs", out["pairs_f64"].shape, "margin", out["pairs_margin_f64"],
          out["pairs_margin_f32"])


def fcc(ncell, a=3.615, sigma=0.05, seed=0):
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    grid = np.stack(np.meshgrid(*[np.arange(ncell)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    pos = (grid + base[None]).reshape(-1, 3) * a
    g = torch.Generator().manual_seed(seed)
    pos = pos + sigma * torch.randn(pos.shape, generator=g, dtype=torch.float64).numpy()
    return pos, np.eye(3) * ncell * a


def main():
    torch.manual_seed(1)
    # ---- case 1: random cluster in a triclinic cell, pbc T,T,F, mixed G1/G2/G4/G5 ----------------
    cart = (torch.rand(3, 40, 3, dtype=torch.float64) * torch.tensor([12.0, 10, 5]) - 3).numpy()
    cell = np.array([[6.0, 0, 0], [1, 5, 0], [0, 0.5, 7]])
    pbc = np.array([1.0, 1, 0])
    sfd = {"features": [4, 4, 4, 5, 5, 2, 1, 2, 5, 4],
           "params": [[3.0, 0.1, 1, -1], [3.0, 0.1, 1.0, 1], [2.5, 0.1, 2.5, -1], [3.0, 0.1, 2.5, -1],
                      [3.0, 0.05, 4, 1], [3.0, 0.5, 1.0], [2.5], [3.0, 1.5, 0.0], [2.8, 0.3, 1.7, 1],
                      [3.0, 0.1, 16.0, 1]]}
    save_sf_case("sf_cluster", cart, cell, pbc, sfd, 3.0)

    # ---- case 2: Cu(111) slab frames from the reference's GPAW log, its own JSON parameter set ----
    n_atoms, frames, cell, pbc = gpaw_parser("/root/reference/samples/gpaw/Cu111.txt")
    cart = np.stack([fr["Cu"]["positions"] for fr in frames[:3]])
    import json
    with open("/root/reference/samples/cu111/input/symm_cu111.json") as fh:
        sfd = _symmetry_functions_parser(json.load(fh))["Cu"]
    save_sf_case("sf_cu111_json", cart, np.asarray(cell), np.asarray(pbc, dtype=float), sfd, 12.0)
    # the same frames with the auto-parameter scheme used by samples/cu111/train.py:19-26 (Q10 trap included)
    prm = calculate_params(8.0, 1, 6, 0, 6, 6)
    sfd = {"features": [1] * 1 + [2] * 6 + [4] * 6 + [5] * 6, "params": prm}
    save_sf_case("sf_cu111_auto", cart[:2], np.asarray(cell), np.asarray(pbc, dtype=float), sfd, 8.0)

    # ---- case 3: jittered Cu fcc 3x3x3 (N=108), rc=6, slice of the bench table ------------------
    pos, cell = fcc(3)
    ang = calculate_params(6.0, 0, 32, 0, 64, 0)[32:]
    rad = [[6.0, 4.0, float(rs)] for rs in np.linspace(0.5, 5.5, 32)]
    sel_r = [0, 9, 13, 20, 31]
    sel_a = [0, 1, 2, 3, 20, 21, 40, 41, 62, 63]
    sfd = {"features": [2] * len(sel_r) + [4] * len(sel_a),
           "params": [rad[i] for i in sel_r] + [ang[i] for i in sel_a]}
    save_sf_case("sf_fcc108", pos[None], cell, np.ones(3), sfd, 6.0)

    # ---- case 4: LJ trimer-like frames, 3 x G1 (samples/lennard-jones/input/symm_lj.yaml) --------
    g = torch.Generator().manual_seed(7)
    tri = np.array([[0, 0, 0], [1.12, 0, 0], [0.56, 0.97, 0]])
    cart = tri[None] + 0.15 * torch.randn(16, 3, 3, generator=g, dtype=torch.float64).numpy()
    sfd = _symmetry_functions_parser({"H": {"G1": {"r_cutoff": [2.0, 2.5, 3.0]}}})["H"]
    save_sf_case("sf_lj_trimer", cart, np.zeros((3, 3)), np.zeros(3), sfd, 3.0)

    # ---- calculate_params vectors (auto_params.py:4-59) ------------------------------------------
    cp = {}
    for args in [(6.0, 0, 32, 0, 64, 0), (38.0, 0, 25, 0, 0, 25), (8.0, 1, 6, 0, 6, 6), (3.54, 2, 4, 1, 3, 5)]:
        rows = calculate_params(*args)
        arr = np.full((len(rows), 4), np.nan)
        for i, r in enumerate(rows):
            arr[i, :len(r)] = [float(x) for x in r]
        cp["args_" + "_".join(str(a) for a in args)] = arr
    np.savez_compressed(os.path.join(HERE, "calculate_params.npz"), **cp, **META)

    # ---- energy / force / train step (atomic_nn.py, bpnn.py:196-215, 308-377) --------------------
    d = np.load(os.path.join(HERE, "sf_cluster.npz"))
    gq = torch.tensor(d["g_f32"])
    dgq = torch.tensor(d["dg_f32"])
    B, N, F = gq.shape
    torch.manual_seed(3)
    e_ref = torch.rand(B, 1)
    f_ref = torch.randn(B, N, 3) * 0.1
    out = dict(g=gq.numpy(), dg=dgq.numpy(), e_ref=e_ref.numpy(), f_ref=f_ref.numpy())
    for opt in ("adam", "adamw"):
        cfg = dict(atom_species=["X"], input_sizes=[F], hidden_sizes=[[16, 8]], learning_rate=0.01,
                   l2_regularization=1e-4, e_loss_coeff=1.0, f_loss_coeff=0.1, load_models=False,
                   path="", optimizer=opt, batch_size=B, epochs=3)
        torch.manual_seed(11)
        net = BPNN(dtype=torch.float32, use_cuda=False)
        net.config(cfg)
        sd0 = {k: v.clone() for k, v in net.atomic_nets[0].state_dict().items()}
        gl = gq.clone().requires_grad_(True)
        ds = TrainAtomicDataset({"X": (gl, dgq)}, e_ref, f_ref, {})
        loader = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False)
        # forward quantities exactly as bpnn.py:342-354 computes them
        e_at = net.atomic_nets[0](gl)
        dE = torch.autograd.grad(e_at.sum(), gl, create_graph=True)[0]
        f_nn = -torch.einsum("ijk,ijkl->ijl", dE, dgq)
        if opt == "adam":
            for k, v in sd0.items():
                out["w0_" + k] = v.numpy()
            out["e_atoms"] = e_at.detach().numpy()
            out["dE"] = dE.detach().numpy()
            out["f_nn"] = f_nn.detach().numpy()
        losses = []
        for ep in range(3):
            e_l, f_l, l = net._train_loop(ep, loader)
            losses.append([e_l.item(), f_l.item(), l.item()])
        out[f"losses_{opt}"] = np.asarray(losses)
        for k, v in net.atomic_nets[0].state_dict().items():
            out[f"w3_{opt}_" + k] = v.detach().numpy()
        e_l, f_l, l = net._validate(0, loader)
        out[f"val_{opt}"] = np.asarray([e_l.item(), f_l.item(), l.item()])
    out["cfg_hidden"] = np.asarray([16, 8])
    out["cfg_scalars"] = np.asarray([0.01, 1e-4, 1.0, 0.1])  # lr, l2, e_coeff, f_coeff
    np.savez_compressed(os.path.join(HERE, "train_step.npz"), **out, **META)
    print("train_step losses", out["losses_adam"], out["losses_adamw"])


if __name__ == "__main__":
    main()
