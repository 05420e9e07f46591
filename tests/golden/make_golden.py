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


def lj_energy_forces(pos, eps=0.05, sigma=2.7):
    """Analytic Lennard-Jones energy and forces of a free cluster (targets for the synthetic training sets: SURVEY 8d)."""
    d = pos[:, None, :] - pos[None, :, :]
    r2 = (d ** 2).sum(-1) + np.eye(len(pos))
    s6 = (sigma ** 2 / r2) ** 3
    np.fill_diagonal(s6, 0.0)
    e = 2.0 * eps * (s6 ** 2 - s6).sum()
    coef = 24.0 * eps * (2.0 * s6 ** 2 - s6) / r2
    return float(e), (coef[:, :, None] * d).sum(axis=1)


def config_with_shipped_weights(net, cfg, spec):
    """net.config(cfg) with load_models=True: the reference calls torch.load without map_location (bpnn.py:139-146), which
    fails for the shipped CUDA-saved weights on a CPU-only box.  Same end state: configure without loading, then put the
    global RNG where a real load_models=True run leaves it (after the nn.Linear default initialisation of the AtomicNN
    constructor, before any Xavier draw) and load the state dict with map_location."""
    state = torch.get_rng_state()
    net.config(dict(cfg, load_models=False))
    torch.set_rng_state(state)
    AtomicNN(input_size=cfg["input_sizes"][0], hidden_sizes=cfg["hidden_sizes"][0], output_size=1)
    net.atomic_nets[0].load_state_dict(torch.load(f"{cfg['path']}/atomic_nn_{spec}.pt", map_location="cpu"))
    net.neural_net_data = dict(cfg)


def reference_fit(net, train, val, test, batch_size, epochs, shuffle):
    """The body of BPNN.fit (bpnn.py:217-306) without its logger / package-metadata dependencies: same loaders, same
    scheduler, same call order (hence the same draws from the global RNG)."""
    from torch.utils.data import DataLoader
    tl = DataLoader(train, batch_size=batch_size, shuffle=shuffle)
    vl = DataLoader(val, batch_size=batch_size, shuffle=shuffle)
    sl = DataLoader(test, batch_size=batch_size, shuffle=shuffle)
    net.sched = torch.optim.lr_scheduler.ExponentialLR(net.optim, gamma=0.99)
    tr, va = [], []
    for ep in range(epochs):
        tr.append([v.item() for v in net._train_loop(ep, tl)])
        va.append([v.item() for v in net._validate(ep, vl)])
        net.sched.step()
    te = [v.item() for v in net._test(sl)]
    return np.asarray(tr), np.asarray(va), np.asarray(te)


def make_train_li():
    """BASELINE.json configs[2] at parity size: the Li sample's OWN parameter set (samples/li/input/symm_li_27.json),
    network shape, optimizer settings (samples/li/input/input.yaml) and shipped weights as the starting point, on synthetic
    27-atom bcc Li frames (the sample's .traj is a missing blob) with analytic LJ targets.  3 epochs of the unmodified
    reference's _train_loop / _validate."""
    import json
    import yaml
    with open("/root/reference/samples/li/input/symm_li_27.json") as fh:
        sfd = _symmetry_functions_parser(json.load(fh))["Li"]
    with open("/root/reference/samples/li/input/input.yaml") as fh:
        cfg = yaml.safe_load(fh)["neural_network"]
    a = 3.51
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    grid = np.stack(np.meshgrid(*[np.arange(3)] * 3, indexing="ij"), -1).reshape(-1, 1, 3)
    lat = ((grid + base[None]).reshape(-1, 3) * a)[:27] + 0.25 * a
    gen = torch.Generator().manual_seed(17)
    M = 24
    cart = lat[None] + 0.08 * torch.randn(M, 27, 3, generator=gen, dtype=torch.float64).numpy()
    cell, pbc = np.eye(3) * 3 * a, np.ones(3)
    ef = [lj_energy_forces(c) for c in cart]
    E = torch.tensor([[e] for e, _ in ef], dtype=torch.float32)
    F = torch.tensor(np.array([f for _, f in ef]), dtype=torch.float32)
    emin, emax = E.min(), E.max()      # dataset.py:85-87
    E = (E - emin) / (emax - emin)
    F = F / (emax - emin)
    g, dg = ref_sf(cart, cell, pbc, sfd, torch.float32)
    g, dg = torch.tensor(g), torch.tensor(dg)
    feats, prm, is_int = pack(sfd)
    out = dict(cart=cart, cell=cell, pbc=pbc, features=feats, params=prm, params_is_int=is_int, g=g.numpy(), dg=dg.numpy(),
               e_ref=E.numpy(), f_ref=F.numpy())
    cfg = dict(cfg, input_sizes=[len(sfd["features"])], path="/root/reference/samples/li/models_adamw", batch_size=8, epochs=3)
    net = BPNN(dtype=torch.float32, use_cuda=False)
    config_with_shipped_weights(net, cfg, "Li")
    for k, v in net.atomic_nets[0].state_dict().items():
        out["w0_" + k] = v.detach().clone().numpy()
    ds = TrainAtomicDataset({"Li": (g.clone().requires_grad_(True), dg)}, E, F, {"Li": sfd})
    tr, va, te = reference_fit(net, ds, ds, ds, 8, 3, shuffle=False)
    out.update(losses_train=tr, losses_val=va, losses_test=te)
    for k, v in net.atomic_nets[0].state_dict().items():
        out["w3_" + k] = v.detach().numpy()
    out["cfg_scalars"] = np.asarray([cfg["learning_rate"], cfg["l2_regularization"], cfg["e_loss_coeff"], cfg["f_loss_coeff"]])
    np.savez_compressed(os.path.join(HERE, "train_li.npz"), **out, **META)
    print("train_li", tr, va, te)


def make_cu111_flow(n_frames=48, epochs=2):
    """samples/cu111/train.py:15-66 executed by the unmodified reference (CPU, float32) on the first `n_frames` frames of
    samples/gpaw/cu111.txt: the parser's output dict is stored as the fixture's INPUT, the dataset / split / training /
    evaluation numbers as its expected OUTPUT.  Deviations forced by the reference's own defects: make_atomic_dataset
    hard-codes "cuda" (dataset.py:48), so its body is executed here line by line on the CPU; fit() needs DEBUG / package
    metadata, so its loop is `reference_fit`; predict() raises TypeError as shipped (bpnn.py:529-531), so
    the energy / force evaluation is bpnn.py:537-544 on calculate_sf's output."""
    import yaml
    n_atoms, frames, cell, pbc = gpaw_parser("/root/reference/samples/gpaw/cu111.txt")
    frames = frames[:n_frames]
    with open("/root/reference/samples/cu111/input/input.yaml") as fh:
        nn_cfg = yaml.safe_load(fh)["neural_network"]
    N1, N2, N3, N4, N5 = 0, 25, 0, 0, 25        # train.py:19-26
    params = calculate_params(38.0, N1, N2, N3, N4, N5)
    sfd = {"params": params, "features": [1] * N1 + [2] * N2 + [3] * N3 + [4] * N4 + [5] * N5}
    nn_cfg["input_sizes"] = [N1 + N2 + N3 + N4 + N5]
    # dataset.py:55-101 on the CPU
    cart = torch.tensor(np.array([fr["Cu"]["positions"] for fr in frames]), dtype=torch.float32)
    E = torch.tensor(np.array([fr["energy"] for fr in frames]), dtype=torch.float32).unsqueeze(1)
    F = torch.tensor(np.array([fr["forces"] for fr in frames]), dtype=torch.float32)
    emin, emax = E.min(), E.max()
    E = (E - emin) / (emax - emin)
    F = F / (emax - emin)
    cell_t, pbc_t = torch.tensor(np.asarray(cell), dtype=torch.float32), torch.tensor(np.asarray(pbc), dtype=torch.float32)
    g, dg = calculate_sf(cart, cell_t, pbc_t, sfd, disable_tqdm=True)
    torch.autograd.set_detect_anomaly(False)
    g.requires_grad = True
    ds = TrainAtomicDataset({"Cu": (g, dg)}, E, F, {"Cu": sfd})
    from nnmd.nn import train_val_test_split
    train, val, test = train_val_test_split(ds, (0.8, 0.1, 0.1))      # train.py:38-41
    net = BPNN(dtype=torch.float32, use_cuda=False)
    config_with_shipped_weights(net, dict(nn_cfg, path="/root/reference/samples/cu111/models_adamw"), "Cu")  # load_models: True
    out = {"w0_" + k: v.detach().clone().numpy() for k, v in net.atomic_nets[0].state_dict().items()}
    # energy + force evaluation with the shipped weights (bpnn.py:537-544), first 3 frames; a working predict() would
    # go through calculate_sf, which reseeds the global RNG (calculate_sf.py:115): reproduce the side effect
    torch.manual_seed(42)
    gp = g[:3].detach().clone().requires_grad_(True)
    e_at = net.atomic_nets[0](gp)
    dE = torch.autograd.grad(e_at.sum(), gp)[0]
    out["pred_e"] = e_at.sum(1).squeeze().detach().numpy()
    out["pred_f"] = (-torch.einsum("ijk,ijkl->ijl", dE, dg[:3])).detach().numpy()
    tr, va, te = reference_fit(net, train, val, test, nn_cfg["batch_size"], epochs, shuffle=True)
    out.update(losses_train=tr, losses_val=va, losses_test=te,
               split_train=np.asarray(train.indices), split_val=np.asarray(val.indices), split_test=np.asarray(test.indices))
    for k, v in net.atomic_nets[0].state_dict().items():
        out["w_end_" + k] = v.detach().numpy()
    # the parser's output dict (input of the flow) and samples of the dataset tensors
    out.update(positions=np.array([fr["Cu"]["positions"] for fr in frames]), forces=np.array([fr["forces"] for fr in frames]),
               energy=np.array([fr["energy"] for fr in frames]), unit_cell=np.asarray(cell), pbc=np.asarray(pbc),
               n_atoms=np.asarray(n_atoms), g_head=g[:2].detach().numpy(), dg_head=dg[:2].numpy(),
               e_scaled=E.numpy(), f_scaled_head=F[:2].numpy(), epochs=np.asarray(epochs),
               nn_cfg_json=np.asarray(__import__("json").dumps(nn_cfg)))
    np.savez_compressed(os.path.join(HERE, "cu111_flow.npz"), **out, **META)
    print("cu111_flow", tr, va, te, "split", train.indices[:5])


def main():
    if len(sys.argv) > 1:   # selected fixtures only: python make_golden.py li cu111flow
        for what in sys.argv[1:]:
            {"li": make_train_li, "cu111flow": make_cu111_flow}[what]()
        return
    make_train_li()
    make_cu111_flow()
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
