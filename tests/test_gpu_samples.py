"""Sample-level drop-in (BASELINE.json north_star: "drops into the existing samples unchanged"): the statements of
samples/cu111/train.py:15-66 are executed through the `nnmd` shim package, on the parser output the reference produced for
the first 48 frames of samples/gpaw/cu111.txt, and compared with what the UNMODIFIED reference computed for the same
sequence on the CPU (tests/golden/cu111_flow.npz, written by tests/golden/make_golden.py:make_cu111_flow)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, col_relerr, relerr

pytestmark = pytest.mark.gpu


def _parser_output(d):
    """The dict `nnmd.io.input_parser("input/input.yaml")` returns (input_parser.py:27-54), rebuilt from the fixture."""
    frames = [{"Cu": {"positions": p}, "forces": f, "energy": float(e)} for p, f, e in zip(d["positions"], d["forces"], d["energy"])]
    nn = json.loads(str(d["nn_cfg_json"]))
    return {"atomic_data": {"reference_data": frames, "symmetry_functions_set": {}, "n_atoms": int(d["n_atoms"]),
                            "unit_cell": d["unit_cell"], "pbc": d["pbc"]},
            "neural_network": nn}


def test_cu111_train_script_sequence_through_the_nnmd_shim(tmp_path, monkeypatch):
    # samples/cu111/train.py:7-10, verbatim (nnmd.io is the reference's own parser module and is not rebuilt: its OUTPUT is
    # the fixture)
    from nnmd.nn import BPNN, train_val_test_split
    from nnmd.nn.dataset import TrainAtomicDataset
    from nnmd.features import calculate_params
    import nnmd_b200
    assert BPNN is nnmd_b200.nn.BPNN and TrainAtomicDataset is nnmd_b200.nn.TrainAtomicDataset

    d = np.load(os.path.join(GOLDEN, "cu111_flow.npz"))
    monkeypatch.chdir(tmp_path)                       # the script runs in its sample directory: caches, net.log, models
    os.mkdir("models_adamw")                          # the shipped samples/cu111/models_adamw/atomic_nn_Cu.pt
    torch.save({k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("w0_")}, "models_adamw/atomic_nn_Cu.pt")
    dtype = torch.float32
    input_data = _parser_output(d)

    # ---- train.py:19-31
    N1, N2, N3, N4, N5 = 0, 25, 0, 0, 25
    r_cutoff = 38.0
    params = calculate_params(r_cutoff, N1, N2, N3, N4, N5)
    features = [1] * N1 + [2] * N2 + [3] * N3 + [4] * N4 + [5] * N5
    input_data["atomic_data"]["symmetry_functions_set"]["Cu"] = {"params": params, "features": features}
    input_data["neural_network"]["input_sizes"] = [N1 + N2 + N3 + N4 + N5]
    # ---- train.py:35
    dataset = TrainAtomicDataset.make_atomic_dataset(input_data["atomic_data"])
    assert sorted(f for f in os.listdir(".") if f.endswith(".pt")) == ["dg_Cu.pt", "g_Cu.pt"]
    g, dg = dataset.sf_data["Cu"]
    assert relerr(g[:2].detach().cpu(), d["g_head"]) < 1e-5
    # two fp32 evaluations (the reference's on the CPU, this one) of the rc = 38 A auto-parameter set, whose G2 rows bind
    # eta := r_s up to 38: the summed gradients of those steep columns move by ~1e-4 with the last bit of the
    # wrapped positions.  Accuracy against the fp64 truth is what tests/test_gpu_parity.py holds to 1e-5.
    assert col_relerr(dg[:2].cpu(), d["dg_head"], 2) < 3e-4
    assert relerr(dataset.energies.cpu(), d["e_scaled"]) < 1e-6 and relerr(dataset.forces[:2].cpu(), d["f_scaled_head"]) < 1e-6
    # ---- train.py:38-41: the split draws from the global RNG that calculate_sf left seeded with 42
    train_dataset, val_dataset, test_dataset = train_val_test_split(dataset, (0.8, 0.1, 0.1))
    assert list(train_dataset.indices) == d["split_train"].tolist() and list(test_dataset.indices) == d["split_test"].tolist()
    # ---- train.py:52-53 (load_models: True -> the shipped weights)
    net = BPNN(dtype=dtype)
    net.config(input_data["neural_network"])
    # energy + forces with the shipped weights (what plot_energies.py / the ASE calculator ask predict for)
    cart = torch.tensor(d["positions"][:3], dtype=dtype)
    E, F = net.predict({"Cu": cart}, torch.tensor(d["unit_cell"], dtype=dtype), {"Cu": input_data["atomic_data"]["symmetry_functions_set"]["Cu"]},
                       pbc=d["pbc"])
    assert relerr(E.cpu(), d["pred_e"]) < 1e-5
    assert float((F.cpu() - torch.tensor(d["pred_f"])).abs().max()) < 1e-4
    # ---- train.py:57-66, two epochs
    input_data["neural_network"]["epochs"] = int(d["epochs"])
    net.fit(train_dataset, val_dataset, test_dataset)
    os.mkdir("out")
    net.save_model("out")
    sd = torch.load("out/atomic_nn_Cu.pt", map_location="cpu")
    assert set(sd) == {f"model.{i}.{w}" for i in range(3) for w in ("weight", "bias")}
    log = open("net.log").read()
    tr = [[float(x.split("=")[1].split(",")[0].split()[0]) for x in line.split("training:")[1].split("MSELoss")[1:]]
          for line in log.splitlines() if "training:" in line]
    va = [[float(x.split("=")[1].split(",")[0].split()[0]) for x in line.split("validation:")[1].split("MSELoss")[1:]]
          for line in log.splitlines() if "validation:" in line]
    print("cu111 flow: train", tr, "ref", d["losses_train"].tolist(), "val", va, "ref", d["losses_val"].tolist())
    assert "Training sample size:   38" in log and "Learning rate changed to 0.099\n" in log and "Learning rate changed to 0.09801\n" in log
    # same frames in the same batches (DataLoader-compatible shuffling), same start: both epochs must follow the reference
    # (measured: 3e-6 on the training losses, 2e-5 on the validation losses of both epochs)
    np.testing.assert_allclose(tr, d["losses_train"], rtol=1e-3)
    np.testing.assert_allclose(va, d["losses_val"], rtol=1e-3)
    assert len(tr) == 2
