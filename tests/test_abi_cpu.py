"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol include/nnmd_b200.h declares;
host logic (plan validation, calculate_params, dense helpers) behaves like the reference.  No compute calls."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    from nnmd_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "nnmd_b200.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(nnmd_[a-z0-9_]+)\s*\(", header))
    declared.discard("nnmd_sf_plan")
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.nnmd_abi_version() == 1


def test_compute_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from nnmd_b200 import _lib
    from nnmd_b200.features import calculate_sf
    with pytest.raises(_lib.NnmdError, match="no CPU fallback"):
        calculate_sf(torch.zeros(1, 3, 3), torch.zeros(3, 3), torch.zeros(3), {"features": [1], "params": [[3.0]]})
    # the atomic network and the fused training step have no CPU path either
    from nnmd_b200.dist import FlatGradients
    from nnmd_b200.nn import AtomicNN
    from nnmd_b200.nn.functional import TrainStepPlan
    net = AtomicNN(3, [4, 4])
    with pytest.raises(_lib.NnmdError):
        net.energy_and_grad(torch.zeros(2, 3))
    flat = FlatGradients(list(net.parameters()))
    with pytest.raises(_lib.NnmdError):
        TrainStepPlan([net], ["X"], {"X": (2, 5, 3)}, 2, "cpu", 1.0, 1.0, flat)


def test_calculate_params_matches_reference_vectors(golden_dir):
    from nnmd_b200.features import calculate_params
    d = np.load(os.path.join(golden_dir, "calculate_params.npz"))
    n = 0
    for key in d.files:
        if not key.startswith("args_"):
            continue
        args = [float(a) if "." in a else int(a) for a in key[5:].split("_")]
        rows = calculate_params(*args)
        arr = np.full((len(rows), 4), np.nan)
        for i, r in enumerate(rows):
            arr[i, :len(r)] = [float(x) for x in r]
        assert np.array_equal(arr, d[key], equal_nan=True), key
        n += 1
    assert n == 4
    with pytest.raises(ZeroDivisionError):
        calculate_params(6.0, 0, 1, 0, 0, 0)


def test_dense_helpers_match_oracle():
    """g*_function(distances, ...) helpers (kept as thin torch code for the plotting samples)."""
    from nnmd_b200 import features as F
    from oracle import dense_oracle as O2
    torch.manual_seed(0)
    pos = torch.rand(2, 12, 3, dtype=torch.float64) * 5
    cell = torch.tensor([[5.0, 0, 0], [0.5, 5, 0], [0, 0, 5]], dtype=torch.float64)
    pbc = torch.tensor([1.0, 1, 0], dtype=torch.float64)
    d = F.calculate_distances(pos - 2.0, cell, pbc)
    assert torch.allclose(d, O2.distance_matrix(pos - 2.0, cell, pbc), atol=1e-14)
    assert torch.allclose(F.f_cutoff(d, 3.0), O2.cutoff_fn(d, 3.0))
    assert torch.allclose(F.g1_function(d, 3.0), O2._radial(d, 1, [3.0]))
    assert torch.allclose(F.g2_function(d, 3.0, 0.7, 1.1), O2._radial(d, 2, [3.0, 0.7, 1.1]))
    assert torch.allclose(F.g4_function(d, 3.0, 0.1, 2.5, -1), O2._angular(d, 4, [3.0, 0.1, 2.5, -1], "exact"))
    assert torch.allclose(F.g5_function(d, 3.0, 0.1, 3, 1), O2._angular(d, 5, [3.0, 0.1, 3, 1], "exact"))
    assert [m.value for m in F.SymmetryFunction] == [1, 2, 4, 5]


def test_atomic_nn_state_dict_layout():
    from nnmd_b200.nn import AtomicNN
    net = AtomicNN(14, [64, 64])
    assert list(net.state_dict().keys()) == [f"model.{i}.{k}" for i in range(3) for k in ("weight", "bias")]
    assert AtomicNN(5, 7).layer_sizes() == [5, 7, 1]
    with pytest.raises(ValueError, match="hidden_sizes must be a list or an integer"):
        AtomicNN(3, (4, 4))
    x = torch.rand(2, 5, 14)
    assert net(x).shape == (2, 5, 1)


def test_bpnn_config_host_logic(tmp_path):
    """Host side of BPNN.config / save_model (reference: bpnn.py:107-194, 556-565): optimizer table and its error text,
    Xavier initialisation with zero biases, amsgrad and the reference's betas, checkpoint file names."""
    from nnmd_b200.nn import BPNN
    cfg = dict(atom_species=["A", "B"], input_sizes=[4, 6], hidden_sizes=[[5, 5], [7]], learning_rate=0.01, l2_regularization=1e-3,
               e_loss_coeff=1.0, f_loss_coeff=2.0, load_models=False, path=str(tmp_path), optimizer="sgd", batch_size=4, epochs=1)
    net = BPNN(dtype=torch.float32, use_cuda=False)
    with pytest.raises(ValueError, match="Optimizer sgd is not available"):
        net.config(cfg)
    for name, betas in (("adam", (0.8, 0.9999)), ("adamw", (0.9, 0.999)), (None, (0.8, 0.9999))):
        net = BPNN(dtype=torch.float32, use_cuda=False)
        net.config(dict(cfg, optimizer=name))
        grp = net.optim.param_groups[0]
        assert grp["amsgrad"] is True and tuple(grp["betas"]) == betas and grp["weight_decay"] == 1e-3
        assert [a.layer_sizes() for a in net.atomic_nets] == [[4, 5, 5, 1], [6, 7, 1]]
        assert all(float(m.bias.detach().abs().max()) == 0.0 for a in net.atomic_nets for m in a.model)
        assert isinstance(net.criterion, torch.nn.MSELoss) and net.species == ["A", "B"]
    net.save_model(str(tmp_path))
    assert sorted(p.name for p in tmp_path.iterdir()) == ["atomic_nn_A.pt", "atomic_nn_B.pt"]
    again = BPNN(dtype=torch.float32, use_cuda=False)
    again.config(dict(cfg, optimizer="adam", load_models=True))
    for a, b in zip(net.atomic_nets, again.atomic_nets):
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.equal(p, q)


def test_batch_loader_and_split_host_logic():
    """FrameBatches reproduces DataLoader(dataset, batch_size, shuffle) semantics on the cached tensors (bpnn.py:234-236):
    every frame exactly once per epoch, ragged last batch, Subset chains from train_val_test_split resolved to indices."""
    from nnmd_b200.nn import TrainAtomicDataset, train_val_test_split
    from nnmd_b200.nn.bpnn import FrameBatches, batch_loader
    M = 23
    g = torch.arange(M, dtype=torch.float32).view(M, 1, 1).expand(M, 3, 2).contiguous()
    dg = torch.zeros(M, 3, 2, 3)
    ds = TrainAtomicDataset({"X": (g, dg)}, torch.arange(M, dtype=torch.float32).view(M, 1), torch.zeros(M, 3, 3), {})
    loader = batch_loader(ds, 5, shuffle=True)
    assert isinstance(loader, FrameBatches) and len(loader) == 5
    torch.manual_seed(0)
    seen = []
    for gb, dgb, e, f in loader:
        assert gb["X"].shape[1:] == (3, 2) and dgb["X"].shape[1:] == (3, 2, 3) and f.shape[1:] == (3, 3)
        assert torch.equal(gb["X"][:, 0, 0], e[:, 0])  # rows of every cached tensor move together
        seen += e[:, 0].tolist()
    assert sorted(seen) == list(range(M)) and seen != sorted(seen)
    sizes = [int(s.numel()) for s in loader.index_batches("cpu")]
    assert sizes == [5, 5, 5, 5, 3]
    train, val, test = train_val_test_split(ds, (0.6, 0.2, 0.2))
    parts = [batch_loader(p, 4, shuffle=False) for p in (train, val, test)]
    assert all(isinstance(p, FrameBatches) for p in parts)
    ids = [sorted(int(i) for b in p.index_batches("cpu") for i in b) for p in parts]
    assert sorted(ids[0] + ids[1] + ids[2]) == list(range(M)) and [len(i) for i in ids] == [len(train), len(val), len(test)]
    # anything that is not a TrainAtomicDataset (or a Subset chain of one) falls back to the stock DataLoader
    assert isinstance(batch_loader(torch.utils.data.TensorDataset(torch.zeros(4, 2)), 2, shuffle=False), torch.utils.data.DataLoader)
