"""The `nnmd` drop-in package resolves the reference's import names to the B200 implementation (no GPU needed)."""
import importlib
import os
import sys

import pytest


def test_reference_import_names_resolve_to_the_b200_package():
    import nnmd
    import nnmd_b200
    from nnmd.features import calculate_params, calculate_sf, SymmetryFunction, f_cutoff, g1_function, g2_function, g4_function, \
        g5_function, calculate_distances                                    # nnmd/features/__init__.py:13-23
    from nnmd.nn import AtomicNN, BPNN, TrainAtomicDataset, train_val_test_split   # nnmd/nn/__init__.py:1-6
    from nnmd.nn.dataset import TrainAtomicDataset as T2                    # samples/li/train.py:9
    from nnmd.md import NNMD_calc                                           # samples/*/md_simulation.py
    assert calculate_sf is nnmd_b200.features.calculate_sf and BPNN is nnmd_b200.nn.BPNN and T2 is TrainAtomicDataset
    assert NNMD_calc is nnmd_b200.md.NNMDCalc
    assert importlib.import_module("nnmd.features.symm_funcs") is importlib.import_module("nnmd_b200.features.symm_funcs")
    assert importlib.import_module("nnmd.nn.bpnn") is importlib.import_module("nnmd_b200.nn.bpnn")
    assert calculate_params(6.0, 0, 2, 0, 2, 0)[0][0] == 6.0
    assert nnmd.__file__.startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.skipif(not os.path.isdir("/root/reference/nnmd/io"), reason="the reference checkout exists in the build container only")
def test_reference_parsers_load_through_the_shim():
    """With NNMD_REFERENCE_ROOT set, `nnmd.io` is the reference's own unmodified parser package (run in a fresh interpreter:
    the search path is fixed when `nnmd` is first imported)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, types\n"
        "ase = types.ModuleType('ase'); ase.Atoms = object; aio = types.ModuleType('ase.io'); aio.Trajectory = object\n"
        "sys.modules.update({'ase': ase, 'ase.io': aio})\n"      # `ase` is not installed here; gpaw_parser does not need it
        "from nnmd.io import input_parser\n"
        "from nnmd.nn import BPNN\n"
        "import os; os.chdir('/root/reference/samples/cu111')\n"
        "d = input_parser('input/input.yaml')\n"
        "print(input_parser.__module__, BPNN.__module__, len(d['atomic_data']['reference_data']), d['neural_network']['input_sizes'])\n")
    env = dict(os.environ, PYTHONPATH=root, NNMD_REFERENCE_ROOT="/root/reference")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["nnmd.io.input_parser", "nnmd_b200.nn.bpnn", "5", "[12]"]
