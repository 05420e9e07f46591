"""`nnmd` drop-in: the reference's import names on the B200 path.

Put this repository on PYTHONPATH (ahead of any installed `nnmd`) and the reference's sample scripts run unchanged:

    samples/cu111/train.py:7-10   from nnmd.io import input_parser
                                  from nnmd.nn import BPNN, train_val_test_split
                                  from nnmd.nn.dataset import TrainAtomicDataset
                                  from nnmd.features import calculate_params
    samples/li/train.py:9         from nnmd.nn.dataset import TrainAtomicDataset

`nnmd.features` and `nnmd.nn` (and their sub-modules, under the reference's module names) ARE `nnmd_b200.features` and
`nnmd_b200.nn`: the hot path runs on the sm_100a kernels.  `nnmd.md` is `nnmd_b200.md` under the reference's names
(`NNMD_calc` = the calculator adaptor with the species-by-symbol fix; the reference's own compares symbols
with atomic numbers, nnmd/md/calculator.py:34 vs :42, and selects no atom).  The host-side text parsers `nnmd.io` are NOT
rebuilt here (out of scope): when a copy of the reference package is reachable (environment variable
NNMD_REFERENCE_ROOT = its checkout, or another `nnmd` directory further down sys.path), its directory is appended to this
package's search path, so `import nnmd.io` resolves to the reference's own, unmodified parsers.
"""
import os
import sys

import nnmd_b200
from nnmd_b200 import _util, features, nn
from nnmd_b200.features import auto_params as _auto_params, calculate_sf as _calculate_sf_mod, symm_funcs as _symm_funcs
from nnmd_b200.nn import atomic_nn as _atomic_nn, bpnn as _bpnn, dataset as _dataset, train_val_test_split as _split

__version__ = nnmd_b200.__version__

_ALIASES = {
    "features": features, "features.auto_params": _auto_params, "features.symm_funcs": _symm_funcs,
    "nn": nn, "nn.atomic_nn": _atomic_nn, "nn.bpnn": _bpnn, "nn.dataset": _dataset,
    "_util": _util,
}
# `from nnmd_b200.features import calculate_sf` rebinds the attribute to the FUNCTION; the sub-module is in sys.modules
_ALIASES["features.calculate_sf"] = sys.modules["nnmd_b200.features.calculate_sf"]
_ALIASES["nn.train_val_test_split"] = sys.modules["nnmd_b200.nn.train_val_test_split"]
for _name, _mod in _ALIASES.items():
    sys.modules[f"{__name__}.{_name}"] = _mod


def _reference_package_dir():
    here = os.path.dirname(os.path.abspath(__file__))
    roots = [os.environ.get("NNMD_REFERENCE_ROOT")] + list(sys.path)
    for root in roots:
        if not root:
            continue
        cand = os.path.join(root, "nnmd")
        if os.path.isdir(os.path.join(cand, "io")) and os.path.abspath(cand) != here:
            return cand
    return None


import types

from nnmd_b200 import md as _md

md = types.ModuleType(f"{__name__}.md")
md.__doc__ = "nnmd_b200.md under the reference's names (nnmd/md/__init__.py:1-7)"
md.NNMD_calc = _md.NNMDCalc              # reference name: nnmd/md/calculator.py:10
md.NNMDCalc = _md.NNMDCalc
md.VelocityVerlet = _md.VelocityVerlet   # device-resident replacement of nnmd/md/verlet.py:8-218 (MDSimulation)
md.__all__ = ["NNMD_calc", "NNMDCalc", "VelocityVerlet"]
sys.modules[f"{__name__}.md"] = md

_REF = _reference_package_dir()
if _REF is not None:
    __path__.append(_REF)  # nnmd.io: the reference's own parsers, imported lazily on first use

__all__ = ["features", "nn", "md"]
