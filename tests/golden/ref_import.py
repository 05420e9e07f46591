"""Import the UNMODIFIED reference (ded-otshelnik/nnmd, /root/reference) in the
build container.  Only used by tests/golden/make_golden.py to produce the
committed fixtures; nothing under tests/ (run time), bench.py or the package
imports this module, because /root/reference does not exist on the GPU box.

The reference's package __init__ pulls in nnmd.md -> `ase`, which is not
installed here; four empty stand-in modules make the import succeed
(SURVEY.md App. B).
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def import_reference():
    if "ase" not in sys.modules:
        ase = types.ModuleType("ase")
        ase.Atoms = object
        calcs = types.ModuleType("ase.calculators")
        calc = types.ModuleType("ase.calculators.calculator")

        class Calculator:  # never instantiated
            def __init__(self, *a, **k):
                pass

        calc.Calculator = Calculator
        aio = types.ModuleType("ase.io")
        aio.Trajectory = object
        sys.modules.update({"ase": ase, "ase.calculators": calcs,
                            "ase.calculators.calculator": calc, "ase.io": aio})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import nnmd  # noqa: F401  (the reference package)
    import nnmd.nn.bpnn as bp
    bp.tqdm = lambda it, **k: it
    return nnmd
