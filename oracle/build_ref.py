"""Stage the UNMODIFIED reference (ded-otshelnik/nnmd) into oracle/_ref/ -- TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it is a pip install of its own pyproject into a target directory:

    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>

(`--no-deps`: its requirements -- torch, numpy, pyyaml, tqdm, ase -- are resolved from the image; `ase` is absent and is
stubbed by `load_reference`, exactly as SURVEY.md App. B does.  The install runs from a copy under /tmp because
/root/reference is read-only and setuptools writes build/ and *.egg-info into the source tree.)

oracle/_ref/ is git-ignored (no reference source ever enters the history) but travels to the GPU box with the gpurun
snapshot, where `bench.py --impl reference` times it on the host cores.  Only `__graft_entry__.build()` (staging),
`bench.py --impl reference` / its cpu_baseline legs and tests/ may touch this; the product package never does.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "nnmd", "features"))


def build(force: bool = False) -> str | None:
    """Install the reference into oracle/_ref/ when its checkout is present (build container only).  Returns the
    directory, or None when there is nothing to install from and nothing staged."""
    if available() and not force:
        return REF_DIR
    if not os.path.isdir(REFERENCE_ROOT):
        return REF_DIR if available() else None
    tmp = tempfile.mkdtemp(prefix="nnmd_ref_src_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REFERENCE_ROOT, src, ignore=shutil.ignore_patterns(".git", "*.pt", "*.traj", "*.png", "*.txt", "docs"))
        for root, dirs, files in os.walk(src):
            for name in dirs + files:
                os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)
        os.chmod(src, 0o755)
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF_DIR, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout + r.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return REF_DIR


def load_reference():
    """Import the staged reference package (as `nnmd`, from oracle/_ref) and return it.  Stubs the absent `ase` the way
    SURVEY.md App. B does and neutralises the two defects that stop `BPNN` from running at all (tqdm NameError without
    DEBUG, bpnn.py:19-22 vs :327).  Must be called before anything else imported a module called `nnmd`."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() in the build container)")
    if "nnmd" in sys.modules and not os.path.abspath(getattr(sys.modules["nnmd"], "__file__", "")).startswith(REF_DIR):
        raise RuntimeError("another package named nnmd is already imported (the drop-in shim?): load the reference first")
    if "ase" not in sys.modules:
        try:
            import ase  # noqa: F401
        except ImportError:
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
            sys.modules.update({"ase": ase, "ase.calculators": calcs, "ase.calculators.calculator": calc, "ase.io": aio})
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import nnmd
    import nnmd.nn.bpnn as bp
    if not hasattr(bp, "tqdm"):
        bp.tqdm = lambda it, **k: it
    return nnmd


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
