"""Build nnmd_b200/csrc/*.cu into nnmd_b200/csrc/libnnmd_b200.so with nvcc for sm_100a.

The library is a plain C ABI (include/nnmd_b200.h); it is built IN-TREE so that it travels to the GPU box
with the repository snapshot.  `python -m nnmd_b200._build` rebuilds when a source is newer than the library.
"""
from __future__ import annotations

import contextlib
import fcntl
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libnnmd_b200.so")
SOURCES = ["api.cu", "nlist.cu", "sf.cu", "mlp.cu", "mlp_tc.cu", "mlp_train.cu", "mlp_fast.cu", "loss.cu", "peer.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


@contextlib.contextmanager
def _build_lock():
    """One builder at a time per checkout: N ranks of a torchrun launch would otherwise run nvcc into the same files."""
    fd = os.open(os.path.join(CSRC, ".build.lock"), os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile stale objects and relink.  Serialised by a file lock; objects and the library are written under temporary
    names and renamed into place, so a concurrent loader never maps a half-written file."""
    with _build_lock():
        return _build_locked(force, verbose)


def _build_locked(force: bool, verbose: bool) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "nnmd_b200.h"))
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o + ".tmp.o"]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            results = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs))
        for cmd, r in zip(jobs, results):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + cmd[-3])
        for cmd in jobs:
            os.replace(cmd[-1], cmd[-1][:-len(".tmp.o")])
    if jobs or force or _stale(LIB, objs):
        tmp = LIB + ".tmp"
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
