"""ctypes front-end of oracle/nl_oracle.c (O3: O(N) cell-list CPU oracle, fp64, OpenMP).

TEST INFRASTRUCTURE ONLY -- see the header of nl_oracle.c.  `build()` compiles the C
file with gcc into oracle/_build/ (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

from . import dense_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "nl_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libnl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.nl_oracle_descriptors.restype = ctypes.c_int
        lib.nl_oracle_descriptors.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int]
        lib.nl_oracle_pairs.restype = ctypes.c_int64
        lib.nl_oracle_pairs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_int]
        lib.nl_oracle_max_threads.restype = ctypes.c_int
        lib.nl_oracle_threshold_margin.restype = ctypes.c_int
        lib.nl_oracle_threshold_margin.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().nl_oracle_max_threads())


def pack_params(sfd: dict):
    """features/params (reference's positional binding) -> int32 kinds[nf], float64 prm[nf,4]."""
    kinds = np.asarray([int(k) for k in sfd["features"]], dtype=np.int32)
    prm = np.zeros((len(kinds), 4), dtype=np.float64)
    for f, (k, p) in enumerate(zip(kinds, sfd["params"])):
        if k not in (1, 2, 4, 5):
            raise ValueError(f"Unknown symmetry function number: {k}")
        need = {1: 1, 2: 3, 4: 4, 5: 4}[int(k)]
        if len(p) != need:
            raise TypeError(f"G{k} takes {need} parameters, got {len(p)}")
        prm[f, :need] = [float(x) for x in p]
    return kinds, prm


def raw_descriptors(pos: torch.Tensor, cell: torch.Tensor, pbc: torch.Tensor, sfd: dict, nthreads: int = 0):
    """Un-normalised G (B,N,F), dg (B,N,F,3) and neighbour counts (B,N); fp64 on CPU."""
    lib = _load()
    pos64 = pos.detach().to("cpu", torch.float64)
    wrapped = dense_oracle.wrap_positions(pos64, cell.detach().to("cpu", torch.float64),
                                          pbc.detach().to("cpu", torch.float64)).contiguous().numpy()
    B, N, _ = wrapped.shape
    kinds, prm = pack_params(sfd)
    nf = len(kinds)
    G = np.zeros((B, N, nf), dtype=np.float64)
    dG = np.zeros((B, N, nf, 3), dtype=np.float64)
    cnt = np.zeros((B, N), dtype=np.int32)
    rc = lib.nl_oracle_descriptors(wrapped.ctypes.data, B, N, kinds.ctypes.data, prm.ctypes.data, nf,
                                   G.ctypes.data, dG.ctypes.data, cnt.ctypes.data, int(nthreads))
    assert rc == 0
    return torch.from_numpy(G), torch.from_numpy(dG), torch.from_numpy(cnt)


def descriptors(pos, cell, pbc, sfd, nthreads: int = 0):
    """calculate_sf semantics: (g_hat, dg) in fp64 (calculate_sf.py:100-109)."""
    G, dG, _ = raw_descriptors(pos, cell, pbc, sfd, nthreads)
    ghat = G / torch.linalg.vector_norm(G, dim=-1, keepdim=True)
    if torch.isnan(ghat).any():
        raise ValueError("NaN values in the symmetry functions")
    if torch.isnan(dG).any():
        raise ValueError("NaN values in the gradients of the symmetry functions")
    return ghat, dG


def pairs(pos_frame: torch.Tensor, cell, pbc, rc: float, nthreads: int = 0) -> np.ndarray:
    """Row-sorted (i,j) pair set of ONE frame with 0 < d_ij < rc (fp64 arithmetic)."""
    lib = _load()
    p = pos_frame.detach().to("cpu", torch.float64)[None]
    wrapped = dense_oracle.wrap_positions(p, cell.detach().to("cpu", torch.float64),
                                          pbc.detach().to("cpu", torch.float64))[0].contiguous().numpy()
    n = wrapped.shape[0]
    total = lib.nl_oracle_pairs(wrapped.ctypes.data, n, float(rc), None, 0, int(nthreads))
    out = np.zeros((max(total, 1), 2), dtype=np.int32)
    lib.nl_oracle_pairs(wrapped.ctypes.data, n, float(rc), out.ctypes.data, total, int(nthreads))
    return out[:total]


def threshold_margin(pos_frame: torch.Tensor, rc: float, nthreads: int = 0) -> torch.Tensor:
    """Per atom of ONE frame of already wrapped positions: the smallest | |cos| - 1e-3 | over the cosines of all closed
    triangles containing the atom -- its distance from the reference's angular on/off threshold (symm_funcs.py:210-211).
    Brute-force neighbour search: sub-blocks of ~1e4 atoms."""
    lib = _load()
    p = pos_frame.detach().to("cpu", torch.float64).contiguous().numpy()
    out = np.zeros(p.shape[0], dtype=np.float64)
    rcode = lib.nl_oracle_threshold_margin(p.ctypes.data, p.shape[0], float(rc), out.ctypes.data, int(nthreads))
    assert rcode == 0
    return torch.from_numpy(out)
