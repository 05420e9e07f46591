"""ctypes binding of the C ABI in include/nnmd_b200.h (libnnmd_b200.so, built in-tree by nnmd_b200._build).

There is NO CPU implementation behind this module: if the library is missing and cannot be built, or no CUDA
device is present, every compute entry point raises.  torch owns device memory and streams; this layer only
passes raw pointers.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _build

F32, F64 = 0, 1
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY = 0, 1, 2, 3
FLAG_NAN_G, FLAG_NAN_DG, FLAG_OVERFLOW = 1, 2, 4
MLP_EXACT, MLP_TF32 = 0, 1

_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "nnmd_abi_version": (ctypes.c_int, []),
    "nnmd_last_error": (ctypes.c_char_p, []),
    "nnmd_device_sm_count": (ctypes.c_int, []),
    "nnmd_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "nnmd_launch_count": (ctypes.c_longlong, []),
    "nnmd_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "nnmd_nvtx_enable": (ctypes.c_int, [ctypes.c_int]),
    "nnmd_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_wrap_positions": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "nnmd_nlist_workspace_bytes": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int64]),
    "nnmd_nlist_count": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "nnmd_nlist_fill": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_nlist_build": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_nlist_row_scratch_bytes": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int64]),
    "nnmd_skin_check": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_sf_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_void_p)]),
    "nnmd_sf_plan_create_species": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "nnmd_set_species": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p]),
    "nnmd_sf_plan_destroy": (None, [ctypes.c_void_p]),
    "nnmd_sf_plan_rc_max": (ctypes.c_double, [ctypes.c_void_p]),
    "nnmd_sf_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_sf_edge_scratch_bytes": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int64]),
    "nnmd_sf_normalize": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_sf_force": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_sf_force_flag": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_virial_workspace_bytes": (ctypes.c_int64, [ctypes.c_int64]),
    "nnmd_virial": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_force_contract": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_mlp_energy_grad": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_energy_force_fused": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_mlp_train_backward": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_force_contract_backward": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_gather_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int64, ctypes.c_void_p]),
    "nnmd_gather_rows_cursor": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "nnmd_loss_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64]),
    "nnmd_loss_mse": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_peer_handle_bytes": (ctypes.c_int, []),
    "nnmd_peer_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p)]),
    "nnmd_peer_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_peer_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_peer_destroy": (None, [ctypes.c_void_p]),
    "nnmd_peer_capacity": (ctypes.c_int64, [ctypes.c_void_p]),
    "nnmd_peer_status": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "nnmd_peer_allreduce_weighted_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "nnmd_peer_allreduce_step_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p]),
    "nnmd_peer_allreduce_sum_f32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "nnmd_peer_halo_exchange": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class NnmdError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first when a source is newer) the CUDA library.  Raises when that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    override = os.environ.get("NNMD_B200_LIB")  # tuning / debugging: load a differently built library
    if override:
        path, build_if_missing = override, False
    if build_if_missing:
        try:
            path = _build.build()
        except Exception as exc:  # no nvcc on the box: fall through to the prebuilt file
            if not os.path.exists(path):
                raise NnmdError(f"libnnmd_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(path):
        raise NnmdError("libnnmd_b200.so is missing (run `python -m nnmd_b200._build`); there is no CPU fallback")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != OK:
        msg = load().nnmd_last_error().decode("utf-8", "replace")
        if code == ERR_INVALID and msg.startswith("Unknown symmetry function number"):
            raise ValueError(msg)
        raise NnmdError(f"nnmd_b200 error {code}: {msg}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise NnmdError("nnmd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.float64:
        return F64
    raise TypeError(f"nnmd_b200 computes in float32 or float64, got {dtype}")


PROF_CLASSES = ("nlist", "radial", "angular", "normalize", "contract", "mlp", "wrap", "gather")


def profile_enable(on: bool) -> None:
    check(load().nnmd_profile_enable(int(on)))


def nvtx_enable(on: bool = True) -> None:
    """NVTX ranges around every kernel class of the library (visible to nsys / `ncu --nvtx`); off by default."""
    check(load().nnmd_nvtx_enable(int(on)))


def profile_read() -> dict:
    """{class: (total ms, launches)} of the spans recorded since the last read (synchronises the device)."""
    n = len(PROF_CLASSES)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_longlong * n)()
    check(load().nnmd_profile_read(ms, cnt))
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(PROF_CLASSES)}


def set_option(name: str, value: int) -> None:
    """Process-wide kernel-selection switch of the library (see include/nnmd_b200.h: nnmd_set_option)."""
    check(load().nnmd_set_option(name.encode(), int(value)))


def launch_count() -> int:
    return int(load().nnmd_launch_count())


def ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
