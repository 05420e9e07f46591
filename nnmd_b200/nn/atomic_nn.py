"""Per-species atomic network (reference: nnmd/nn/atomic_nn.py:5-48).

Linear -> sigmoid -> ... -> Linear, one output per atom.  Parameters live in `self.model` (an nn.Sequential of
nn.Linear) so `state_dict()` keys are `model.{i}.weight|bias` and the reference's shipped weights
(samples/*/models_*/atomic_nn_*.pt) load unchanged.

Two execution paths:
  * `energy_and_grad(x)`  -- K6 (csrc/mlp.cu): e and dE/dx in one fused CUDA pass, no autograd graph; used by
    `BPNN.predict` and the evaluation engine.
  * `energy_and_grad_autograd(x)` -- the same forward, wired into autograd with K7 (csrc/mlp_train.cu) as the
    backward: gradients of a loss of (e, dE/dx) w.r.t. weights and biases.  `BPNN._train_loop` uses this.
  * `forward(x)`          -- plain torch ops, kept for API compatibility (`net(x)` works like the reference's
    module); not used by the training or prediction paths of this package.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from .. import _lib


class AtomicNN(nn.Module):
    def __init__(self, input_size: int, hidden_sizes, output_size: int = 1, activation=torch.sigmoid):
        super().__init__()
        if isinstance(hidden_sizes, list):
            widths = [input_size] + list(hidden_sizes) + [output_size]
        elif isinstance(hidden_sizes, int):
            widths = [input_size, hidden_sizes, output_size]
        else:
            raise ValueError("hidden_sizes must be a list or an integer")
        self.model = nn.Sequential(*[nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:])])
        self.activation = activation

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        last = len(self.model) - 1
        for i, layer in enumerate(self.model):
            x = layer(x)
            if i != last:
                x = self.activation(x)
        return x

    # ---- CUDA training path ---------------------------------------------------------------------------------
    def energy_and_grad_autograd(self, x: torch.Tensor):
        """(e, dE/dx) on the CUDA kernels (K6), differentiable w.r.t. the parameters through both outputs (K7)."""
        if self.activation is not torch.sigmoid:
            raise NotImplementedError("the CUDA atomic network implements the reference's sigmoid activation only")
        from .functional import atomic_energy_and_grad
        return atomic_energy_and_grad(x, [m.weight for m in self.model], [m.bias for m in self.model])

    # ---- CUDA inference path ------------------------------------------------------------------------------
    def layer_sizes(self) -> list[int]:
        return [self.model[0].in_features] + [m.out_features for m in self.model]

    def tensor_core_shape(self) -> bool:
        """True when csrc/mlp_tc.cu has an instantiation for this network (two hidden layers, <= 96 inputs, <= 64 wide)."""
        s = self.layer_sizes()
        return len(s) == 4 and s[3] == 1 and s[0] <= 96 and s[1] <= 64 and s[2] <= 64

    def fused_shape(self, dtype: torch.dtype) -> bool:
        """True when csrc/mlp_tc.cu has a fused K5 + K6 + K4c instantiation for this network and precision."""
        s = self.layer_sizes()
        return dtype == torch.float32 and self.activation is torch.sigmoid and self.tensor_core_shape() and s[1] > 32

    def energy_force_fused(self, G: torch.Tensor, dG: torch.Tensor, flag: torch.Tensor, out_g: torch.Tensor | None = None):
        """K5 + K6 + K4c in one pass (csrc/mlp_tc.cu): un-normalised G (n_rows, F) and dG (n_rows, F, 3) in, per-atom energies
        (n_rows,) and reference-semantic forces (n_rows, 3) out; the normalised descriptors go to `out_g` when given (may be
        G itself), the NaN bit to `flag`.  fp32, tensor cores (3xTF32)."""
        _lib.require_cuda()
        lib = _lib.load()
        sizes = self.layer_sizes()
        ws = [m.weight.detach().to(device=G.device, dtype=torch.float32).contiguous() for m in self.model]
        bs = [m.bias.detach().to(device=G.device, dtype=torch.float32).contiguous() for m in self.model]
        L = len(ws)
        n_rows = G.shape[0]
        e = torch.empty(n_rows, dtype=torch.float32, device=G.device)
        f = torch.empty((n_rows, 3), dtype=torch.float32, device=G.device)
        c_sizes = (ctypes.c_int32 * (L + 1))(*sizes)
        c_w = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
        c_b = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
        with torch.cuda.device(G.device):
            _lib.check(lib.nnmd_energy_force_fused(L, c_sizes, c_w, c_b, _lib.ptr(G), _lib.ptr(dG), n_rows, _lib.ptr(out_g), _lib.ptr(e),
                                                   None, _lib.ptr(f), _lib.ptr(flag), _lib.stream_ptr()))
        return e, f

    def energy_and_grad(self, x: torch.Tensor, need_grad: bool = True, precision: str = "auto"):
        """K6: per-atom energies e (..., 1) and dE/dx (..., F) for x (..., F) on the GPU; no autograd graph.

        precision: "exact" = fp32/fp64 FMA kernels; "tf32" = tensor cores with 3xTF32 error compensation (fp32 only,
        ~2e-6 relative); "auto" = "tf32" when the data is fp32 and the shape has a tensor-core instantiation."""
        if self.activation is not torch.sigmoid:
            raise NotImplementedError("the CUDA atomic network implements the reference's sigmoid activation only")
        _lib.require_cuda()
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.NnmdError("AtomicNN.energy_and_grad needs CUDA tensors")
        sizes = self.layer_sizes()
        dtype = x.dtype
        x2 = x.detach().reshape(-1, sizes[0]).contiguous()
        n_rows = x2.shape[0]
        ws = [m.weight.detach().to(device=x.device, dtype=dtype).contiguous() for m in self.model]
        bs = [m.bias.detach().to(device=x.device, dtype=dtype).contiguous() for m in self.model]
        L = len(ws)
        e = torch.empty(n_rows, dtype=dtype, device=x.device)
        dE = torch.empty_like(x2) if need_grad else None
        c_sizes = (ctypes.c_int32 * (L + 1))(*sizes)
        c_w = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
        c_b = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
        if precision == "auto":
            precision = "tf32" if (dtype == torch.float32 and self.tensor_core_shape()) else "exact"
        prec = {"exact": _lib.MLP_EXACT, "tf32": _lib.MLP_TF32}[precision]
        with torch.cuda.device(x.device):
            _lib.check(lib.nnmd_mlp_energy_grad(_lib.dtype_code(dtype), prec, L, c_sizes, c_w, c_b, _lib.ptr(x2), n_rows,
                                                _lib.ptr(e), _lib.ptr(dE), _lib.stream_ptr()))
        e = e.view(*x.shape[:-1], 1)
        return e, (dE.view_as(x) if need_grad else None)
