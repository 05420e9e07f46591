"""Autograd bindings of the training kernels (K6 forward, K7 backward, K4c / K4c').

`atomic_energy_and_grad(x, weights, biases)` returns (e, dE/dx) like
    e = net(x); dE = torch.autograd.grad(e.sum(), x, create_graph=True)[0]          (nnmd/nn/bpnn.py:343-349)
and is differentiable w.r.t. the network parameters through BOTH outputs (the double backward the force loss needs);
`force_contract(dE, dG)` is  -einsum("ijk,ijkl->ijl", dE, dG)  (bpnn.py:353), differentiable w.r.t. dE.
There is no gradient w.r.t. x or dG: the training step never needs one (the reference accumulates a useless .grad on
the cached descriptors, SURVEY 3.2 note ii).  CUDA only.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib


def _param_arrays(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _sizes(weights):
    sizes = [weights[0].shape[1]] + [w.shape[0] for w in weights]
    return sizes, (ctypes.c_int32 * len(sizes))(*sizes)


class _AtomicEnergyGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_layers, *params):
        _lib.require_cuda()
        if not x.is_cuda:
            raise _lib.NnmdError("the atomic network runs on CUDA tensors only (no CPU fallback)")
        lib = _lib.load()
        dtype = x.dtype
        weights = [p.detach().to(dtype).contiguous() for p in params[:n_layers]]
        biases = [p.detach().to(dtype).contiguous() for p in params[n_layers:]]
        sizes, c_sizes = _sizes(weights)
        x2 = x.detach().reshape(-1, sizes[0]).contiguous()
        n_rows = x2.shape[0]
        e = torch.empty(n_rows, dtype=dtype, device=x.device)
        dE = torch.empty_like(x2)
        with torch.cuda.device(x.device):
            _lib.check(lib.nnmd_mlp_energy_grad(_lib.dtype_code(dtype), _lib.MLP_EXACT, n_layers, c_sizes, _param_arrays(weights),
                                                _param_arrays(biases), _lib.ptr(x2), n_rows, _lib.ptr(e), _lib.ptr(dE),
                                                _lib.stream_ptr()))
        ctx.n_layers = n_layers
        ctx.x_shape = x.shape
        ctx.param_dtypes = [p.dtype for p in params]
        ctx.save_for_backward(x2, *weights, *biases)
        return e.view(*x.shape[:-1], 1), dE.view(x.shape)

    @staticmethod
    def backward(ctx, grad_e, grad_w):
        lib = _lib.load()
        L = ctx.n_layers
        x2, *pb = ctx.saved_tensors
        weights, biases = pb[:L], pb[L:]
        sizes, c_sizes = _sizes(weights)
        dtype = x2.dtype
        n_rows = x2.shape[0]
        ge = (torch.zeros(n_rows, dtype=dtype, device=x2.device) if grad_e is None
              else grad_e.detach().to(dtype).reshape(-1).contiguous())
        u = None if grad_w is None else grad_w.detach().to(dtype).reshape(-1, sizes[0]).contiguous()
        gW = [torch.zeros_like(w) for w in weights]
        gb = [torch.zeros_like(b) for b in biases]
        with torch.cuda.device(x2.device):
            _lib.check(lib.nnmd_mlp_train_backward(_lib.dtype_code(dtype), L, c_sizes, _param_arrays(weights), _param_arrays(biases),
                                                   _lib.ptr(x2), n_rows, _lib.ptr(ge), _lib.ptr(u), _param_arrays(gW),
                                                   _param_arrays(gb), _lib.stream_ptr()))
        grads = [g.to(dt) for g, dt in zip(gW + gb, ctx.param_dtypes)]
        return (None, None, *grads)


def atomic_energy_and_grad(x: torch.Tensor, weights, biases):
    """(e (...,1), dE/dx (...,F)) of the sigmoid MLP given by torch.nn.Linear-layout weights/biases."""
    return _AtomicEnergyGrad.apply(x, len(weights), *weights, *biases)


class _ForceContract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dE, dG):
        lib = _lib.load()
        nf = dE.shape[-1]
        dE2 = dE.detach().reshape(-1, nf).contiguous()
        dG2 = dG.detach().to(dE.dtype).reshape(-1, nf, 3).contiguous()
        out = torch.empty((dE2.shape[0], 3), dtype=dE.dtype, device=dE.device)
        with torch.cuda.device(dE.device):
            _lib.check(lib.nnmd_force_contract(_lib.dtype_code(dE.dtype), _lib.ptr(dE2), _lib.ptr(dG2), dE2.shape[0], nf,
                                               _lib.ptr(out), _lib.stream_ptr()))
        ctx.save_for_backward(dG2)
        ctx.dE_shape = dE.shape
        return out.view(*dE.shape[:-1], 3)

    @staticmethod
    def backward(ctx, gf):
        lib = _lib.load()
        (dG2,) = ctx.saved_tensors
        nf = dG2.shape[1]
        gf2 = gf.detach().to(dG2.dtype).reshape(-1, 3).contiguous()
        u = torch.empty((dG2.shape[0], nf), dtype=dG2.dtype, device=dG2.device)
        with torch.cuda.device(dG2.device):
            _lib.check(lib.nnmd_force_contract_backward(_lib.dtype_code(dG2.dtype), _lib.ptr(gf2), _lib.ptr(dG2), dG2.shape[0], nf,
                                                        _lib.ptr(u), _lib.stream_ptr()))
        return u.view(ctx.dE_shape), None


def force_contract(dE: torch.Tensor, dG: torch.Tensor) -> torch.Tensor:
    """F[...,a,:] = - sum_f dE[...,a,f] dG[...,a,f,:]   (differentiable w.r.t. dE)."""
    return _ForceContract.apply(dE, dG)
