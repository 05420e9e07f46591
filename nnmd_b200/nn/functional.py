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


class TrainStepPlan:
    """Static buffers and the launch sequence of one training step for ONE batch shape, without an autograd graph:

        K6 (e, dE/dg per species) -> K4c (forces) -> K8 (loss, dL/de, dL/df) -> K4c' (dL/d(dE/dg)) -> K7 (weight gradients)

    i.e. bpnn.py:342-366 (forward, `_loss`, `loss.backward()`) in 2 + 4 * n_species launches.  The gradients are ADDED to
    the `.grad` tensors of the parameters (views of a FlatGradients buffer that `run` zeroes first).  All buffers are
    static, so `run` can be captured into a CUDA graph and replayed; `load` / `load_indexed` refill them.
    float32 CUDA tensors only (the reference trains in float32: dataset.py:48, atomic_nn.py).
    """

    def __init__(self, nets, species, shapes, n_frames, device, e_coeff, f_coeff, flat_grads):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.species, self.nets, self.device, self.flat = list(species), list(nets), torch.device(device), flat_grads
        self.n_frames, self.e_coeff, self.f_coeff = int(n_frames), float(e_coeff), float(f_coeff)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.n_atoms = [int(shapes[s][1]) for s in self.species]
        self.n_feat = [int(shapes[s][2]) for s in self.species]
        B = self.n_frames
        self.g = {s: torch.empty(B, n, f, **f32) for s, n, f in zip(self.species, self.n_atoms, self.n_feat)}
        self.dG = {s: torch.empty(B, n, f, 3, **f32) for s, n, f in zip(self.species, self.n_atoms, self.n_feat)}
        self.e = [torch.empty(B * n, **f32) for n in self.n_atoms]
        self.dE = [torch.empty(B * n, f, **f32) for n, f in zip(self.n_atoms, self.n_feat)]
        self.f = [torch.empty(B * n, 3, **f32) for n in self.n_atoms]
        self.ge = [torch.empty(B * n, **f32) for n in self.n_atoms]
        self.gf = [torch.empty(B * n, 3, **f32) for n in self.n_atoms]
        self.u = [torch.empty(B * n, f, **f32) for n, f in zip(self.n_atoms, self.n_feat)]
        self.E = torch.empty(B, **f32)
        self.F = torch.empty(B, sum(self.n_atoms), 3, **f32)
        self.out = torch.zeros(3, **f32)   # (energy term, force term, total) of the last step
        self.acc = torch.zeros(3, **f32)   # running sum over the steps of an epoch
        self.ws = torch.zeros(int(self.lib.nnmd_loss_workspace_bytes(B)), dtype=torch.uint8, device=self.device)
        self.graphs = None
        self.graph_key = None  # what the captured step reads its batch from (BPNN._step_plan)
        self.chunk_graphs = {}  # repeat count -> one graph holding that many consecutive steps
        # parameter / gradient pointer tables (the optimizer updates parameters in place: the pointers are stable)
        self._keep = []
        self.tables = []
        for net in self.nets:
            ws = [m.weight for m in net.model]
            bs = [m.bias for m in net.model]
            for p in ws + bs:
                if p.dtype != torch.float32 or not p.is_cuda or p.grad is None or not p.is_contiguous():
                    raise _lib.NnmdError("TrainStepPlan needs contiguous float32 CUDA parameters with FlatGradients views")
            sizes, c_sizes = _sizes(ws)
            self.tables.append((len(ws), c_sizes, _param_arrays(ws), _param_arrays(bs),
                                _param_arrays([p.grad for p in ws]), _param_arrays([p.grad for p in bs])))
        S = len(self.species)
        self.c_natoms = (ctypes.c_int32 * S)(*self.n_atoms)
        self.c_e, self.c_f = _param_arrays(self.e), _param_arrays(self.f)
        self.c_ge, self.c_gf = _param_arrays(self.ge), _param_arrays(self.gf)

    # ---- filling the static buffers
    def load(self, g, dG, energy, forces):
        for s in self.species:
            self.g[s].copy_(g[s].detach())
            self.dG[s].copy_(dG[s])
        self.E.copy_(energy.reshape(-1))
        self.F.copy_(forces)

    def load_indexed(self, base, sel, cursor=None):
        """Gather the frames `sel` (device int64) of a TrainAtomicDataset straight into the static buffers (one launch).
        cursor (device int64[1]): `sel` is the whole epoch's order and the batch is sel[cursor : cursor + n_frames]; the call
        advances the cursor on the device, so it can sit inside the captured step."""
        src, dst = [], []
        for s in self.species:
            gs, dgs = base.sf_data[s]
            src += [gs.detach(), dgs]
            dst += [self.g[s], self.dG[s]]
        src += [base.energies.reshape(-1), base.forces]
        dst += [self.E, self.F]
        for a, b in zip(src, dst):
            if a.dtype != b.dtype or not a.is_contiguous() or a.shape[1:] != b.shape[1:] or a.device != b.device:
                raise _lib.NnmdError("load_indexed: the cached tensors do not match the plan (float32, contiguous, same frame shape)")
        n = len(src)
        rows = (ctypes.c_int64 * n)(*[(a.numel() // max(a.shape[0], 1)) * a.element_size() for a in src])
        sel = sel.contiguous()
        with torch.cuda.device(self.device):
            if cursor is not None:
                _lib.check(self.lib.nnmd_gather_rows_cursor(n, _param_arrays(src), _param_arrays(dst), rows, _lib.ptr(sel),
                                                            _lib.ptr(cursor), self.n_frames, _lib.stream_ptr()))
            else:
                _lib.check(self.lib.nnmd_gather_rows(n, _param_arrays(src), _param_arrays(dst), rows, _lib.ptr(sel), sel.numel(),
                                                     _lib.stream_ptr()))

    def source_key(self, base):
        """Identity of the cached tensors a captured gather reads (a graph is only valid for the dataset it was captured on)."""
        return tuple(int(t.data_ptr()) for s in self.species for t in base.sf_data[s]) + (int(base.energies.data_ptr()), int(base.forces.data_ptr()))

    # ---- the step
    def run(self):
        lib, st, F32 = self.lib, _lib.stream_ptr(), _lib.dtype_code(torch.float32)
        self.flat.flat.zero_()
        with torch.cuda.device(self.device):
            for i, s in enumerate(self.species):
                L, c_sizes, c_w, c_b, _, _ = self.tables[i]
                rows = self.n_frames * self.n_atoms[i]
                _lib.check(lib.nnmd_mlp_energy_grad(F32, _lib.MLP_EXACT, L, c_sizes, c_w, c_b, _lib.ptr(self.g[s]), rows,
                                                    _lib.ptr(self.e[i]), _lib.ptr(self.dE[i]), st))
                _lib.check(lib.nnmd_force_contract(F32, _lib.ptr(self.dE[i]), _lib.ptr(self.dG[s]), rows, self.n_feat[i],
                                                   _lib.ptr(self.f[i]), st))
            _lib.check(lib.nnmd_loss_mse(F32, len(self.species), self.c_natoms, self.c_e, self.c_f, _lib.ptr(self.E),
                                         _lib.ptr(self.F), self.n_frames, self.e_coeff, self.f_coeff, self.c_ge, self.c_gf,
                                         _lib.ptr(self.ws), _lib.ptr(self.out), st))
            for i, s in enumerate(self.species):
                L, c_sizes, c_w, c_b, c_gw, c_gb = self.tables[i]
                rows = self.n_frames * self.n_atoms[i]
                _lib.check(lib.nnmd_force_contract_backward(F32, _lib.ptr(self.gf[i]), _lib.ptr(self.dG[s]), rows,
                                                            self.n_feat[i], _lib.ptr(self.u[i]), st))
                _lib.check(lib.nnmd_mlp_train_backward(F32, L, c_sizes, c_w, c_b, _lib.ptr(self.g[s]), rows,
                                                       _lib.ptr(self.ge[i]), _lib.ptr(self.u[i]), c_gw, c_gb, st))
        self.acc.add_(self.out)
