"""Behler-Parrinello network: one atomic MLP per species, trained on energies and forces
(reference: nnmd/nn/bpnn.py:45-565).  Public surface, config keys, optimizer settings, loss, log format and
checkpoint layout follow the reference; evaluation (`predict`) runs on the CUDA kernels K0..K6.

Deviations from the reference, all of them places where the reference raises as shipped:
  * `predict(cartesians, cell, symm_func_params, pbc=None)`: the reference forgets `pbc` and raises TypeError
    (bpnn.py:529-531).  Here pbc defaults to `self.pbc` when `config` received one, else all-False for a
    zero-determinant cell and all-True otherwise.
  * `fit` works without the DEBUG environment variable and without installed package metadata.
  * Two or more species: the training step concatenates species along the atom axis (the reference's dim-0
    concatenation raises a shape error there); descriptors stay per-species isolated like the reference's (Q9).
  * The training step does not build a torch graph through the MLP: forward (K6) and the double backward (K7) are
    CUDA kernels behind two autograd Functions; `retain_graph=True` and the `.grad` accumulation on the cached
    descriptors (SURVEY 3.2 notes ii, iii) have no counterpart.
"""
from __future__ import annotations

import os

import torch
from torch.utils.data import DataLoader

from .. import dist
from .._util._logger import Logger
from ..dist import FlatGradients
from ..engine import force_contract
from .functional import force_contract as autograd_force_contract
from ..features import calculate_sf
from .atomic_nn import AtomicNN
from .dataset import TrainAtomicDataset

torch.manual_seed(0)  # the reference seeds at import (bpnn.py:18); sample runs depend on it
DEBUG = os.environ.get("DEBUG", False)

if DEBUG:
    from tqdm import tqdm
else:
    def tqdm(iterable, **_kwargs):
        return iterable


class FrameBatches:
    """Iterable over shuffled frame batches of a TrainAtomicDataset (or a Subset of one): yields the same
    (g_dict, dG_dict, E, F) tuples a DataLoader with default collation would, built by tensor indexing."""

    def __init__(self, base: TrainAtomicDataset, indices: torch.Tensor, batch_size: int, shuffle: bool):
        self.base, self.indices, self.batch_size, self.shuffle = base, indices, int(batch_size), shuffle

    def __len__(self):
        return (self.indices.numel() + self.batch_size - 1) // self.batch_size

    def _epoch_order(self) -> torch.Tensor:
        """Frame order of one epoch.  Consumes the global RNG exactly as `iter(DataLoader(ds, batch_size, shuffle=...))`
        does -- one int64 draw for the iterator's base seed, and with shuffle one more that seeds a private generator
        for the permutation (torch.utils.data: _BaseDataLoaderIter.__init__, RandomSampler.__iter__) -- so a sample
        script sees the same batches, in the same order, as with the reference's DataLoader (bpnn.py:234-236)."""
        idx = self.indices
        torch.empty((), dtype=torch.int64).random_()
        if self.shuffle and idx.numel() > 0:
            gen = torch.Generator()
            gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
            idx = idx[torch.randperm(idx.numel(), generator=gen)]
        return idx

    def index_batches(self, device):
        """The frame indices of every batch of one epoch as device tensors (one host-to-device copy per epoch)."""
        idx = self._epoch_order().to(device)
        for s in range(0, idx.numel(), self.batch_size):
            yield idx[s:s + self.batch_size]

    def __iter__(self):
        idx = self._epoch_order()
        for s in range(0, idx.numel(), self.batch_size):
            sel = idx[s:s + self.batch_size]
            g, dG = {}, {}
            for spec, (gs, dgs) in self.base.sf_data.items():
                sel_d = sel.to(gs.device)
                g[spec] = gs.detach().index_select(0, sel_d)
                dG[spec] = dgs.index_select(0, sel_d)
            yield g, dG, self.base.energies.index_select(0, sel.to(self.base.energies.device)), \
                self.base.forces.index_select(0, sel.to(self.base.forces.device))


def batch_loader(dataset, batch_size: int, shuffle: bool):
    """FrameBatches for a TrainAtomicDataset / Subset chain, else a stock DataLoader."""
    indices = None
    base = dataset
    while isinstance(base, torch.utils.data.Subset):
        sub = torch.as_tensor(base.indices, dtype=torch.int64)
        indices = sub if indices is None else sub[indices]
        base = base.dataset
    if isinstance(base, TrainAtomicDataset):
        if indices is None:
            indices = torch.arange(len(base), dtype=torch.int64)
        return FrameBatches(base, indices, batch_size, shuffle)
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle)


class RMSELoss(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.mse = torch.nn.MSELoss()

    def forward(self, yhat, y):
        return torch.sqrt(self.mse(yhat, y))


available_optimizers = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW}
available_losses = {"mse": torch.nn.MSELoss, "mae": torch.nn.L1Loss, "rmse": RMSELoss}


def _short_lr(lr: float) -> str:
    import numpy as np
    return str(np.float32(lr))


def _package_version() -> str:
    try:
        from importlib.metadata import version
        return version("nnmd")
    except Exception:
        return "0.2.0.dev0+b200"


class BPNN(torch.nn.Module):
    """High-dimensional NN potential: per-species AtomicNN, shared optimizer (reference: bpnn.py:45-62)."""

    def __init__(self, dtype=torch.float32, use_cuda: bool = True) -> None:
        super().__init__()
        self.dtype = dtype
        self.use_cuda = use_cuda
        self.device = torch.device("cuda") if self.use_cuda else torch.device("cpu")
        self.pbc = None

    # ------------------------------------------------------------------------------------------ configuration
    def config(self, neural_net_data: dict):
        """reference: bpnn.py:106-194 (Xavier-uniform weights, zero biases, Adam(0.8,0.9999)/AdamW(0.9,0.999) + amsgrad)."""
        self.neural_net_data = neural_net_data
        self.species = neural_net_data["atom_species"]
        self.input_sizes = neural_net_data["input_sizes"]
        self.hidden_sizes = neural_net_data["hidden_sizes"]
        self.learning_rate = neural_net_data["learning_rate"]
        self.l2_regularization = neural_net_data["l2_regularization"]
        self.e_loss_coeff = neural_net_data["e_loss_coeff"]
        self.f_loss_coeff = neural_net_data["f_loss_coeff"]
        if "pbc" in neural_net_data:
            self.pbc = neural_net_data["pbc"]

        self.atomic_nets = []
        self._flat_grads = None
        for i, spec in enumerate(self.species):
            net = AtomicNN(input_size=self.input_sizes[i], hidden_sizes=self.hidden_sizes[i], output_size=1)
            net = net.to(self.device)
            if neural_net_data["load_models"]:
                net.load_state_dict(torch.load(f"{neural_net_data['path']}/atomic_nn_{spec}.pt", map_location=self.device))
            else:
                for layer in net.model:
                    if isinstance(layer, torch.nn.Linear):
                        torch.nn.init.xavier_uniform_(layer.weight)
                        torch.nn.init.constant_(layer.bias, 0)
            self.atomic_nets.append(net)

        self.criterion = torch.nn.MSELoss()

        params = [p for net in self.atomic_nets for p in net.parameters()]
        name = neural_net_data["optimizer"]
        if name is None:
            name = "adam"
            print("Optimizer is not specified, using default: Adam")
        elif name not in available_optimizers:
            raise ValueError(f"Optimizer {name} is not available. "
                             f"Available optimizers: {list(available_optimizers.keys())}")
        betas = (0.9, 0.999) if name == "adamw" else (0.8, 0.9999)
        kwargs = dict(params=params, lr=self.learning_rate, weight_decay=self.l2_regularization, betas=betas, amsgrad=True)
        on_gpu = self.device.type == "cuda"
        try:  # same update rule; fused = one kernel per step, capturable = step counter and lr live on the device so
            #   the whole training step can be replayed from a CUDA graph (the step is launch-bound otherwise)
            if on_gpu:  # a device-resident lr is what a captured optimizer step reads; schedulers update it in place
                kwargs["lr"] = torch.tensor(float(self.learning_rate), dtype=torch.float32, device=self.device)
            self.optim = available_optimizers[name](fused=on_gpu, capturable=on_gpu, **kwargs)
            self._graphs_ok = on_gpu and os.environ.get("NNMD_NO_CUDA_GRAPH", "0") != "1"
        except (RuntimeError, TypeError, ValueError):
            kwargs["lr"] = self.learning_rate
            self.optim = available_optimizers[name](**kwargs)
            self._graphs_ok = False
        self._plans, self._opt_ready, self._acc_carry = {}, False, None

    # ------------------------------------------------------------------------------------------------- loss
    def _loss(self, e_nn, energies, f_nn, forces):
        """e_coeff * MSE(E) + f_coeff / 3 * MSE(F)   (reference: bpnn.py:196-215)."""
        E_loss = self.e_loss_coeff * self.criterion(e_nn, energies)
        F_loss = self.f_loss_coeff / 3 * self.criterion(f_nn, forces)
        return E_loss + F_loss, E_loss, F_loss

    def _energies_forces(self, g: dict, dG: dict):
        """Per-species MLP, dE/dg and the force contraction (bpnn.py:342-364) on K6 / K4c; the backward of the loss
        through both (the reference's double backward) is K7 / K4c' (nn/functional.py)."""
        e_nn, f_nn = [], []
        for i, spec in enumerate(self.species):
            e_spec, dE = self.atomic_nets[i].energy_and_grad_autograd(g[spec])
            e_nn.append(e_spec)
            f_nn.append(autograd_force_contract(dE, dG[spec]))
        # The reference concatenates the species along dim 0 (bpnn.py:357-358), which for batched (B, N_s, .) tensors
        # stacks species on the FRAME axis and fails against (B,1) energies as soon as there are two species
        # (SURVEY Q9).  One species: identical.  Several species: atoms are concatenated along the atom axis, so that
        # e.sum(1) is the frame energy and f matches the (B, N_total, 3) reference forces in species order.
        cat_dim = 0 if len(e_nn) == 1 else 1
        return torch.cat(e_nn, dim=cat_dim), torch.cat(f_nn, dim=cat_dim)

    # ------------------------------------------------------------------------------------------------ loops
    def _train_step_eager(self, g, dG, energy, forces, out):
        """One optimisation step: K6 forward, loss, K7/K4c' backward, gradient all-reduce, optimizer."""
        self._flat_grads.zero()
        e_nn, f_nn = self._energies_forces(g, dG)
        loss, e_loss, f_loss = self._loss(e_nn.sum(dim=1), energy, f_nn, forces)
        loss.backward()
        self._flat_grads.allreduce(int(energy.shape[0]))  # frames are sharded across ranks (no-op on 1 rank)
        self.optim.step()
        out.copy_(torch.stack([e_loss.detach(), f_loss.detach(), loss.detach()]))

    def _train_step_empty(self):
        """This rank has no batch left in a step the other ranks still take: zero gradient with weight 0, the SAME
        collective and optimizer step as everybody else (the weights stay identical on all ranks, nobody hangs)."""
        self._flat_grads.zero()
        self._flat_grads.allreduce(0)
        self.optim.step()

    def _epoch_steps(self, n_local: int, loader=None) -> int:
        """Steps per epoch = the longest rank's batch count: per-rank frame shards can differ by one frame, hence by one
        batch, and every step issues a collective.  One scalar all-reduce (and one host sync) per LOADER, not per epoch: the
        answer is remembered per loader object -- every rank runs the same program, so the same loader is met again on all
        ranks at once (a weak dictionary: a freed loader cannot alias a new one)."""
        if dist._world()[1] == 1:
            return n_local
        import weakref
        import torch.distributed as td
        cache = getattr(self, "_steps_cache", None)
        if cache is None:
            cache = self._steps_cache = weakref.WeakKeyDictionary()
        try:
            hit = cache.get(loader) if loader is not None else None
        except TypeError:  # not weak-referenceable
            hit, loader = None, None
        if hit is not None and hit[0] == n_local:
            return hit[1]
        t = torch.tensor([n_local], dtype=torch.int64, device=self.device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        n = int(t.item())
        if loader is not None:
            cache[loader] = (n_local, n)
        return n

    def _upload_order(self, order: torch.Tensor) -> None:
        """The epoch's frame order -> self._perm without a host sync: two pinned staging buffers used in turn, each guarded
        by an event (the host may be a whole epoch ahead of the device)."""
        n = order.numel()
        st = getattr(self, "_perm_stage", None)
        if st is None or st[0][0].numel() < n:
            cap = max(n, self._perm.numel())
            st = self._perm_stage = [[torch.empty(cap, dtype=torch.int64).pin_memory(), None] for _ in range(2)]
            self._perm_turn = 0
        buf, ev = st[self._perm_turn]
        if ev is not None:
            ev.synchronize()
        buf[:n].copy_(order)
        self._perm[:n].copy_(buf[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st[self._perm_turn][1] = ev
        self._perm_turn ^= 1

    # ---- fused step: K6 -> K4c -> K8 -> K4c' -> K7 on static buffers, replayed from CUDA graphs
    def _fused_ok(self, g, dG) -> bool:
        if self.device.type != "cuda" or len(self.species) > 8 or os.environ.get("NNMD_NO_FUSED_STEP", "0") == "1":
            return False
        if not isinstance(self.criterion, torch.nn.MSELoss) or self.criterion.reduction != "mean":
            return False
        ok = all(net.activation is torch.sigmoid for net in self.atomic_nets)
        ok = ok and all(p.dtype == torch.float32 and p.is_cuda for net in self.atomic_nets for p in net.parameters())
        return ok and all(g[s].dtype == torch.float32 and g[s].is_cuda and g[s].dim() == 3 and dG[s].dtype == torch.float32
                          for s in self.species)

    def _plan_for(self, shapes: dict, n_frames: int):
        from .functional import TrainStepPlan
        key = (n_frames,) + tuple((s, tuple(shapes[s])) for s in self.species)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:  # full batches plus the ragged last one; anything else is churn
                old = self._plans.pop(next(iter(self._plans)))
                self._acc_carry = old.acc.clone() if self._acc_carry is None else self._acc_carry + old.acc
            plan = TrainStepPlan(self.atomic_nets, self.species, shapes, n_frames, self.device, self.e_loss_coeff,
                                 self.f_loss_coeff, self._flat_grads)
            self._plans[key] = plan
        return plan

    def _step_plan(self, plan, feed=None, repeat: int = 1):
        """Run one plan (`repeat` > 1: that many consecutive steps of it -- needs a feed).  feed=None: its buffers were loaded by the caller.  feed=(dataset, perm, cursor): the batch is
        perm[cursor : cursor + n_frames] of the dataset, gathered on the device and the cursor advanced there -- the gather
        is part of the captured step, so a batch costs the host one graph replay and nothing else.

        One rank: the whole step is one CUDA graph.  Several ranks with a peer window (dist.PeerWindow: the gradient
        all-reduce is a kernel over NVLink peer memory): still ONE graph -- gather, forward, backward, all-reduce, optimizer,
        loss accounting.  Several ranks on NCCL: gather + forward + backward (graph 1) | all-reduce | optimizer (graph 2);
        capturing the NCCL all-reduce hung on this stack (torch 2.11 / NCCL 2.28, 2 GPUs).

        repeat > 1 (one-graph case only): ONE graph holds `repeat` whole steps -- a 0.12 ms step is launch-bound on the host
        as soon as several ranks share the box's cores and wait for each other in every all-reduce; a chunk of steps per
        replay takes the host out of the loop."""
        multi = self._flat_grads.world_size() > 1
        key = None if feed is None else (plan.source_key(feed[0]), int(feed[1].data_ptr()), int(feed[2].data_ptr()))
        if plan.graphs is not None and plan.graph_key != key:
            plan.graphs, plan.chunk_graphs = None, {}  # captured for another feed (dataset, permutation buffer): capture again
        if repeat > 1:
            if plan.graphs is None or len(plan.graphs) != 1 or feed is None:
                for _ in range(repeat):  # no single-graph step (yet): one at a time
                    self._step_plan(plan, feed)
                return
            g = plan.chunk_graphs.get(repeat)
            if g is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=plan.graphs[0].pool()):
                    for _ in range(repeat):
                        plan.load_indexed(feed[0], feed[1], cursor=feed[2])
                        plan.run()
                        if multi and not self._flat_grads.allreduce(plan.n_frames, plan.out, self._acc_multi):
                            self._acc_multi.addcdiv_(plan.out, self._flat_grads.count, value=float(plan.n_frames))
                        self.optim.step()
                plan.chunk_graphs[repeat] = g
            g.replay()
            return

        def load():
            if feed is not None:
                plan.load_indexed(feed[0], feed[1], cursor=feed[2])

        if self._graphs_ok and plan.graphs is None and self._opt_ready:
            try:  # the optimizer state exists (one eager step was done): capture the step
                torch.cuda.synchronize()
                if not multi or self._flat_grads.graph_capturable():
                    whole = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(whole):
                        load()
                        plan.run()
                        # this rank's share of the global-batch loss, out * b_r / sum_r b_r, is added by the all-reduce kernel
                        if multi and not self._flat_grads.allreduce(plan.n_frames, plan.out, self._acc_multi):
                            self._acc_multi.addcdiv_(plan.out, self._flat_grads.count, value=float(plan.n_frames))
                        self.optim.step()
                    plan.graphs = (whole,)
                else:
                    fwd_bwd, opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(fwd_bwd):
                        load()
                        plan.run()
                    with torch.cuda.graph(opt, pool=fwd_bwd.pool()):
                        self.optim.step()
                    plan.graphs = (fwd_bwd, opt)
                plan.graph_key = key
                # a capture does not execute: fall through to the replay below
            except Exception as exc:  # capture not possible here: stay eager, loudly
                import sys
                print(f"nnmd_b200: CUDA-graph capture of the training step failed ({exc!r}); continuing without graphs",
                      file=sys.stderr, flush=True)
                self._graphs_ok = False
        if plan.graphs is not None and len(plan.graphs) == 1:
            plan.graphs[0].replay()
            return
        if plan.graphs is not None:
            plan.graphs[0].replay()
            self._flat_grads.allreduce(plan.n_frames)  # frames are sharded across ranks: weighted by the batch sizes
            plan.graphs[1].replay()
        else:
            load()
            plan.run()
            self._flat_grads.allreduce(plan.n_frames)
            self.optim.step()
            self._opt_ready = True
        if multi:
            self._acc_multi.addcdiv_(plan.out, self._flat_grads.count, value=float(plan.n_frames))

    def _train_step(self, g, dG, energy, forces, out):
        """One optimisation step on a batch given as tensors (DataLoader-style tuples)."""
        if not self._fused_ok(g, dG):
            return self._train_step_eager(g, dG, energy, forces, out)
        plan = self._plan_for({s: g[s].shape for s in self.species}, int(energy.shape[0]))
        plan.load(g, dG, energy, forces)
        self._step_plan(plan)
        out.copy_(plan.out)

    def _run_epoch(self, loader, train: bool, desc: str):
        """One pass over `loader`.  One rank: epoch loss = mean of the batch losses (bpnn.py:368-377).  Frames sharded over
        several ranks: every rank takes the same number of steps (`_epoch_steps`), the gradient of a step is that of the
        GLOBAL batch (`FlatGradients.allreduce`), and the logged losses are global too (training: mean over steps of the
        global-batch loss; validation / test: frame-weighted mean), identical on every rank."""
        for net in self.atomic_nets:
            net.train(train)
        world = dist._world()[1]
        tot = torch.zeros(3, device=self.device)
        out = torch.zeros(3, device=self.device)
        if train and getattr(self, "_flat_grads", None) is None:
            # gradients live in one flat buffer: a single in-place all-reduce per step when frames are sharded
            self._flat_grads = FlatGradients([p for net in self.atomic_nets for p in net.parameters()])
            self._plans, self._opt_ready, self._acc_carry = {}, False, None
        n_steps = self._epoch_steps(len(loader), loader)
        if getattr(self, "_acc_multi", None) is None:  # persistent: the captured step adds into it
            self._acc_multi = torch.zeros(3, device=self.device)
        self._acc_multi.zero_()
        if train and isinstance(loader, FrameBatches) and loader.base.energies.is_cuda \
                and self._fused_ok({s: v[0] for s, v in loader.base.sf_data.items()}, {s: v[1] for s, v in loader.base.sf_data.items()}):
            # batches are gathered straight into the static buffers of the step (no per-batch tensors, no copies)
            shapes = {s: loader.base.sf_data[s][0].shape for s in self.species}
            self._acc_carry = None  # loss sums of plans evicted during the epoch
            for plan in self._plans.values():
                plan.acc.zero_()
            # the epoch's frame order goes to the device once; every step then reads its batch through a device cursor
            order = loader._epoch_order()
            if getattr(self, "_perm", None) is None or self._perm.numel() < order.numel():
                # sized for the whole dataset: a longer epoch later (train after a short warm-up subset) must not move
                # the buffer, or every captured step would have to be captured again
                self._perm = torch.empty(max(order.numel(), len(loader.base), 1), dtype=torch.int64, device=self.device)
                self._cursor = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._upload_order(order)
            self._cursor.zero_()
            feed = (loader.base, self._perm, self._cursor)
            left = order.numel()
            chunk = max(int(os.environ.get("NNMD_STEP_CHUNK", "8")), 1)  # full-batch steps per graph replay
            bar = iter(tqdm(range(n_steps), desc=desc, total=n_steps))  # progress bar only with DEBUG set (as the reference)
            k = 0
            while k < n_steps:
                if left <= 0:
                    self._train_step_empty()
                    k += 1
                    next(bar, None)
                    continue
                nb = min(loader.batch_size, left)
                plan = self._plan_for(shapes, nb)
                rep = chunk if (chunk > 1 and left >= chunk * loader.batch_size and plan.graphs is not None
                                and len(plan.graphs) == 1) else 1
                self._step_plan(plan, feed, repeat=rep)
                left -= nb * rep
                k += rep
                for _ in range(rep):
                    next(bar, None)
            if world > 1:
                tot = dist.allreduce_sum(self._acc_multi.clone()) / n_steps  # peer-window kernel where available: no host sync
                return tot[0], tot[1], tot[2]
            for plan in self._plans.values():
                tot += plan.acc
                plan.acc.zero_()
            if self._acc_carry is not None:
                tot += self._acc_carry
            tot /= len(loader)
            return tot[0], tot[1], tot[2]
        batches = iter(loader)
        frames = torch.zeros(1, device=self.device)
        for _ in tqdm(range(n_steps), desc=desc, total=n_steps):
            item = next(batches, None)
            if item is None:
                if train:
                    self._train_step_empty()
                continue
            g, dG, energy, forces = item
            energy = energy.to(device=self.device)
            forces = forces.to(device=self.device)
            if train:
                self._train_step(g, dG, energy, forces, out)
                if world > 1 and not self._fused_ok(g, dG):  # the fused step adds its share itself (`_step_plan`)
                    self._acc_multi.addcdiv_(out, self._flat_grads.count, value=float(energy.shape[0]))
            else:
                with torch.no_grad():
                    e_nn, f_nn = self._energies_forces(g, dG)
                    loss, e_loss, f_loss = self._loss(e_nn.sum(dim=1), energy, f_nn, forces)
                out.copy_(torch.stack([e_loss, f_loss, loss]))
                self._acc_multi += out * float(energy.shape[0])
                frames += float(energy.shape[0])
            tot += out
        if world > 1:
            import torch.distributed as td
            if train:
                td.all_reduce(self._acc_multi)
                tot = self._acc_multi / n_steps
            else:
                both = torch.cat([self._acc_multi, frames])
                td.all_reduce(both)
                tot = both[:3] / both[3]
            return tot[0], tot[1], tot[2]
        tot /= len(loader)
        return tot[0], tot[1], tot[2]

    def _train_loop(self, epoch: int, train_loader: DataLoader):
        """reference: bpnn.py:308-377; returns epoch means (E loss, F loss, total)."""
        return self._run_epoch(train_loader, True, f"Epoch {epoch + 1}, Training")

    def _validate(self, epoch: int, val_loader: DataLoader):
        """reference: bpnn.py:379-444."""
        return self._run_epoch(val_loader, False, f"Epoch {epoch + 1}, Validation")

    def _test(self, test_loader: DataLoader):
        """reference: bpnn.py:446-505."""
        return self._run_epoch(test_loader, False, "Testing")

    def _describe_training(self, n_train, n_val, n_test, batch_size, epochs):
        """Log header, same lines as the reference (bpnn.py:64-104; samples/*/plot_errors.py parse the log)."""
        log = self.net_log.info
        log(f"NNMD v{_package_version()}\n")
        log("---Neural network parameters---")
        log(f"device:                             {self.device}")
        log(f"training epochs:                    {epochs}")
        log(f"batch size:                         {batch_size}")
        log(f"optimizer:                          {self.optim.__class__.__name__}")
        log(f"loss function:                      {self.criterion.__class__.__name__}")
        log(f"energies loss function coefficient: {self.e_loss_coeff}")
        log(f"forces loss function coefficient:   {self.f_loss_coeff}")
        log(f"learning rate:                      {self.learning_rate}")
        log(f"scheduler:                          {self.sched.__class__.__name__}")
        log(f"L2 regularization coefficient:      {self.l2_regularization}\n\n")
        log("----Atomic NNs parameters----")
        log(f"input sizes (number of descriptors): {self.input_sizes}")
        log(f"hidden sizes:                        {self.hidden_sizes}\n\n")
        log("---Dataset parameters---")
        log(f"Training sample size:   {n_train}")
        log(f"Validation sample size: {n_val}")
        log(f"Test sample size:       {n_test}\n\n")

    def fit(self, train_dataset: TrainAtomicDataset, val_dataset: TrainAtomicDataset, test_dataset: TrainAtomicDataset):
        """reference: bpnn.py:217-306 (ExponentialLR gamma 0.99, net.log, per-epoch train/validation lines)."""
        batch_size = self.neural_net_data["batch_size"]
        epochs = self.neural_net_data["epochs"]
        # same batching as DataLoader(dataset, batch_size, shuffle=True) (bpnn.py:234-236), but a batch is ONE
        # index_select per cached tensor instead of batch_size per-frame lookups and a Python collate
        train_loader = batch_loader(train_dataset, batch_size, shuffle=True)
        val_loader = batch_loader(val_dataset, batch_size, shuffle=True)
        test_loader = batch_loader(test_dataset, batch_size, shuffle=True)
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.optim, gamma=0.99)
        if dist._world()[0] == 0:
            self.net_log = Logger.get_logger("Net training info", "net.log")
        else:  # the losses are global (identical on every rank): one writer
            import logging
            self.net_log = logging.getLogger("nnmd_b200.null")
            self.net_log.addHandler(logging.NullHandler())
            self.net_log.propagate = False
        self._describe_training(len(train_dataset), len(val_dataset), len(test_dataset), batch_size, epochs)
        curr_lr = float(self.optim.param_groups[0]["lr"])
        line = "{loss} E = {e_loss:e}, {loss} F = {f_loss:e}, {loss} total = {total:e} eV"
        name = self.criterion.__class__.__name__
        for epoch in range(epochs):
            e_loss, f_loss, loss = self._train_loop(epoch, train_loader)
            fg = getattr(self, "_flat_grads", None)
            if fg is not None and fg.window is not None:
                fg.window.check()  # a peer that never arrived at an all-reduce (the .item() below synchronises anyway)
            self.net_log.info(f"Epoch {epoch + 1}: training:   "
                              + line.format(e_loss=e_loss.item(), f_loss=f_loss.item(), loss=name, total=loss.item()))
            e_loss, f_loss, loss = self._validate(epoch, val_loader)
            self.net_log.info(f"Epoch {epoch + 1}: validation: "
                              + line.format(e_loss=e_loss.item(), f_loss=f_loss.item(), loss=name, total=loss.item()))
            self.sched.step()
            if float(self.optim.param_groups[0]["lr"]) != curr_lr:
                curr_lr = float(self.optim.param_groups[0]["lr"])
                # the captured optimizer keeps lr as a float32 device scalar: print its shortest round-trip form
                # (0.099, like the reference's Python float, not 0.0989999994635582)
                self.net_log.info(f"Learning rate changed to {_short_lr(curr_lr)}")
        e_loss, f_loss, loss = self._test(test_loader)
        self.net_log.info("Testing: " + line.format(e_loss=e_loss.item(), f_loss=f_loss.item(), loss=loss.item(),
                                                    total=loss.item()))

    # ---------------------------------------------------------------------------------------------- predict
    def _default_pbc(self, cell: torch.Tensor) -> torch.Tensor:
        if self.pbc is not None:
            return torch.as_tensor(self.pbc, dtype=self.dtype)
        has_cell = bool(torch.det(cell.detach().to("cpu", torch.float64)) != 0)
        return torch.full((3,), 1.0 if has_cell else 0.0, dtype=self.dtype)

    def predict(self, cartesians: dict, cell: torch.Tensor, symm_func_params: dict, pbc=None):
        """Energies and forces for frames of atoms (reference: bpnn.py:507-554), on K0..K6.

        cartesians: {species: (B,N,3) or (N,3)}; returns (E per frame squeezed, F squeezed) like the reference:
        species are concatenated along dim 0 exactly as bpnn.py:546-552 does.

        Repeated calls with the same system size (MD, scans) reuse a per-species `NeighborWorkspace`: the neighbour build
        then runs in one pass without a host synchronisation, and for fp32 nets with a tensor-core instantiation the
        normalisation, the network and the force contraction are ONE kernel.  The call synchronises once, at its end, to
        raise the reference's ValueError on NaN descriptors (calculate_sf.py:106-109).  `release_workspaces()` frees the
        cached buffers (the edge scratch of a million-atom structure is ~19 GB).
        """
        from ..engine import energy_forces
        from ..features.calculate_sf import get_plan, raise_on_flag
        from ..neighbor import NeighborWorkspace
        for net in self.atomic_nets:
            net.eval()
        torch.manual_seed(42)  # side effect of the reference's calculate_sf call inside predict (calculate_sf.py:115)
        pbc_t = self._default_pbc(cell) if pbc is None else torch.as_tensor(pbc)
        cache = self.__dict__.setdefault("_predict_ws", {})
        e_nn = f_nn = None
        for i, spec in enumerate(self.species):
            cart = cartesians[spec].to(device=self.device)
            if cart.dim() == 2:
                cart = cart.unsqueeze(0)
            cartesians[spec] = cart
            B, N, _ = cart.shape
            plan = get_plan(symm_func_params[spec], cart.dtype, cart.device)
            key = (spec, B, N, cart.dtype, plan.rc_max)
            for attempt in range(3):
                ws = cache.get(key)
                if ws is None:
                    if len(cache) >= 4:  # a handful of shapes at most: drop the oldest (its buffers go back to the allocator)
                        cache.pop(next(iter(cache)))
                    ws = cache[key] = NeighborWorkspace(B, N, plan.rc_max, cart.dtype, cart.device)
                ws.status.zero_()  # the status word is sticky across steps: a predict call starts clean
                res = energy_forces(cart, cell, pbc_t.to(cart.dtype), symm_func_params[spec], self.atomic_nets[i],
                                    mode="materialize", check=False, workspace=ws)
                if ws.validate():  # one host sync per species; False: the structure outgrew the cached buffers (regrown)
                    break
            raise_on_flag(res.flag)
            e_spec, f_spec = res.e_atoms.unsqueeze(-1), res.forces
            e_spec, f_spec = e_spec.squeeze(0), f_spec.squeeze(0)  # the reference squeezes g/dg before the net
            e_nn = e_spec if e_nn is None else torch.cat((e_nn, e_spec), dim=0)
            f_nn = f_spec if f_nn is None else torch.cat((f_nn, f_spec), dim=0)
        return e_nn.sum(1).squeeze(), f_nn.squeeze()

    def release_workspaces(self) -> None:
        """Free the neighbour-list workspaces `predict` keeps between calls."""
        self.__dict__.pop("_predict_ws", None)

    def save_model(self, path: str):
        """One state_dict per species at <path>/atomic_nn_<spec>.pt (reference: bpnn.py:556-565)."""
        for spec, net in zip(self.neural_net_data["atom_species"], self.atomic_nets):
            torch.save(net.state_dict(), f"{path}/atomic_nn_{spec}.pt")
