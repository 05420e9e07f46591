"""Training dataset with cached descriptors (reference: nnmd/nn/dataset.py:8-123).

`make_atomic_dataset` keeps the reference's contract -- per-species positions, float32, min-max scaled
energies, forces scaled by (emax - emin), descriptors per species ISOLATED from the other species,
`g_{spec}.pt` / `dg_{spec}.pt` caches in the working directory -- but all frames of a species go through one
pass of the CUDA descriptor path instead of the reference's 2-frame chunk loop.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .. import dist
from ..features import calculate_sf

_NOT_SPECIES = ("forces", "energy", "velocities")


def scale_targets(energies: torch.Tensor, forces: torch.Tensor, presharded: bool = False):
    """Min-max scaling of the targets (reference: dataset.py:85-90): E -> (E - emin) / (emax - emin), F -> F / (emax - emin).

    The extrema are GLOBAL over the training set.  `presharded`: the tensors hold only this rank's frames, so emin / emax
    are completed with one MIN and one MAX all-reduce; otherwise the caller passed every frame and no
    collective is needed.  Returns (energies, forces, emin, emax).
    """
    emin, emax = energies.min(), energies.max()
    if presharded and dist._world()[1] > 1:
        import torch.distributed as td
        emin, emax = emin.clone(), emax.clone()
        td.all_reduce(emin, op=td.ReduceOp.MIN)
        td.all_reduce(emax, op=td.ReduceOp.MAX)
    return (energies - emin) / (emax - emin), forces / (emax - emin), emin, emax


def cache_name(kind: str, species: str, rank: int, world: int) -> str:
    """`g_{spec}.pt` / `dg_{spec}.pt` as in the reference; one file per rank when the frame axis is sharded."""
    return f"{kind}_{species}.pt" if world == 1 else f"{kind}_{species}.rank{rank}of{world}.pt"


class TrainAtomicDataset(Dataset):
    def __init__(self, sf_data: dict, energies: torch.Tensor, forces: torch.Tensor, symm_func_params: dict) -> None:
        self.sf_data = sf_data              # {species: (g (M,N,F), dg (M,N,F,3))}
        self.energies = energies            # (M, 1)
        self.forces = forces                # (M, N_total, 3)
        self.symm_func_params = symm_func_params
        self.len = len(self.energies)

    def __getitem__(self, index):
        g = {spec: pair[0][index] for spec, pair in self.sf_data.items()}
        dG = {spec: pair[1][index] for spec, pair in self.sf_data.items()}
        return g, dG, self.energies[index], self.forces[index]

    def __len__(self):
        return self.len

    @classmethod
    def make_atomic_dataset(cls, dataset: dict, **kwargs) -> "TrainAtomicDataset":
        """dict from `input_parser` -> dataset of (g, dG, E, F) on the GPU (reference: dataset.py:33-123).

        kwargs: saved=True reloads g_/dg_ caches from the CWD; disable_tqdm accepted.  Extensions (the reference has none
        of them): cache=False skips writing the .pt files; shard=True (default when torch.distributed runs with
        world > 1) keeps only this rank's contiguous block of frames -- descriptors are computed for that block alone,
        targets are scaled with the extrema of ALL frames, caches are written per rank; presharded=True says the dict
        already holds only this rank's frames (extrema then come from an all-reduce); frames_per_pass=n bounds the
        descriptor workspace by sending n frames at a time through the CUDA path.
        """
        device = torch.device("cuda")
        frames = dataset["reference_data"]
        params = dataset["symmetry_functions_set"]
        f32 = dict(dtype=torch.float32, device=device)
        cell = torch.tensor(np.asarray(dataset["unit_cell"]), **f32)
        pbc = torch.tensor(np.asarray(dataset["pbc"]), **f32)
        species = [k for k in frames[0].keys() if k not in _NOT_SPECIES]
        rank, world = kwargs.get("rank"), kwargs.get("world")
        if rank is None or world is None:
            rank, world = dist._world()
        presharded = bool(kwargs.get("presharded", False))
        shard = bool(kwargs.get("shard", world > 1)) and not presharded

        energies = torch.tensor(np.array([fr["energy"] for fr in frames]), **f32)
        if energies.ndim == 1:
            energies = energies.unsqueeze(1)
        forces = torch.tensor(np.array([fr["forces"] for fr in frames]), **f32)
        energies, forces, _, _ = scale_targets(energies, forces, presharded)
        mine = dist.shard_frames(len(frames), rank, world) if shard else range(len(frames))
        if shard:
            energies, forces = energies[mine.start:mine.stop], forces[mine.start:mine.stop]
        cart = {s: torch.tensor(np.array([frames[i][s]["positions"] for i in mine]), **f32) for s in species}
        files_world = world if (shard or presharded) else 1

        sf_data = {}
        per_pass = int(kwargs.get("frames_per_pass", 0)) or len(mine)
        for s in species:
            if kwargs.get("saved"):
                g = torch.load(cache_name("g", s, rank, files_world), map_location=device)
                dg = torch.load(cache_name("dg", s, rank, files_world), map_location=device)
            else:
                parts = [calculate_sf(cart[s][lo:lo + per_pass], cell, pbc, params[s], disable_tqdm=kwargs.get("disable_tqdm", False))
                         for lo in range(0, max(len(mine), 1), max(per_pass, 1))]
                g = parts[0][0] if len(parts) == 1 else torch.cat([q[0] for q in parts], dim=0)
                dg = parts[0][1] if len(parts) == 1 else torch.cat([q[1] for q in parts], dim=0)
                if kwargs.get("cache", True):
                    torch.save(g, cache_name("g", s, rank, files_world))
                    torch.save(dg, cache_name("dg", s, rank, files_world))
            if not g.requires_grad:
                g.requires_grad = True
            sf_data[s] = (g, dg)

        print("Atomic dataset created with the following species:", ", ".join(sf_data.keys()))
        print(f"Total species in dataset: {len(sf_data)}")
        print(f"Shapes of g and dg tensors: {g.shape}, {dg.shape}")
        print(f"Shapes of energies and forces tensors: {energies.shape}, {forces.shape}")
        return cls(sf_data, energies, forces, params)
