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

from ..features import calculate_sf

_NOT_SPECIES = ("forces", "energy", "velocities")


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

        kwargs: saved=True reloads g_/dg_ caches from the CWD; disable_tqdm accepted; cache=False skips writing
        the .pt files (extension; the reference always writes them).
        """
        device = torch.device("cuda")
        frames = dataset["reference_data"]
        params = dataset["symmetry_functions_set"]
        f32 = dict(dtype=torch.float32, device=device)
        cell = torch.tensor(np.asarray(dataset["unit_cell"]), **f32)
        pbc = torch.tensor(np.asarray(dataset["pbc"]), **f32)
        species = [k for k in frames[0].keys() if k not in _NOT_SPECIES]
        cart = {s: torch.tensor(np.array([fr[s]["positions"] for fr in frames]), **f32) for s in species}
        energies = torch.tensor(np.array([fr["energy"] for fr in frames]), **f32)
        if energies.ndim == 1:
            energies = energies.unsqueeze(1)
        forces = torch.tensor(np.array([fr["forces"] for fr in frames]), **f32)

        emin, emax = energies.min(), energies.max()
        energies = (energies - emin) / (emax - emin)
        forces = forces / (emax - emin)

        sf_data = {}
        for s in species:
            if kwargs.get("saved"):
                g = torch.load(f"g_{s}.pt", map_location=device)
                dg = torch.load(f"dg_{s}.pt", map_location=device)
            else:
                g, dg = calculate_sf(cart[s], cell, pbc, params[s], disable_tqdm=kwargs.get("disable_tqdm", False))
                if kwargs.get("cache", True):
                    torch.save(g, f"g_{s}.pt")
                    torch.save(dg, f"dg_{s}.pt")
            if not g.requires_grad:
                g.requires_grad = True
            sf_data[s] = (g, dg)

        print("Atomic dataset created with the following species:", ", ".join(sf_data.keys()))
        print(f"Total species in dataset: {len(sf_data)}")
        print(f"Shapes of g and dg tensors: {g.shape}, {dg.shape}")
        print(f"Shapes of energies and forces tensors: {energies.shape}, {forces.shape}")
        return cls(sf_data, energies, forces, params)
