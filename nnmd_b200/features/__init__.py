"""`nnmd.features` surface of the reference (nnmd/features/__init__.py:13-23) on the B200 path."""
from .symm_funcs import (
    f_cutoff,
    g1_function,
    g2_function,
    g4_function,
    g5_function,
    calculate_distances,
    SymmetryFunction,
)
from .calculate_sf import calculate_sf
from .auto_params import calculate_params
from .autograd import symmetry_functions
from .multispecies import calculate_sf_species

__all__ = [
    "calculate_sf",
    "calculate_params",
    "SymmetryFunction",
    "f_cutoff",
    "g1_function",
    "g2_function",
    "g4_function",
    "g5_function",
    "calculate_distances",
    "symmetry_functions",
    "calculate_sf_species",
]
