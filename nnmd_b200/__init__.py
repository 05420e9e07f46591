"""nnmd_b200 -- B200-native (sm_100a) implementation of nnmd's descriptor -> energy -> force hot path.

Sub-packages mirror the reference's public surface: `features` (calculate_sf & co) and `nn` (AtomicNN, BPNN,
TrainAtomicDataset).  The CUDA kernels live in csrc/ behind the C ABI of include/nnmd_b200.h; there is no CPU
implementation of the hot path in this package.
"""
__version__ = "0.1.0"

from . import features, nn  # noqa: F401,E402
