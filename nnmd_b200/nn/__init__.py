"""`nnmd.nn` surface of the reference (nnmd/nn/__init__.py:1-6) on the B200 path."""
from .atomic_nn import AtomicNN
from .bpnn import BPNN
from .dataset import TrainAtomicDataset
from .train_val_test_split import train_val_test_split

__all__ = ["AtomicNN", "BPNN", "TrainAtomicDataset", "train_val_test_split"]
