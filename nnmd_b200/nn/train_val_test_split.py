"""80/10/10 random split (reference: nnmd/nn/train_val_test_split.py:4-25) -- stock torch, not on the hot path."""
from torch.utils.data import random_split


def train_val_test_split(dataset, train_val_test_ratio: tuple = (0.8, 0.1, 0.1)):
    n = len(dataset)
    n_train = int(n * train_val_test_ratio[0])
    n_val = int(n * train_val_test_ratio[1])
    return random_split(dataset, [n_train, n_val, n - n_train - n_val])
