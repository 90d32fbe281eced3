"""Dataset glue — drop-in for ``multimodalDataset`` / ``get_goldstein_params`` of the reference's
``VAESNe/data_util.py:10-20,76-79`` (the image datasets of that file are outside the accelerated path)."""
import re

import numpy as np
from torch.utils.data import Dataset

_SCI_FLOAT = re.compile(r"[-+]?\d*\.\d+e[-+]?\d+")     # e.g. "1.50e+00" inside a Goldstein model file name


class multimodalDataset(Dataset):
    """Zips equally long datasets; item i is the tuple of every modality's item i."""

    def __init__(self, *datasets):
        assert all(len(d) == len(datasets[0]) for d in datasets), "All datasets must be the same length"
        self.datasets = datasets
        self.num_modes = len(datasets)

    def __len__(self):
        return len(self.datasets[0])

    def __getitem__(self, idx):
        return tuple(d[idx] for d in self.datasets)


def get_goldstein_params(filename):
    """Physical parameters written in scientific notation inside a Goldstein model file name."""
    return np.array([float(tok) for tok in _SCI_FLOAT.findall(filename)])
