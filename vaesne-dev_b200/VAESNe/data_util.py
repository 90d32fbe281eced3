"""Dataset glue — drop-in for ``multimodalDataset`` / ``get_goldstein_params`` of the reference's
``VAESNe/data_util.py:10-20,76-79`` (the image datasets of that file are outside the accelerated path)."""
import re

import numpy as np
import torch
from torch.utils.data import Dataset, TensorDataset

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


class ResidentLoader:
    """Batches served from device memory: the replacement for ``DataLoader(dataset, batch_size, shuffle=True)``
    (`cannon/ZTF_photospect.py:71-77`) when the whole (augmented) training set fits in HBM — it always does: 10x-augmented
    ZTF is tens of MB against 180 GB.

    ``dataset`` is a ``TensorDataset`` (single modality) or a ``multimodalDataset`` of ``TensorDataset``s.  One epoch is a
    fresh permutation (``torch.randperm`` on the device, seeded like ``DataLoader``'s sampler from torch's global RNG unless a
    generator is given); a batch is one ``index_select`` per tensor — no per-sample Python, no collate, no host-to-device
    copies, so a large-global-batch data-parallel run is not host-bound.  Yields what ``training_step`` consumes: a tuple of
    tensors, or (multimodal) a list with one tuple per modality.  ``rank``/``world`` deal the permutation out to data-parallel
    ranks (every rank draws the same permutation: pass generators seeded alike, or seed torch's global RNG alike)."""

    def __init__(self, dataset, batch_size, shuffle=True, device=None, drop_last=False, generator=None, rank=0, world=1):
        self.multimodal = isinstance(dataset, multimodalDataset)
        parts = dataset.datasets if self.multimodal else (dataset,)
        if not all(isinstance(d, TensorDataset) for d in parts):
            raise TypeError("ResidentLoader needs TensorDataset modalities (tensors that can live on the device)")
        self.device = torch.device(device) if device is not None else parts[0].tensors[0].device
        self.mods = [tuple(t.to(self.device) for t in d.tensors) for d in parts]
        self.n = len(parts[0])
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator, self.rank, self.world = generator, int(rank), int(world)

    def _local_count(self):
        # every rank serves the SAME number of samples (the n % world leftovers of an epoch's permutation are dropped): ranks
        # with different step counts, or different last-batch sizes, would dead-lock in / skew the gradient all-reduce
        return self.n // self.world

    def __len__(self):
        m = self._local_count()
        return m // self.batch_size if self.drop_last else (m + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_(generator=self.generator).item())
            g = torch.Generator(device=self.device).manual_seed(seed)
            perm = torch.randperm(self.n, device=self.device, generator=g)
        else:
            perm = torch.arange(self.n, device=self.device)
        perm = perm[:(self.n // self.world) * self.world][self.rank::self.world]
        for i in range(len(self)):
            idx = perm[i * self.batch_size:(i + 1) * self.batch_size]
            batch = [tuple(t.index_select(0, idx) for t in mod) for mod in self.mods]
            yield batch if self.multimodal else batch[0]


def _image_dataset_stub(name):
    class _Stub(Dataset):
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"VAESNe-B200: data_util.{name} (host-galaxy image datasets, reference data_util.py:23-73) is outside "
                                      "the accelerated photometry/spectra path and is not provided")
    _Stub.__name__ = name
    return _Stub


ImagePathDataset = _image_dataset_stub("ImagePathDataset")
ImagePathDatasetAug = _image_dataset_stub("ImagePathDatasetAug")
