"""The step BEFORE the hot path: the on-disk contract of the training sets and their augmentation, on the device.

Every reference training script does the same thing on the host before building its DataLoader
(cannon/test_photospectra.py:22-78, cannon/ZTF_photospect.py:20-66): read one ``.npz`` (keys ``flux, wavelength, mask, phase,
photoflux, phototime, photomask, photowavelength`` [+ ``training_idx, testing_idx``]), turn ``mask == 0`` into the boolean
"unobserved" mask, optionally repeat the set 10x, add Gaussian noise to fluxes / times / phases, shift each light curve in time
by one draw, and OR a random mask on top.  Here the arrays go to the device once and the augmentation is ONE kernel per
tensor (``vaesne_augment``: repeat + per-element noise + per-row shift + random masking, counter-based generator), so a
large-global-batch data-parallel run re-draws its augmentation every epoch without touching the host.

torch's Philox stream cannot be reproduced element by element by a fused kernel, so the augmentation is STATISTICALLY equal
to the scripts' (noise scale, shift scale, masking rate) and exactly equal in its index logic (row r of the output is source
row r % B; a masked point stays masked)."""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np
import torch
from torch.utils.data import TensorDataset

from . import _ops as P
from .data_util import multimodalDataset

NPZ_KEYS = ("flux", "wavelength", "mask", "phase", "photoflux", "phototime", "photomask", "photowavelength")


def load_photospectra_npz(source, split: Optional[str] = None, device=None):
    """``source``: a path to the ``.npz`` or an already loaded mapping.  ``split``: None (all rows), ``"train"`` / ``"test"``
    (rows ``training_idx`` / ``testing_idx`` of the file, cannon/test_photospectra.py:23-31).
    -> multimodalDataset(photometry TensorDataset(flux, time, band, mask), spectra TensorDataset(flux, wavelength, phase, mask))
    with the scripts' dtypes: float32 values, int64 bands, bool masks where True = unobserved (``npz_mask == 0``)."""
    data: Mapping = np.load(source) if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__") else source
    missing = [k for k in NPZ_KEYS if k not in data]
    if missing:
        raise KeyError(f"photospectra npz misses keys {missing}")
    if split is None:
        rows = slice(None)
    else:
        key = {"train": "training_idx", "test": "testing_idx"}[split]
        rows = np.asarray(data[key])

    def f32(k):
        return torch.tensor(np.asarray(data[k])[rows], dtype=torch.float32, device=device)

    def unobserved(k):
        return torch.tensor(np.asarray(data[k])[rows] == 0, device=device)
    photo = TensorDataset(f32("photoflux"), f32("phototime"),
                          torch.tensor(np.asarray(data["photowavelength"])[rows], dtype=torch.long, device=device), unobserved("photomask"))
    spec = TensorDataset(f32("flux"), f32("wavelength"), f32("phase"), unobserved("mask"))
    return multimodalDataset(photo, spec)


class GpuAugmenter:
    """repeat x `copies`, flux noise, time noise (per point) and / or time shift (per light curve), phase noise, random masking.

    Presets restate the two scripts: ``GpuAugmenter.goldstein()`` (cannon/test_photospectra.py:45-47,75-78: sigma 0.02 on both
    fluxes, one N(0, 0.1) shift per light curve, 5 % extra masking, no repeat) and ``GpuAugmenter.ztf()``
    (cannon/ZTF_photospect.py:46-66: 10 copies, sigma 0.01 on fluxes, 0.001 on times and phases, 10 % masking)."""

    def __init__(self, copies=1, flux_noise=0.0, time_noise=0.0, time_shift=0.0, phase_noise=0.0, mask_p=0.0, seed=0):
        self.copies, self.flux_noise, self.time_noise, self.time_shift = int(copies), float(flux_noise), float(time_noise), float(time_shift)
        self.phase_noise, self.mask_p, self.seed = float(phase_noise), float(mask_p), int(seed)
        self._draw = 0

    @classmethod
    def goldstein(cls, seed=0):
        return cls(copies=1, flux_noise=0.02, time_shift=0.1, mask_p=0.05, seed=seed)

    @classmethod
    def ztf(cls, seed=0):
        return cls(copies=10, flux_noise=0.01, time_noise=0.001, phase_noise=0.001, mask_p=0.1, seed=seed)

    def _seed(self, device):
        self._draw += 1
        return torch.tensor([(self.seed * 0x9E3779B97F4A7C15 + self._draw * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)

    def __call__(self, dataset, device=None):
        """dataset: the multimodalDataset of `load_photospectra_npz` (or a pair of TensorDatasets) -> a new one, augmented, on `device`."""
        photo, spec = dataset.datasets if isinstance(dataset, multimodalDataset) else dataset
        dev = torch.device(device) if device is not None else photo.tensors[0].device
        pf, pt, pb, pm = (t.to(dev).contiguous() for t in photo.tensors)
        sf, sw, sp, sm = (t.to(dev).contiguous() for t in spec.tensors)
        seed = self._seed(dev)
        c = self.copies
        pf2, pm2 = P.augment(pf, pm, c, self.flux_noise, 0.0, self.mask_p, seed, 1)
        pt2, _ = P.augment(pt, None, c, self.time_noise, self.time_shift, 0.0, seed, 2)
        pb2 = pb.repeat((c, 1)) if c > 1 else pb
        sf2, sm2 = P.augment(sf, sm, c, self.flux_noise, 0.0, self.mask_p, seed, 3)
        sw2 = sw.repeat((c, 1)) if c > 1 else sw
        sp2, _ = P.augment(sp.reshape(-1, 1), None, c, self.phase_noise, 0.0, 0.0, seed, 4)
        return multimodalDataset(TensorDataset(pf2, pt2, pb2, pm2), TensorDataset(sf2, sw2, sp2.reshape(-1), sm2))
