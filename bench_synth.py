"""Synthetic Goldstein / ZTF-shaped batches for bench.py — shared by the product arm and the reference arm (SURVEY §8d:
photometry flux/time ~ N(0,1), band ~ U{0..nb-1}, 30 % masked with the first point observed; spectra flux ~ N(0,1), wavelength =
linspace(-1.7, 1.7, L), phase ~ N(0,1), 10 % masked plus a padded tail of up to 200 bins on every second row).  Same recipe
as the test-suite's generator; restated here so that the product arm of the benchmark never imports oracle/."""
import torch


def synth_photometry(B, L=60, num_bands=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    flux = torch.randn(B, L, generator=g)
    time = torch.randn(B, L, generator=g)
    band = torch.randint(0, num_bands, (B, L), generator=g)
    mask = torch.rand(B, L, generator=g) < 0.3
    mask[:, 0] = False
    return flux, time, band, mask


def synth_spectra(B, L=982, seed=0):
    g = torch.Generator().manual_seed(seed + 1000)
    flux = torch.randn(B, L, generator=g)
    wavelength = torch.linspace(-1.7, 1.7, L)[None].repeat(B, 1)
    phase = torch.randn(B, generator=g)
    mask = torch.rand(B, L, generator=g) < 0.1
    tail = torch.randint(0, min(201, L), (B,), generator=g)
    rows = torch.arange(0, B, 2)
    cols = torch.arange(L)[None, :]
    pad = cols >= (L - tail[rows])[:, None]
    mask[rows] |= pad & (tail[rows] > 0)[:, None]
    mask[:, 0] = False
    return flux, wavelength, phase, mask


def synth_batch(B, seed, num_bands=2, Lp=60, Ls=982):
    return [synth_photometry(B, Lp, num_bands, seed=seed), synth_spectra(B, Ls, seed=seed)]
