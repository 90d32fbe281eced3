"""Spectra transformers — drop-in for the reference's ``VAESNe/SpectraLayers.py`` (decoder :11-63,
encoder :66-138).  Same constructor signatures, attribute names and registration order; forward
passes run as single fused stacks of sm_100a kernels (see ``_stacks.py``)."""
import torch
from torch import nn

from . import _stacks as S
from ._functions import run_stack, _prep, model_dim_of
from .util_layers import (MLP, SinusoidalMLPPositionalEmbedding, SinusoidalPositionalEmbedding,
                          TransformerBlock, aux_tables, singlelayerMLP)


class spectraTransformerDecoder(nn.Module):
    """latent [N, latent_len, bottleneck_dim] + (wavelength, phase) -> flux [N, L]."""

    def __init__(self, bottleneck_dim, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=0.1, selfattn=False):
        super().__init__()
        self.transformerblocks = nn.ModuleList(
            [TransformerBlock(model_dim, num_heads, ff_dim, dropout, selfattn) for _ in range(num_layers)])
        self.wavelength_embd_layer = SinusoidalMLPPositionalEmbedding(model_dim)
        self.phase_embd_layer = SinusoidalMLPPositionalEmbedding(model_dim)
        self.contextfc = MLP(bottleneck_dim, model_dim, [model_dim])
        self.get_flux = singlelayerMLP(model_dim, 1)
        self._model_dim = model_dim
        self._drop_p = float(dropout)

    def decode_replicated(self, wavelength, phase, z, mask, copies):
        """wavelength/mask [B, L], phase [B] un-replicated; z [copies*B, T, Z] with row r = c*B + b
        (SpectraVAE.py:189-194 without materialising the copies)."""
        aux = aux_tables(model_dim_of(self), z.device)
        wavelength, phase, mask, z = _prep(wavelength, torch.float32), _prep(phase, torch.float32), _prep(mask), _prep(z, torch.float32)

        def run(tape, pv, w, p, zz, m):
            return S.spectra_decoder_forward(tape, pv, aux, w, p, zz, m, copies)
        return run_stack(self, run, (wavelength, phase, z, mask))

    def forward(self, wavelength, phase, bottleneck, mask=None):
        return self.decode_replicated(wavelength, phase, bottleneck, mask, 1)


class spectraTransformerEncoder(nn.Module):
    """forward(wavelength, flux, phase, mask): the first argument feeds the parameter-free sinusoid, the
    second the Linear(1 -> model_dim) named ``flux_embd`` (SpectraLayers.py:112-123).  The VAE wrappers
    call it with (flux, wavelength) — see SpectraVAE.SpectraEnc."""

    def __init__(self, bottleneck_length, bottleneck_dim, model_dim, num_heads, num_layers, ff_dim,
                 dropout=0.1, selfattn=False, concat=True):
        super().__init__()
        self.initbottleneck = nn.Parameter(torch.randn(bottleneck_length, model_dim))
        self.flux_embd = nn.Linear(1, model_dim)
        self.transformerblocks = nn.ModuleList(
            [TransformerBlock(model_dim, num_heads, ff_dim, dropout, selfattn) for _ in range(num_layers)])
        self.bottleneckfc = singlelayerMLP(model_dim, bottleneck_dim)
        self.concat = concat
        if concat:
            self.spectrafc = MLP(2 * model_dim, model_dim, [model_dim])
            self.wavelength_embd_layer = SinusoidalPositionalEmbedding(model_dim)
        else:
            self.spectrafc = None
            self.wavelength_embd_layer = SinusoidalMLPPositionalEmbedding(model_dim)
        self.phase_embd_layer = SinusoidalMLPPositionalEmbedding(model_dim)
        self._model_dim = model_dim
        self._drop_p = float(dropout)

    def forward(self, wavelength, flux, phase, mask=None):
        aux = aux_tables(model_dim_of(self), flux.device)
        a1, a2, phase, mask = _prep(wavelength, torch.float32), _prep(flux, torch.float32), _prep(phase, torch.float32), _prep(mask)

        def run(tape, pv, x1, x2, ph, m):
            return S.spectra_encoder_forward(tape, pv, aux, x1, x2, ph, m)
        return run_stack(self, run, (a1, a2, phase, mask))
