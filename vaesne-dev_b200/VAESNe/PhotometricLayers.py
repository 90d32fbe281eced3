"""Photometry (light-curve) transformers — drop-in for the reference's ``VAESNe/PhotometricLayers.py``
(decoder :10-69, encoder :72-143).  Same constructor signatures, attribute names and registration
order; the forward passes run as single fused stacks of sm_100a kernels (see ``_stacks.py``)."""
import torch
from torch import nn

from . import _stacks as S
from ._functions import run_stack, _prep, model_dim_of
from .util_layers import (MLP, SinusoidalMLPPositionalEmbedding, SinusoidalPositionalEmbedding,
                          TransformerBlock, aux_tables, singlelayerMLP)


class photometricTransformerDecoder(nn.Module):
    """latent [N, latent_len, bottleneck_dim] + (time, band) -> flux [N, L]."""

    def __init__(self, bottleneck_dim, num_bands, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                 dropout=0.1, donotmask=False, selfattn=False):
        super().__init__()
        self.transformerblocks = nn.ModuleList(
            [TransformerBlock(model_dim, num_heads, ff_dim, dropout, selfattn) for _ in range(num_layers)])
        self.model_dim = model_dim
        self.sinusoidal_time_embd = SinusoidalMLPPositionalEmbedding(model_dim)
        self.bandembd = nn.Embedding(num_bands, model_dim)
        self.contextfc = MLP(bottleneck_dim, model_dim, [model_dim])
        self.get_photo = singlelayerMLP(model_dim, 1)
        self.donotmask = donotmask
        self._drop_p = float(dropout)

    def decode_replicated(self, time, band, z, mask, copies):
        """time/band/mask are the un-replicated [B, L] inputs; z is [copies*B, T, Z] with row r = c*B + b
        (the K-sample / source replication of PhotometricVAE.py:191-197 without materialising the copies)."""
        if getattr(self, "donotmask", False):
            mask = None
        aux = aux_tables(model_dim_of(self), z.device)
        time, band, mask, z = _prep(time, torch.float32), _prep(band, torch.int64), _prep(mask), _prep(z, torch.float32)

        def run(tape, pv, t, b, zz, m):
            return S.photo_decoder_forward(tape, pv, aux, t, b, zz, m, copies)
        return run_stack(self, run, (time, band, z, mask))

    def forward(self, time, band, bottleneck, mask=None):
        return self.decode_replicated(time, band, bottleneck, mask, 1)


class photometricTransformerEncoder(nn.Module):
    """(flux, time, band, mask) [B, L] -> bottleneck [B, bottleneck_length, bottleneck_dim]."""

    def __init__(self, num_bands, bottleneck_length, bottleneck_dim, model_dim=32, num_heads=4, ff_dim=32,
                 num_layers=4, dropout=0.1, selfattn=False, concat=True):
        super().__init__()
        self.model_dim = model_dim
        self.initbottleneck = nn.Parameter(torch.randn(bottleneck_length, model_dim))
        self.bottleneckfc = singlelayerMLP(model_dim, bottleneck_dim)
        self.transformerblocks = nn.ModuleList(
            [TransformerBlock(model_dim, num_heads, ff_dim, dropout, selfattn) for _ in range(num_layers)])
        self.concat = concat
        self.bandembd = nn.Embedding(num_bands, model_dim)
        self.fluxfc = nn.Linear(1, model_dim)
        if concat:
            self.time_embd = SinusoidalMLPPositionalEmbedding(model_dim)
            self.LCfc = MLP(3 * model_dim, model_dim, [model_dim])
        else:
            self.time_embd = SinusoidalPositionalEmbedding(model_dim)
            self.LCfc = None
        self._drop_p = float(dropout)

    def forward(self, flux, time, band, mask=None):
        aux = aux_tables(model_dim_of(self), flux.device)
        flux, time, band, mask = _prep(flux, torch.float32), _prep(time, torch.float32), _prep(band, torch.int64), _prep(mask)

        def run(tape, pv, f, t, b, m):
            return S.photo_encoder_forward(tape, pv, aux, f, t, b, m)
        return run_stack(self, run, (flux, time, band, mask))
