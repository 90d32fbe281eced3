"""Contrastive photometry/spectra pre-training net — drop-in for ``ContraPhotSpec`` of the reference's
``VAESNe/contrastiveNets.py:20-101`` (both encoders with bottleneck_length = latent_len, projection heads)."""
import torch
import torch.nn as nn

from .PhotometricLayers import photometricTransformerEncoder
from .SpectraLayers import spectraTransformerEncoder
from .util_layers import singlelayerMLP


class ContraPhotSpec(nn.Module):
    def __init__(self, latent_len, latent_dim, proj_dim,
                 num_bands, photo_model_dim, photo_num_heads, photo_ff_dim, photo_num_layers, photo_dropout,
                 spec_model_dim, spec_num_heads, spec_num_layers, spec_ff_dim, spec_dropout, selfattn):
        super().__init__()
        self.photometry_encoder = photometricTransformerEncoder(
            num_bands, latent_len, latent_dim, photo_model_dim, photo_num_heads, photo_ff_dim, photo_num_layers,
            photo_dropout, selfattn)
        self.photo_proj = singlelayerMLP(latent_len * latent_dim, proj_dim)
        self.spectra_encoder = spectraTransformerEncoder(
            latent_len, latent_dim, spec_model_dim, spec_num_heads, spec_num_layers, spec_ff_dim, spec_dropout, selfattn)
        self.spectra_proj = singlelayerMLP(latent_len * latent_dim, proj_dim)
        self.latent_dim = latent_dim
        self.latent_len = latent_len
        self.proj_dim = proj_dim

    def forward(self, x):
        photo_flux, time, band, photo_mask = x[0]
        spec_flux, wavelength, phase, spec_mask = x[1]
        z1 = self.photometry_encoder(photo_flux, time, band, photo_mask)
        # positional hand-off as in the reference (:84): (flux, wavelength) into forward(wavelength, flux)
        z2 = self.spectra_encoder(spec_flux, wavelength, phase, spec_mask)
        return self.photo_proj(z1.reshape(z1.shape[0], -1)), self.spectra_proj(z2.reshape(z2.shape[0], -1))

    def photo_enc(self, x):
        photo_flux, time, band, photo_mask = x
        self.eval()
        with torch.no_grad():
            return self.photometry_encoder(photo_flux, time, band, photo_mask)

    def spectra_enc(self, x):
        self.eval()
        spec_flux, wavelength, phase, spec_mask = x
        with torch.no_grad():
            return self.spectra_encoder(spec_flux, wavelength, phase, spec_mask)
