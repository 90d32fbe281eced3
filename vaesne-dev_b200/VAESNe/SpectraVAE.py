"""Spectra VAE — drop-in for the reference's ``VAESNe/SpectraVAE.py`` (SpectraEnc :11-51,
SpectraDec :53-87, SpectraVAE :90-206)."""
import torch
import torch.distributions as dist
from torch import nn

from ._functions import latent_step
from ._vae_common import FusedVAEMixin, masked_scale_tensor
from .base_vae import VAE
from .util_layers import MLP
from .SpectraLayers import spectraTransformerDecoder, spectraTransformerEncoder


class SpectraEnc(nn.Module):
    def __init__(self, latent_len, latent_dim, model_dim, num_heads, num_layers, ff_dim, dropout=0.1, selfattn=False, concat=True):
        super().__init__()
        self.inference_transformer = spectraTransformerEncoder(
            2 * latent_len, latent_dim, model_dim, num_heads, num_layers, ff_dim, dropout, selfattn, concat)
        self.latent_dim = latent_dim
        self.latent_len = latent_len

    def bottleneck(self, flux, wavelength, phase, mask=None):
        # Positional hand-off exactly as the reference (SpectraVAE.py:40-44): the transformer's signature is
        # (wavelength, flux, ...), so its Linear(1->D) sees the wavelength values and its sinusoid the flux values.
        return self.inference_transformer(flux, wavelength, phase, mask)

    def forward(self, flux, wavelength, phase, mask=None):
        bott = self.bottleneck(flux, wavelength, phase, mask)
        zero = torch.zeros(1, bott.shape[0], self.latent_len, bott.shape[2], device=bott.device)
        _, _, mus, ss = latent_step([bott], [zero], [0], self.latent_len)
        return mus[0], ss[0]


class SpectraDec(nn.Module):
    def __init__(self, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout=0.1, selfattn=False):
        super().__init__()
        self.generativetransformer = spectraTransformerDecoder(latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn)

    def pxz(self, wavelength, phase, z, mask=None):
        return self.generativetransformer(wavelength, phase, z, mask)

    def forward(self, wavelength, phase, z, mask=None):
        x_rec = self.pxz(wavelength, phase, z, mask)
        return x_rec, masked_scale_tensor(mask, 1e10, x_rec)


class SpectraVAE(FusedVAEMixin, VAE):
    _big = 1e10

    def __init__(self, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=0.1,
                 selfattn=False, concat=True, beta=1., prior=dist.Laplace, likelihood=dist.Laplace, posterior=dist.Laplace,
                 **legacy_kwargs):
        # legacy_kwargs swallows `spectra_length=` still passed by cannon/ZTF_photospect.py:89 / ZTF_spectonly.py:57
        VAE.__init__(
            self, prior, likelihood, posterior,
            SpectraEnc(latent_len, latent_dim, model_dim, num_heads, num_layers, ff_dim, dropout, selfattn, concat),
            SpectraDec(latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout),
            params=[latent_len, latent_dim, model_dim, num_heads, num_layers, ff_dim, dropout, selfattn])
        self._pz_params = nn.ParameterList([
            nn.Parameter(torch.zeros(latent_len, latent_dim), requires_grad=False),
            nn.Parameter(torch.ones(latent_len, latent_dim), requires_grad=False)])
        self.llik_scaling = 1. / beta
        self.modelName = 'spectrum'
        self.latent_len = latent_len
        self.latent_dim = latent_dim

    def _bottleneck(self, x):
        flux, wavelength, phase, mask = x
        return self.enc.bottleneck(flux, wavelength, phase, mask)

    def _decode_loc(self, zs, x):
        _, wavelength, phase, mask = x
        R, B = zs.shape[0], zs.shape[1]
        loc = self.dec.generativetransformer.decode_replicated(wavelength, phase, zs.reshape(R * B, zs.shape[-2], zs.shape[-1]), mask, R)
        return loc.view(R, B, wavelength.shape[1])

    def generate(self, N, x):
        self.eval()
        with torch.no_grad():
            zs = self.pz(*self.pz_params).rsample(torch.Size([N, 1]))
            return self._decode_loc(zs, x).unsqueeze(0)


class BrightSpectraVAE(SpectraVAE):
    """Drop-in for ``BrightSpectraVAE`` (reference ``SpectraVAE.py:208-322``): ``loc = dec(z) - mean_L(dec(z)) +
    brightnessfc([z[:, :, 0, :], phase])`` — the brightness head also sees the phase."""

    def __init__(self, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=0.1,
                 selfattn=False, beta=1., prior=dist.Laplace, likelihood=dist.Laplace, posterior=dist.Laplace):
        assert latent_len > 1, "Need at least one token for overall brightness"
        super().__init__(latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn, True, beta,
                         prior, likelihood, posterior)
        self.brightnessfc = MLP(latent_dim + 1, 1, [model_dim])          # phase is added

    def _decode_loc(self, zs, x):
        loc = super()._decode_loc(zs, x)
        phase = x[2]
        feats = torch.cat([zs[:, :, 0, :], phase[None, :, None].expand(zs.shape[0], -1, 1).to(zs.dtype)], dim=-1)
        brightness = self.brightnessfc(feats)                               # [R, B, 1]
        return loc + brightness - loc.mean(dim=2, keepdim=True)
