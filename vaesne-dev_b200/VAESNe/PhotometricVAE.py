"""Photometric (light-curve) VAE — drop-in for the reference's ``VAESNe/PhotometricVAE.py``
(PhotometricEnc :10-56, PhotometricDec :58-94, PhotometricVAE :97-222)."""
import torch
import torch.distributions as dist
from torch import nn

from . import _noise
from . import _ops as P
from ._functions import latent_step
from ._vae_common import FusedVAEMixin, masked_scale_tensor
from .base_vae import VAE
from .util_layers import MLP
from .PhotometricLayers import photometricTransformerDecoder, photometricTransformerEncoder


class PhotometricEnc(nn.Module):
    def __init__(self, num_bands, latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers,
                 dropout=0.1, selfattn=False, concat=True):
        super().__init__()
        self.inference_transformer = photometricTransformerEncoder(
            num_bands, 2 * latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn, concat)
        self.latent_dim = latent_dim
        self.latent_len = latent_len

    def forward(self, flux, time, band, mask=None):
        """-> (mu, scale): first latent_len tokens, softplus of the last latent_len tokens."""
        bott = self.inference_transformer(flux, time, band, mask)
        zero = torch.zeros(1, bott.shape[0], self.latent_len, bott.shape[2], device=bott.device)
        _, _, mus, ss = latent_step([bott], [zero], [0], self.latent_len)
        return mus[0], ss[0]


class PhotometricDec(nn.Module):
    def __init__(self, latent_dim, num_bands, model_dim, num_heads, ff_dim, num_layers, dropout=0.1, selfattn=False):
        super().__init__()
        self.generativetransformer = photometricTransformerDecoder(
            latent_dim, num_bands, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn)

    def pxz(self, time, band, z, mask=None):
        return self.generativetransformer(time, band, z, mask)

    def forward(self, time, band, z, mask=None):
        x_rec = self.pxz(time, band, z, mask)
        return x_rec, masked_scale_tensor(mask, 1e8, x_rec)


class PhotometricVAE(FusedVAEMixin, VAE):
    _big = 1e8

    def __init__(self, num_bands=6, latent_len=8, latent_dim=4, model_dim=64, num_heads=4, ff_dim=64, num_layers=4,
                 dropout=0.1, selfattn=False, concat=True, beta=1., prior=dist.Laplace, likelihood=dist.Laplace,
                 posterior=dist.Laplace, **legacy_kwargs):
        # legacy_kwargs swallows `photometric_length=` still passed by cannon/test_photometry.py:58
        VAE.__init__(
            self, prior, likelihood, posterior,
            PhotometricEnc(num_bands, latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn, concat),
            PhotometricDec(latent_dim, num_bands, model_dim, num_heads, ff_dim, num_layers, dropout),
            params=[num_bands, latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn])
        self._pz_params = nn.ParameterList([
            nn.Parameter(torch.zeros(latent_len, latent_dim), requires_grad=False),
            nn.Parameter(torch.ones(latent_len, latent_dim), requires_grad=False)])
        self.llik_scaling = 1. / beta
        self.modelName = 'light_curve'
        self.latent_len = latent_len
        self.latent_dim = latent_dim

    def _bottleneck(self, x):
        flux, time, band, mask = x
        return self.enc.inference_transformer(flux, time, band, mask)

    def _decode_loc(self, zs, x, copies=None):
        """zs [K, B, T, Z] (or any [R, B, T, Z]) -> loc [R, B, L]; row r*B + b pairs z[r, b] with x[b]."""
        _, time, band, mask = x
        R, B = zs.shape[0], zs.shape[1]
        loc = self.dec.generativetransformer.decode_replicated(time, band, zs.reshape(R * B, zs.shape[-2], zs.shape[-1]), mask, R)
        return loc.view(R, B, time.shape[1])

    def generate(self, N, time, band, mask=None):
        # the reference's PhotometricVAE.generate (:211-222) references an undefined K; this is the evident intent
        self.eval()
        with torch.no_grad():
            pz = self.pz(*self.pz_params)
            zs = pz.rsample(torch.Size([N, time.shape[0]]))
            return self._decode_loc(zs, (None, time, band, mask))


class BrightPhotometricVAE(PhotometricVAE):
    """Drop-in for ``BrightPhotometricVAE`` (reference ``PhotometricVAE.py:225-332``): the first latent token carries the
    overall brightness — ``loc = dec(z) - mean_L(dec(z)) + brightnessfc(z[:, :, 0, :])``.  Same encoder / decoder stacks
    and fused objectives as ``PhotometricVAE``; the brightness head is a 2-layer MLP on K*B rows."""

    def __init__(self, num_bands=6, latent_len=8, latent_dim=4, model_dim=64, num_heads=4, ff_dim=64, num_layers=4,
                 dropout=0.1, selfattn=False, beta=1., prior=dist.Laplace, likelihood=dist.Laplace, posterior=dist.Laplace):
        assert latent_len > 1, "first token for overall brightness"
        super().__init__(num_bands, latent_len, latent_dim, model_dim, num_heads, ff_dim, num_layers, dropout, selfattn,
                         True, beta, prior, likelihood, posterior)
        self.brightnessfc = MLP(latent_dim, 1, [model_dim])

    def _decode_loc(self, zs, x, copies=None):
        loc = super()._decode_loc(zs, x)
        brightness = self.brightnessfc(zs[:, :, 0, :])                       # [R, B, 1]
        return loc + brightness - loc.mean(dim=2, keepdim=True)
