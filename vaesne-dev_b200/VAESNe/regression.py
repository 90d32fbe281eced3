"""Parameter-regression heads — drop-in for the reference's ``VAESNe/regression.py:9-141``: frozen-VAE and
frozen-contrastive encoders followed by an MLP, and the end-to-end encoder + MLP regressors."""
from torch import nn

from .PhotometricLayers import photometricTransformerEncoder
from .SpectraLayers import spectraTransformerEncoder
from .util_layers import MLP


def _freeze(module):
    for p in module.parameters():
        p.requires_grad = False


class VAEregressionHead(nn.Module):
    """vae.encode(x, mean) -> flatten -> MLP; the consumer of the 'encode latents/s' path."""

    def __init__(self, vae, outdim, freeze_vae=True, MLPlatent=[64, 64]):
        super().__init__()
        if freeze_vae:
            _freeze(vae)
        self.vae = vae
        self.outfc = MLP(self.vae.latent_len * self.vae.latent_dim, outdim, MLPlatent)

    def forward(self, x):
        h = self.vae.encode(x, True)
        return self.outfc(h.reshape(h.shape[0], -1))


class contrasphotoregressionHead(nn.Module):
    def __init__(self, contrastnet, outdim, freeze_contrastnet=True, MLPlatent=[64, 64]):
        super().__init__()
        if freeze_contrastnet:
            _freeze(contrastnet)
        self.contrastnet = contrastnet
        self.outfc = MLP(self.contrastnet.latent_len * self.contrastnet.latent_dim, outdim, MLPlatent)

    def forward(self, x):
        h = self.contrastnet.photo_enc(x)
        return self.outfc(h.reshape(h.shape[0], -1))


class contrasspecregressionHead(nn.Module):
    def __init__(self, contrastnet, outdim, freeze_contrastnet=True, MLPlatent=[64, 64]):
        super().__init__()
        if freeze_contrastnet:
            _freeze(contrastnet)
        self.contrastnet = contrastnet
        self.outfc = MLP(self.contrastnet.latent_len * self.contrastnet.latent_dim, outdim, MLPlatent)

    def forward(self, x):
        h = self.contrastnet.spectra_enc(x)
        return self.outfc(h.reshape(h.shape[0], -1))


class photoend2endregression(nn.Module):
    def __init__(self, outdim, num_bands=6, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32,
                 num_layers=4, dropout=0.1, selfattn=False, MLPlatent=[64, 64]):
        super().__init__()
        self.enc = photometricTransformerEncoder(num_bands, latent_len, latent_dim, model_dim, num_heads, ff_dim,
                                                 num_layers, dropout, selfattn)
        self.outfc = MLP(latent_dim * latent_len, outdim, MLPlatent)
        self.latent_dim = latent_dim
        self.latent_len = latent_len

    def forward(self, x):
        flux, time, band, mask = x
        h = self.enc(flux, time, band, mask)
        return self.outfc(h.reshape(h.shape[0], -1))


class specend2endregression(nn.Module):
    def __init__(self, outdim, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, num_layers=4, ff_dim=32,
                 dropout=0.1, selfattn=False, MLPlatent=[64, 64]):
        super().__init__()
        self.enc = spectraTransformerEncoder(latent_len, latent_dim, model_dim, num_heads, num_layers, ff_dim, dropout, selfattn)
        self.outfc = MLP(latent_dim * latent_len, outdim, MLPlatent)
        self.latent_dim = latent_dim
        self.latent_len = latent_len

    def forward(self, x):
        flux, wavelength, phase, mask = x
        # positional hand-off as in the reference (:139)
        h = self.enc(flux, wavelength, phase, mask)
        return self.outfc(h.reshape(h.shape[0], -1))
