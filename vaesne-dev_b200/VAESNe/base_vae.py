"""VAE base class — drop-in for the reference's ``VAESNe/base_vae.py:8-60``: holds the prior /
likelihood / posterior distribution classes, the encoder and decoder modules and ``llik_scaling``."""
import torch
import torch.nn as nn

from .util_layers import get_mean


class VAE(nn.Module):
    def __init__(self, prior_dist, likelihood_dist, post_dist, enc, dec, params):
        super().__init__()
        self.pz = prior_dist
        self.px_z = likelihood_dist
        self.qz_x = post_dist
        self.enc = enc
        self.dec = dec
        self.modelName = None
        self.params = params
        self._pz_params = None       # set by the subclass
        self._qz_x_params = None     # set by forward()
        self.llik_scaling = 1.0

    @property
    def pz_params(self):
        return self._pz_params

    @property
    def qz_x_params(self):
        if self._qz_x_params is None:
            raise NameError("qz_x params not initalised yet!")
        return self._qz_x_params

    @staticmethod
    def getDataLoaders(batch_size, shuffle=True, device="cuda"):
        raise NotImplementedError

    def forward(self, x, K=1):
        self._qz_x_params = self.enc(x)
        qz_x = self.qz_x(*self._qz_x_params)
        zs = qz_x.rsample(torch.Size([K]))
        return qz_x, self.px_z(*self.dec(zs)), zs

    def generate(self, N, K):
        self.eval()
        with torch.no_grad():
            latents = self.pz(*self.pz_params).rsample(torch.Size([N]))
            data = self.px_z(*self.dec(latents)).sample(torch.Size([K]))
        return data.view(-1, *data.size()[3:])

    def reconstruct(self, data):
        self.eval()
        with torch.no_grad():
            latents = self.qz_x(*self.enc(data)).rsample()
            return get_mean(self.px_z(*self.dec(latents)))
