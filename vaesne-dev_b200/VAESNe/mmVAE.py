"""Mixture-of-experts multimodal VAE — drop-in for ``photospecMMVAE`` of the reference's
``VAESNe/mmVAE.py:71-126`` (MoE-MMVAE of Shi et al. 2019).

The M x M cross-modal decode matrix is produced with ONE pass per decoder: the K samples of every
source modality are stacked (row = (source*K + k)*B + b) and each decoder runs once on M*K*B rows,
instead of M separate passes of K*B rows."""
import torch
import torch.distributions as dist
import torch.nn as nn

from . import _noise
from . import _ops as P
from ._functions import latent_step


class MMVAE(nn.Module):
    """Generic base of the reference (mmVAE.py:17-67): unused by every script and inconsistent with the decoder signatures
    (its off-diagonal `vae.dec(zs)` call, :47); outside the accelerated path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("VAESNe-B200: the generic MMVAE base (reference mmVAE.py:17-67) is not provided; use photospecMMVAE")


class photospecMMVAE(nn.Module):
    def __init__(self, vaes, prior_dist=dist.Laplace, beta=1., length_ratio=982 / 60):
        super().__init__()
        self.pz = prior_dist
        self.vaes = nn.ModuleList(vaes)
        self.modelName = "photospectra"
        self._pz_params = nn.ParameterList([
            nn.Parameter(torch.zeros(vaes[0].latent_len, vaes[0].latent_dim), requires_grad=False),
            nn.Parameter(torch.ones(vaes[0].latent_len, vaes[0].latent_dim), requires_grad=False)])
        self.vaes[0].llik_scaling = 1. / beta
        self.vaes[1].llik_scaling = 1. / beta
        self.vaes[0].llik_scaling *= length_ratio

    @property
    def pz_params(self):
        return self._pz_params

    # ---- fused building blocks (also used by losses.m_iwae) ------------------------------------
    def _encode_sample(self, x, K, want_lat):
        botts = [vae._bottleneck(x[m]) for m, vae in enumerate(self.vaes)]
        fams = [P.FAMILY[_noise.family_of(vae.qz_x)] for vae in self.vaes]
        T = self.vaes[0].latent_len
        noises = [_noise.draw(_noise.family_of(vae.qz_x), (K, b.shape[0], T, b.shape[2]), b) for vae, b in zip(self.vaes, botts)]
        fp = P.FAMILY[_noise.family_of(self.pz)]
        z, lat, mus, ss = latent_step(botts, noises, fams, T, fp, self._pz_params[0], self._pz_params[1], want_lat)
        return z, lat, mus, ss

    def _decode_all(self, z, x):
        """z [M, K, B, T, Z] -> list over d of loc_d [M, K, B, L_d]."""
        M, K, B = z.shape[0], z.shape[1], z.shape[2]
        zz = z.reshape(M * K, B, z.shape[3], z.shape[4])
        return [vae._decode_loc(zz, x[d]).view(M, K, B, -1) for d, vae in enumerate(self.vaes)]

    def forward(self, x, K=1):
        from ._vae_common import masked_scale_tensor
        z, _, mus, ss = self._encode_sample(x, K, want_lat=False)
        locs = self._decode_all(z, x)
        M = len(self.vaes)
        qz_xs = [vae.qz_x(mus[m], ss[m]) for m, vae in enumerate(self.vaes)]
        for m, vae in enumerate(self.vaes):
            vae._qz_x_params = (mus[m], ss[m])
        px_zs = [[None] * M for _ in range(M)]
        for d, vae in enumerate(self.vaes):
            scale = masked_scale_tensor(x[d][3], vae._big, locs[d][0, 0]).unsqueeze(0)
            for e in range(M):
                px_zs[e][d] = vae.px_z(locs[d][e], scale.expand(locs[d][e].shape))
        return qz_xs, px_zs, [z[m] for m in range(M)]

    def generate(self, N, x):
        self.eval()
        with torch.no_grad():
            pz = self.pz(*self.pz_params)
            latents = pz.rsample(torch.Size([N, x[0][0].shape[0]]))
            return [vae._decode_loc(latents, x[d]) for d, vae in enumerate(self.vaes)]

    def reconstruct(self, data, K=1):
        self.eval()
        with torch.no_grad():
            z, _, _, _ = self._encode_sample(data, K, want_lat=False)
            locs = self._decode_all(z, data)
            M = len(self.vaes)
            return [[locs[d][e] for d in range(M)] for e in range(M)]
