"""Logic shared by PhotometricVAE and SpectraVAE: posterior heads + sampling through the fused latent
kernel, K-replicated decoding, likelihood scale from the mask."""
import torch

from . import _noise
from . import _ops as P
from ._functions import latent_step, LikSpec


def masked_scale_tensor(mask, big: float, like: torch.Tensor):
    """ones_like(x) + big*mask, evaluated in fp32 (PhotometricVAE.py:91-93 / SpectraVAE.py:84-86)."""
    one = torch.ones((), dtype=torch.float32, device=like.device)
    if mask is None:
        return one.expand(like.shape)
    return torch.where(mask, one + torch.tensor(big, dtype=torch.float32, device=like.device), one)


class FusedVAEMixin:
    """Expects: self.enc.inference_transformer, self.dec.generativetransformer, self.qz_x / px_z / pz,
    self.latent_len, self._big (1e8 | 1e10) and self._bottleneck(x)."""

    def _families(self):
        return _noise.family_of(self.qz_x), _noise.family_of(self.px_z), _noise.family_of(self.pz)

    def _sample(self, x, K):
        bott = self._bottleneck(x)
        fq = P.FAMILY[_noise.family_of(self.qz_x)]
        B, _, Z = bott.shape
        noise = _noise.draw(_noise.family_of(self.qz_x), (K, B, self.latent_len, Z), bott)
        z, _, mus, ss = latent_step([bott], [noise], [fq], self.latent_len)
        return z[0], mus[0], ss[0]

    def lik_spec(self, x) -> LikSpec:
        return LikSpec(x[0], x[3], P.FAMILY[_noise.family_of(self.px_z)], P.masked_scale(self._big), float(self.llik_scaling))

    def forward(self, x, K=1):
        zs, mu, s = self._sample(x, K)
        self._qz_x_params = (mu, s)
        return self.qz_x(mu, s), self.decode(zs, x), zs

    def decode(self, zs, x):
        loc = self._decode_loc(zs, x)
        scale = masked_scale_tensor(x[3], self._big, loc[0])
        return self.px_z(loc, scale.unsqueeze(0).expand(loc.shape))

    def _posterior_params(self, x):
        """(mu, scale) of q(z|x) without drawing a sample: neither the global RNG nor an injected noise tensor is consumed
        (the reference's encode() does not sample, PhotometricVAE.py:179-186)."""
        bott = self._bottleneck(x)
        fq = P.FAMILY[_noise.family_of(self.qz_x)]
        zero = torch.zeros(1, bott.shape[0], self.latent_len, bott.shape[2], device=bott.device)
        _, _, mus, ss = latent_step([bott], [zero], [fq], self.latent_len)
        return mus[0], ss[0]

    def encode(self, x, mean=True):
        self.eval()
        with torch.no_grad():
            mu, s = self._posterior_params(x)
            qz_x = self.qz_x(mu, s)
        return qz_x.mean if mean else qz_x

    def reconstruct(self, x, K=1):
        self.eval()
        with torch.no_grad():
            zs, _, _ = self._sample(x, K)
            return self._decode_loc(zs, x)        # px_z.mean == loc for Laplace and Normal
