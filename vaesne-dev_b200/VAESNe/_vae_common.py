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


def _encode_graphs_enabled(model, x) -> bool:
    import os
    if not x[0].is_cuda:
        return False
    flag = model.__dict__.get("graph_encode")
    if flag is None:
        flag = os.environ.get("VAESNE_CUDA_GRAPH", "0") not in ("", "0")
    return bool(flag) and not torch.cuda.is_current_stream_capturing()


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
            mu, s = self._graphed_posterior_params(x) if _encode_graphs_enabled(self, x) else self._posterior_params(x)
            qz_x = self.qz_x(mu, s)
        return qz_x.mean if mean else qz_x

    # ---- CUDA-graph replay of the encode path (the "encode latents/s" metric) ------------------------------------------
    # The photometry encoder is ~60 kernels of a few microseconds at the batch sizes the regression scripts use (32 .. 512):
    # launch-bound.  Per input signature the third call captures the whole posterior-parameter computation into one graph over
    # static input buffers; later calls copy the batch in and replay.  Opt-in (`vae.graph_encode = True` or VAESNE_CUDA_GRAPH=1);
    # the capture is keyed on the parameters' storage, so an optimiser that re-homes them (FusedAdamW) or a load_state_dict
    # into new storage re-captures instead of replaying stale pointers.
    def _graphed_posterior_params(self, x):
        cache = self.__dict__.setdefault("_encode_graphs", {})
        sig = tuple((tuple(t.shape), t.dtype) for t in x) + (next(self.enc.parameters()).data_ptr(),)
        e = cache.setdefault(sig, {"seen": 0})
        if "graph" not in e:
            if e["seen"] < 2:                       # ordinary calls: they also run every one-time host initialisation
                e["seen"] += 1
                return self._posterior_params(x)
            static = tuple(t.clone() for t in x)
            torch.cuda.synchronize(x[0].device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._posterior_params(static)
            e.update(graph=graph, static=static, out=out)
        else:
            for st, t in zip(e["static"], x):
                st.copy_(t, non_blocking=True)
        e["graph"].replay()
        return e["out"][0].clone(), e["out"][1].clone()

    def reconstruct(self, x, K=1):
        self.eval()
        with torch.no_grad():
            zs, _, _ = self._sample(x, K)
            return self._decode_loc(zs, x)        # px_z.mean == loc for Laplace and Normal
