"""Building blocks of the VAESNe transformers — B200-native drop-in for the reference's
``VAESNe/util_layers.py`` (hot-path symbols only: :9-34 MLPs, :113-149 sinusoidal embeddings,
:257-309 TransformerBlock, :313-336 get_mean / log_mean_exp / kl_divergence).

The modules here are parameter containers with the reference's attribute names, registration order
and initialisers (so ``state_dict`` files and seeded initialisation are interchangeable); their
arithmetic is executed by the sm_100a kernels behind ``include/vaesne_b200.h``.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _ops as P
from . import _stacks as S
from ._functions import run_stack, _prep

_MODEL_DIM = 32


def _freq_table(dim: int, step: int) -> torch.Tensor:
    # same fp32 CPU evaluation as the reference constructors (util_layers.py:122,138)
    return torch.exp(torch.arange(0, dim, step).float() * (-torch.log(torch.tensor(10000.0)) / dim))


class _Aux:
    """Per-device copies of the two sinusoid frequency tables."""

    def __init__(self, dim: int):
        self.dim = dim
        self._cpu = {"div_full": _freq_table(dim, 1), "div_half": _freq_table(dim, 2)}
        self._dev = {}

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {k: v.to(device) for k, v in self._cpu.items()}
        return self._dev[key]


_AUX = {}


def aux_tables(dim: int, device):
    if dim not in _AUX:
        _AUX[dim] = _Aux(dim)
    return _AUX[dim].on(device)


def _flat2d(x: torch.Tensor):
    x = _prep(x, torch.float32)
    return x.reshape(-1, x.shape[-1]), x.shape[:-1]


# ---------------------------------------------------------------------------------------------
class singlelayerMLP(nn.Module):
    """Linear(in,in) -> ReLU -> Linear(in,out)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, in_dim)
        self.fc2 = nn.Linear(in_dim, out_dim)

    def forward(self, x):
        x2, lead = _flat2d(x)

        def run(tape, pv, xin):
            h = S.t_lin(tape, pv, xin, "fc1.weight", "fc1.bias", act=P.ACT_RELU)
            return S.t_lin(tape, pv, h, "fc2.weight", "fc2.bias")
        return run_stack(self, run, (x2,)).view(*lead, -1)


class MLP(nn.Module):
    """Linear/ReLU pairs over ``hidden_dim`` followed by a final Linear (attribute ``mlp``)."""

    def __init__(self, in_dim, out_dim, hidden_dim=[64, 64]):
        super().__init__()
        dims = [in_dim] + list(hidden_dim)
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.ReLU()]
        layers.append(nn.Linear(dims[-1], out_dim))
        self.mlp = nn.Sequential(*layers)

    def forward(self, x):
        x2, lead = _flat2d(x)
        idx = [i for i, m in enumerate(self.mlp) if isinstance(m, nn.Linear)]

        def run(tape, pv, xin):
            h = xin
            for i in idx[:-1]:
                h = S.t_lin(tape, pv, h, f"mlp.{i}.weight", f"mlp.{i}.bias", act=P.ACT_RELU)
            return S.t_lin(tape, pv, h, f"mlp.{idx[-1]}.weight", f"mlp.{idx[-1]}.bias")
        return run_stack(self, run, (x2,)).view(*lead, -1)


class SinusoidalPositionalEmbedding(nn.Module):
    """[sin(x*w_j), cos(x*w_j)] with dim/2 frequencies; parameter free."""

    def __init__(self, dim=64):
        super().__init__()
        self.dim = dim
        self.div_term = _freq_table(dim, 2)

    def forward(self, x):
        x = _prep(x, torch.float32)
        div = self.div_term.to(x.device)
        out = torch.empty(x.numel(), self.dim, device=x.device, dtype=torch.float32)
        P.sincos_feat(x.reshape(-1), div, out)
        return out.view(*x.shape, self.dim)


class SinusoidalMLPPositionalEmbedding(nn.Module):
    """dim frequencies -> 2*dim features -> fc1 -> ReLU -> fc2."""

    def __init__(self, dim=64):
        super().__init__()
        self.dim = dim
        self.div_term = _freq_table(dim, 1)
        self.fc1 = nn.Linear(2 * dim, dim)
        self.fc2 = nn.Linear(dim, dim)

    def forward(self, x):
        x = _prep(x, torch.float32)
        div = self.div_term.to(x.device)

        def run(tape, pv, xin):
            return S.sin_mlp(tape, pv, "", xin.reshape(-1), div)
        out = run_stack(self, run, (x,))
        return out.view(*x.shape, self.dim)


class TransformerBlock(nn.Module):
    """Post-LN block: self-attention, optional context self-attention, cross-attention, GELU FFN."""

    def __init__(self, embed_dim, num_heads, ff_dim, dropout=0.1, context_self_attn=False):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
        self.cross_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
        if context_self_attn:
            self.context_self_attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
            self.layernorm_context = nn.LayerNorm(embed_dim)
        else:
            self.context_self_attn = None
        self.ffn = nn.Sequential(nn.Linear(embed_dim, ff_dim), nn.GELU(), nn.Linear(ff_dim, embed_dim))
        self.layernorm1 = nn.LayerNorm(embed_dim)
        self.layernorm2 = nn.LayerNorm(embed_dim)
        self.layernorm3 = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(dropout)
        self._drop_p = float(dropout)
        if embed_dim != _MODEL_DIM or num_heads != 4:
            self._unsupported = f"embed_dim={embed_dim}, num_heads={num_heads}"
        else:
            self._unsupported = None

    def forward(self, x, context=None, mask=None, context_mask=None):
        if self.__dict__.get("_unsupported"):
            raise NotImplementedError(f"VAESNe-B200 kernels support embed_dim=32 / 4 heads only ({self._unsupported})")
        if context is None:
            raise NotImplementedError("TransformerBlock without a context is not used by the VAESNe hot path")
        x = _prep(x, torch.float32); context = _prep(context, torch.float32)
        Nb, Lq, Lc = x.shape[0], x.shape[1], context.shape[1]
        m, cm = _prep(mask), _prep(context_mask)

        def run(tape, pv, xin, cin):
            return S.block_forward(tape, pv, "", xin.view(Nb * Lq, 32), cin.view(Nb * Lc, 32), m, cm, Nb, Lq, Lc)
        return run_stack(self, run, (x, context)).view(Nb, Lq, 32)


# ---------------------------------------------------------------------------------------------
def get_mean(d, K=100):
    """Mean of a distribution, estimated from K samples when no closed form exists."""
    try:
        return d.mean
    except NotImplementedError:
        return d.rsample(torch.Size([K])).mean(0)


def log_mean_exp(value, dim=0, keepdim=False):
    return torch.logsumexp(value, dim, keepdim=keepdim) - math.log(value.size(dim))


def kl_divergence(d1, d2, K=100):
    """Closed-form KL when registered, otherwise a K-sample Monte-Carlo estimate."""
    if (type(d1), type(d2)) in torch.distributions.kl._KL_REGISTRY:
        return torch.distributions.kl_divergence(d1, d2)
    z = d1.rsample(torch.Size([K]))
    return (d1.log_prob(z) - d2.log_prob(z)).mean(0)
