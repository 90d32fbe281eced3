"""Kernel-level orchestration of the encoder / decoder stacks and the latent / objective steps.

Each stack (one photometry or spectra encoder, one decoder) is ONE ``torch.autograd.Function``:
its forward enqueues the C-ABI kernels in order and records, on a small tape, the matching backward
launches; its backward replays the tape in reverse, accumulating parameter gradients straight into
one flat fp32 buffer per stack (the data-parallel all-reduce bucket and the fused-AdamW operand).
No torch arithmetic op runs on the path — torch only provides memory and the autograd hand-off.

What is restated (paths under /root/reference/package/VAESNe):
  TransformerBlock.forward                       util_layers.py:285-309
  photometricTransformerEncoder / Decoder        PhotometricLayers.py:117-143 / :49-69
  spectraTransformerEncoder / Decoder            SpectraLayers.py:112-138 / :46-63
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence

import os

import torch

from . import _ops as P

_call_counter = itertools.count(1)


# ------------------------------------------------------------------------------------------------
# tape machinery
# ------------------------------------------------------------------------------------------------
class _Bwd:
    """State of one backward pass: activation grads (by tensor identity) and parameter grads."""

    def __init__(self, pg: Dict[str, Optional[torch.Tensor]]):
        self.pg = pg
        self.g: Dict[tuple, torch.Tensor] = {}

    @staticmethod
    def key(t: torch.Tensor):
        # views of one buffer that start at the same element and cover the same number of elements
        # denote the same activation (x and x.view(...)); column slices differ in numel.
        return (t.data_ptr(), t.numel())

    @staticmethod
    def _as(g: torch.Tensor, t: torch.Tensor):
        return g if g.shape == t.shape else g.view(t.shape)

    def slot(self, t: torch.Tensor):
        """(grad buffer of activation t, already_holds_a_contribution)."""
        k = self.key(t)
        g = self.g.get(k)
        if g is None:
            g = torch.empty(t.shape, device=t.device, dtype=torch.float32)
            self.g[k] = g
            return g, False
        return self._as(g, t), True

    def seed(self, t: torch.Tensor, grad: torch.Tensor):
        self.g[self.key(t)] = grad

    def seed_key(self, key: tuple, grad: torch.Tensor):
        assert key not in self.g and key[1] > 0, "ambiguous activation identity (zero-sized or aliased stack output)"
        self.g[key] = grad

    def peek(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        g = self.g.get(self.key(t))
        return None if g is None else self._as(g, t)

    def take(self, t: torch.Tensor) -> Optional[torch.Tensor]:
        g = self.g.pop(self.key(t), None)
        return None if g is None else self._as(g, t)


class Tape:
    def __init__(self, enabled: bool, drop_p: float, device):
        self.enabled = enabled
        self.ops: List = []
        self.drop_p = float(drop_p)
        self.seed = P.next_seed(device) if self.drop_p > 0 else None
        self.base = (next(_call_counter) * 4096) & 0xFFFFFFFF
        self.nsites = 0

    def drop(self) -> P.Drop:
        if self.seed is None:
            return P.NO_DROP
        self.nsites += 1
        return P.Drop(self.drop_p, self.seed, self.base + self.nsites)

    def push(self, fn):
        if self.enabled:
            self.ops.append(fn)

    def backward(self, bw: _Bwd):
        for fn in reversed(self.ops):
            fn(bw)
        self.ops = []


class PView:
    """Parameters of one stack, addressed by their local names (those of the reference's modules)."""

    def __init__(self, names: Sequence[str], tensors: Sequence[torch.Tensor]):
        self.p = dict(zip(names, tensors))

    def __getitem__(self, k):
        return self.p[k]

    def has(self, k):
        return k in self.p


def _pgv(bw: _Bwd, name: str, rows: Optional[slice] = None):
    g = bw.pg.get(name)
    if g is None:
        return None
    return g[rows] if rows is not None else g


# ------------------------------------------------------------------------------------------------
# taped primitives
# ------------------------------------------------------------------------------------------------
def t_lin(tape: Tape, pv: PView, X, wname: str, bname: str, *, rows: Optional[slice] = None, Xadd=None, act=P.ACT_NONE,
          R=None, ln: Optional[str] = None, need_dX=True, Y=None, dropout=False):
    """Y = act((X [+Xadd]) W^T + b)  or  LayerNorm_ln(R + dropout(X W^T + b)).  `rows` selects a row block of a
    packed projection (q / kv parts of in_proj_weight)."""
    W = pv[wname][rows] if rows is not None else pv[wname]
    b = pv[bname][rows] if rows is not None else pv[bname]
    T = X.shape[0]
    H = torch.empty(T, W.shape[0], device=X.device, dtype=torch.float32) if (act == P.ACT_GELU and tape.enabled) else None
    S = torch.empty(T, 32, device=X.device, dtype=torch.float32) if (R is not None and tape.enabled) else None
    drop = tape.drop() if (dropout and R is not None) else P.NO_DROP
    gamma = pv[ln + ".weight"] if ln else None
    beta = pv[ln + ".bias"] if ln else None
    Y = P.lin_fwd(X, W, b, Xadd=Xadd, act=act, H=H, R=R, gamma=gamma, beta=beta, S=S, drop=drop, Y=Y)

    def bwd(bw: _Bwd):
        dY = bw.take(Y)
        if dY is None:
            return
        dW, db = _pgv(bw, wname, rows), _pgv(bw, bname, rows)
        dX = dXacc = None
        if need_dX:
            dX, dXacc = bw.slot(X)
        kw = {}
        if R is not None:
            dR, dRacc = bw.slot(R)
            kw = dict(S=S, gamma=gamma, dgamma=_pgv(bw, ln + ".weight"), dbeta=_pgv(bw, ln + ".bias"), dR=dR, dR_acc=dRacc)
        A = H if act == P.ACT_GELU else (Y if act == P.ACT_RELU else None)
        P.lin_bwd(dY, X if dW is not None else None, W, Xadd=Xadd if dW is not None else None, act=act, A=A, dW=dW, db=db,
                  dX=dX, dX_acc=bool(dXacc), drop=drop, **kw)
        if Xadd is not None and need_dX:
            assert not dXacc, "Xadd input must receive the first gradient contribution"
            if bw.peek(Xadd) is None:
                # d(X + Xadd) is the gradient of both addends: hand the SAME buffer to Xadd instead of copying it (1 GB at the
                # spectra head).  Safe on the tape: Xadd's producer takes (reads) it in the very next backward ops, X keeps
                # accumulating into it only later (X is the stack input, consumed by the first block / the expansion).
                bw.seed(Xadd, dX.view(Xadd.shape) if dX.shape != Xadd.shape else dX)
            else:
                dA, acc = bw.slot(Xadd)
                P.copy3d(dX, 0, 0, dA, 0, 0, 1, 1, dX.numel(), accumulate=acc)
    tape.push(bwd)
    return Y


def t_attn(tape: Tape, qbuf, q, kvbuf, k, v, mask, dropout=True):
    """q/k/v are [N,L,32] views into the packed projection buffers qbuf / kvbuf (which may be the same)."""
    drop = tape.drop() if dropout else P.NO_DROP
    O, LSE = P.attn_fwd(q, k, v, mask, drop)

    def bwd(bw: _Bwd):
        dO = bw.take(O)
        if dO is None:
            return
        if qbuf is kvbuf:
            g, acc = bw.slot(qbuf)
            assert not acc
            C = q.shape[-1]
            P.attn_bwd(q, k, v, mask, O, LSE, dO, g[..., 0:C], g[..., C:2 * C], g[..., 2 * C:3 * C], drop)
        else:
            gq, acc1 = bw.slot(qbuf)
            gkv, acc2 = bw.slot(kvbuf)
            assert not acc1 and not acc2
            C = q.shape[-1]
            P.attn_bwd(q, k, v, mask, O, LSE, dO, gq, gkv[..., 0:C], gkv[..., C:2 * C], drop)
    tape.push(bwd)
    return O


def t_expand(tape: Tape, src, copies: int, pgname: Optional[str] = None):
    """dst[c*Bs + b] = src[b].  With `pgname`, src is a parameter and its gradient goes to the flat buffer."""
    dst = P.expand_rows(src, copies)

    def bwd(bw: _Bwd):
        d = bw.take(dst)
        if d is None:
            return
        if pgname is not None:
            g = bw.pg.get(pgname)
            if g is not None:
                P.expand_rows_bwd(d, src.shape[0], copies, g.view(src.shape), accumulate=True)
            return
        g, acc = bw.slot(src)
        P.expand_rows_bwd(d, src.shape[0], copies, g, accumulate=acc)
    tape.push(bwd)
    return dst


def t_concat_tokens(tape: Tape, a, b):
    """cat([a [G,Ra,32], b [G,Rb,32]], dim=1)."""
    G, Ra, Rb = a.shape[0], a.shape[1], b.shape[1]
    out = torch.empty(G, Ra + Rb, 32, device=a.device, dtype=torch.float32)
    P.copy3d(a, Ra * 32, 32, out, (Ra + Rb) * 32, 32, G, Ra, 32)
    P.copy3d(b, Rb * 32, 32, out, (Ra + Rb) * 32, 32, G, Rb, 32, dst_off=Ra * 32)

    def bwd(bw: _Bwd):
        d = bw.take(out)
        if d is None:
            return
        ga, acca = bw.slot(a)
        P.copy3d(d, (Ra + Rb) * 32, 32, ga, Ra * 32, 32, G, Ra, 32, accumulate=acca)
        gb, accb = bw.slot(b)
        P.copy3d(d, (Ra + Rb) * 32, 32, gb, Rb * 32, 32, G, Rb, 32, accumulate=accb, src_off=Ra * 32)
    tape.push(bwd)
    return out


def t_gather(tape: Tape, pv: PView, idx, tname: str, out, accumulate: bool):
    """out (+)= table[idx]; `out` is a row-strided [T,32] view."""
    P.gather_rows(idx, pv[tname], out, accumulate)

    def bwd(bw: _Bwd, _out=out):
        dt = _pgv(bw, tname)
        if dt is None:
            return
        d = bw.peek(_out)             # the gradient stays in place for the co-producers of `out`
        if d is not None:
            P.scatter_rows(idx, d, dt)
    tape.push(bwd)


def _j(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


def sin_mlp(tape: Tape, pv: PView, prefix: str, x_flat, div, Y=None):
    """SinusoidalMLPPositionalEmbedding (util_layers.py:142-149): sincos(x*div) -> fc1 -> ReLU -> fc2."""
    T = x_flat.numel()
    feats = torch.empty(T, 2 * div.numel(), device=x_flat.device, dtype=torch.float32)
    P.sincos_feat(x_flat, div, feats)
    h = t_lin(tape, pv, feats, _j(prefix, "fc1.weight"), _j(prefix, "fc1.bias"), act=P.ACT_RELU, need_dX=False)
    return t_lin(tape, pv, h, _j(prefix, "fc2.weight"), _j(prefix, "fc2.bias"), Y=Y)


def mlp2(tape: Tape, pv: PView, prefix: str, X, need_dX=True):
    """util_layers.MLP with one hidden layer: mlp.0 -> ReLU -> mlp.2."""
    h = t_lin(tape, pv, X, _j(prefix, "mlp.0.weight"), _j(prefix, "mlp.0.bias"), act=P.ACT_RELU, need_dX=need_dX)
    return t_lin(tape, pv, h, _j(prefix, "mlp.2.weight"), _j(prefix, "mlp.2.bias"))


def head(tape: Tape, pv: PView, prefix: str, X, Xadd):
    """singlelayerMLP(x0 + h) (util_layers.py:9-18)."""
    h = t_lin(tape, pv, X, _j(prefix, "fc1.weight"), _j(prefix, "fc1.bias"), Xadd=Xadd, act=P.ACT_RELU)
    return t_lin(tape, pv, h, _j(prefix, "fc2.weight"), _j(prefix, "fc2.bias"))


# ------------------------------------------------------------------------------------------------
# TransformerBlock (util_layers.py:285-309)
# ------------------------------------------------------------------------------------------------
_NO_SHARED = os.environ.get("VAESNE_NO_SHARED_LAYER0", "0") not in ("", "0")      # A/B switch for the measurement in bench.py
_INPROJ_ONLY = False      # tests: take the "shared projection, per-replica attention" form of a decoder's first block at dropout 0 too


def block_forward(tape: Tape, pv: PView, pre: str, x, ctx, mask, ctx_mask, Nb: int, Lq: int, Lc: int, shared=None):
    """x [Nb*Lq, 32], ctx [Nb*Lc, 32] (2-D, contiguous) -> [Nb*Lq, 32].

    `shared = (xs, copies)`: x is `copies` replicas of xs [Nb/copies * Lq, 32] (a decoder's first block: its queries are
    the position embeddings of the data, identical for all K samples and both sources).  Without dropout the first
    sub-layer LN1(x + SelfMHA(x)) is then identical across the replicas as well, so it runs once per object and is
    replicated afterwards — for the 982-token spectra decoder that removes one of its four 982 x 982 attentions from
    (copies-1)/copies of the rows (reconstruct / generate at K=100; training only when dropout is 0, because the
    reference draws an independent dropout mask per replica)."""
    sa, ca = _j(pre, "self_attn"), _j(pre, "cross_attn")
    pre = pre + "." if pre else ""
    if shared is not None and tape.drop_p == 0.0 and shared[1] > 1 and not _NO_SHARED and not _INPROJ_ONLY:
        xs, copies = shared
        Ns = Nb // copies
        qkv = t_lin(tape, pv, xs, sa + ".in_proj_weight", sa + ".in_proj_bias")
        q3 = qkv.view(Ns, Lq, 96)
        a = t_attn(tape, q3, q3[..., 0:32], q3, q3[..., 32:64], q3[..., 64:96], mask)
        x1s = t_lin(tape, pv, a.view(Ns * Lq, 32), sa + ".out_proj.weight", sa + ".out_proj.bias", R=xs, ln=pre + "layernorm1", dropout=True)
        x1 = t_expand(tape, x1s.view(Ns, Lq * 32), copies).view(Nb * Lq, 32)
    else:
        if shared is not None and shared[1] > 1 and not _NO_SHARED:
            # a decoder's first block under dropout: the replicas draw independent masks, but the block INPUT is still `copies`
            # replicas of xs, so q|k|v are projected once per object and replicated (same values bit for bit) — and in the
            # backward the replicas' d(q|k|v) are summed first, so the projection's weight / input gradients run on 1/copies
            # of the tokens (the packed 32 -> 96 backward is the most expensive linear call of the step)
            xs, copies = shared
            qkv_s = t_lin(tape, pv, xs, sa + ".in_proj_weight", sa + ".in_proj_bias")
            qkv = t_expand(tape, qkv_s.view(Nb // copies, Lq * 96), copies).view(Nb * Lq, 96)
        else:
            qkv = t_lin(tape, pv, x, sa + ".in_proj_weight", sa + ".in_proj_bias")
        q3 = qkv.view(Nb, Lq, 96)
        a = t_attn(tape, q3, q3[..., 0:32], q3, q3[..., 32:64], q3[..., 64:96], mask)
        x1 = t_lin(tape, pv, a.view(Nb * Lq, 32), sa + ".out_proj.weight", sa + ".out_proj.bias", R=x, ln=pre + "layernorm1", dropout=True)

    c = ctx
    if pv.has(pre + "context_self_attn.in_proj_weight"):        # :296-299, the update stays local to this block
        cs = pre + "context_self_attn"
        cqkv = t_lin(tape, pv, ctx, cs + ".in_proj_weight", cs + ".in_proj_bias")
        c3 = cqkv.view(Nb, Lc, 96)
        cattn = t_attn(tape, c3, c3[..., 0:32], c3, c3[..., 32:64], c3[..., 64:96], ctx_mask)
        c = t_lin(tape, pv, cattn.view(Nb * Lc, 32), cs + ".out_proj.weight", cs + ".out_proj.bias", R=ctx,
                  ln=pre + "layernorm_context", dropout=True)

    qc = t_lin(tape, pv, x1, ca + ".in_proj_weight", ca + ".in_proj_bias", rows=slice(0, 32))
    kvc = t_lin(tape, pv, c, ca + ".in_proj_weight", ca + ".in_proj_bias", rows=slice(32, 96))
    qc3, kv3 = qc.view(Nb, Lq, 32), kvc.view(Nb, Lc, 64)
    ac = t_attn(tape, qc3, qc3, kv3, kv3[..., 0:32], kv3[..., 32:64], ctx_mask)
    x2 = t_lin(tape, pv, ac.view(Nb * Lq, 32), ca + ".out_proj.weight", ca + ".out_proj.bias", R=x1, ln=pre + "layernorm2", dropout=True)

    g = t_lin(tape, pv, x2, pre + "ffn.0.weight", pre + "ffn.0.bias", act=P.ACT_GELU)
    return t_lin(tape, pv, g, pre + "ffn.2.weight", pre + "ffn.2.bias", R=x2, ln=pre + "layernorm3", dropout=True)


def _num_blocks(pv: PView) -> int:
    i = 0
    while pv.has(f"transformerblocks.{i}.self_attn.in_proj_weight"):
        i += 1
    return i


# ------------------------------------------------------------------------------------------------
# stacks
# ------------------------------------------------------------------------------------------------
def _check_geometry(pv: PView, key: str):
    w = pv[key]
    if w.shape[-1] != 32:
        raise NotImplementedError(
            f"VAESNe-B200 kernels support model_dim=32, num_heads=4 (the geometry of every reference script); got model_dim={w.shape[-1]}")


def photo_encoder_forward(tape, pv: PView, aux, flux, time, band, mask):
    """PhotometricLayers.py:117-143: concat=True -> LCfc(cat[fluxfc, SinMLP(time), bandembd]); concat=False -> their sum with
    the parameter-free sinusoid (:132-135)."""
    _check_geometry(pv, "initbottleneck")
    B, L = flux.shape
    T = B * L
    if not pv.has("LCfc.mlp.0.weight"):
        ctx = t_lin(tape, pv, flux.reshape(T, 1), "fluxfc.weight", "fluxfc.bias", need_dX=False)
        pos = torch.empty(T, 32, device=flux.device, dtype=torch.float32)
        P.sincos_feat(time.reshape(T), aux["div_half"], pos)
        ctx.add_(pos)                                   # no parameters behind it: d(ctx) passes through unchanged
        t_gather(tape, pv, band.reshape(T), "bandembd.weight", ctx, accumulate=True)
        return _encoder_tail(tape, pv, ctx, B, L, mask)
    feats = torch.empty(T, 96, device=flux.device, dtype=torch.float32)
    t_lin(tape, pv, flux.reshape(T, 1), "fluxfc.weight", "fluxfc.bias", need_dX=False, Y=feats[:, 0:32])
    sin_mlp(tape, pv, "time_embd", time.reshape(T), aux["div_full"], Y=feats[:, 32:64])
    t_gather(tape, pv, band.reshape(T), "bandembd.weight", feats[:, 64:96], accumulate=False)
    # the three feature producers share one gradient buffer: its column ranges become their output grads
    _alias_feature_grads(tape, feats, [(0, 32), (32, 64), (64, 96)])
    ctx = mlp2(tape, pv, "LCfc", feats, need_dX=True)
    return _encoder_tail(tape, pv, ctx, B, L, mask)


def spectra_encoder_forward(tape, pv: PView, aux, arg1, arg2, phase, mask):
    """SpectraLayers.py:112-138: forward(wavelength=arg1, flux=arg2, phase, mask).  concat=True -> spectrafc(cat[flux_embd(arg2),
    Sin(arg1)]); concat=False -> flux_embd(arg2) + SinMLP(arg1) (:124-126)."""
    _check_geometry(pv, "initbottleneck")
    B, L = arg1.shape
    T = B * L
    if not pv.has("spectrafc.mlp.0.weight"):
        emb = sin_mlp(tape, pv, "wavelength_embd_layer", arg1.reshape(T), aux["div_full"])
        lin = t_lin(tape, pv, arg2.reshape(T, 1), "flux_embd.weight", "flux_embd.bias", need_dX=False)
        emb.add_(lin)

        def share_grad(bw: _Bwd, _a=emb, _b=lin):      # d(sum) is the gradient of both addends
            d = bw.peek(_a)
            if d is not None:
                bw.seed(_b, d)
        tape.push(share_grad)
    else:
        feats = torch.empty(T, 64, device=arg1.device, dtype=torch.float32)
        t_lin(tape, pv, arg2.reshape(T, 1), "flux_embd.weight", "flux_embd.bias", need_dX=False, Y=feats[:, 0:32])
        P.sincos_feat(arg1.reshape(T), aux["div_half"], feats[:, 32:64])
        _alias_feature_grads(tape, feats, [(0, 32)])
        emb = mlp2(tape, pv, "spectrafc", feats, need_dX=True)
    pe = sin_mlp(tape, pv, "phase_embd_layer", phase.reshape(B), aux["div_full"])
    ctx = t_concat_tokens(tape, emb.view(B, L, 32), pe.view(B, 1, 32)).view(B * (L + 1), 32)
    # one extra, never-masked key for the phase token (:129-131): handled in-kernel by mask_len = L
    return _encoder_tail(tape, pv, ctx, B, L + 1, mask)


def _alias_feature_grads(tape: Tape, feats, col_ranges):
    """Call BEFORE the consumer of `feats` is taped.  At backward time (i.e. right after the consumer wrote
    d(feats)) the column ranges of that gradient are exposed as the gradients of the row-strided views the
    feature producers wrote into."""
    views = [feats[:, a:b] for a, b in col_ranges]

    def bwd(bw: _Bwd):
        g = bw.take(feats)
        if g is None:
            return
        for v, (a, b) in zip(views, col_ranges):
            bw.seed(v, g[:, a:b])
    tape.push(bwd)


def _encoder_tail(tape, pv: PView, ctx, B: int, Lc: int, mask):
    nq = pv["initbottleneck"].shape[0]
    x0 = t_expand(tape, pv["initbottleneck"].view(1, nq * 32), B, pgname="initbottleneck").view(B * nq, 32)
    h = x0
    for i in range(_num_blocks(pv)):
        h = block_forward(tape, pv, f"transformerblocks.{i}", h, ctx, None, mask, B, nq, Lc)
    out = head(tape, pv, "bottleneckfc", x0, h)
    return out.view(B, nq, out.shape[-1])


def photo_decoder_forward(tape, pv: PView, aux, time, band, z, mask, copies: int):
    """PhotometricLayers.py:49-69.  z [copies*B, T, Z]; time/band/mask are the un-replicated [B, L] inputs."""
    _check_geometry(pv, "bandembd.weight")
    B, L = time.shape
    Nb = copies * B
    q0 = sin_mlp(tape, pv, "sinusoidal_time_embd", time.reshape(B * L), aux["div_full"])
    t_gather(tape, pv, band.reshape(B * L), "bandembd.weight", q0, accumulate=True)
    x0 = t_expand(tape, q0.view(B, L * 32), copies).view(Nb * L, 32)
    Tl = z.shape[1]
    ctx = mlp2(tape, pv, "contextfc", z.reshape(Nb * Tl, z.shape[2]))
    h = x0
    for i in range(_num_blocks(pv)):
        h = block_forward(tape, pv, f"transformerblocks.{i}", h, ctx, mask, None, Nb, L, Tl, shared=(q0, copies) if i == 0 else None)
    return head(tape, pv, "get_photo", x0, h).view(Nb, L)


def spectra_decoder_forward(tape, pv: PView, aux, wavelength, phase, z, mask, copies: int):
    """SpectraLayers.py:46-63."""
    _check_geometry(pv, "get_flux.fc1.weight")
    B, L = wavelength.shape
    Nb = copies * B
    q0 = sin_mlp(tape, pv, "wavelength_embd_layer", wavelength.reshape(B * L), aux["div_full"])
    x0 = t_expand(tape, q0.view(B, L * 32), copies).view(Nb * L, 32)
    pe = sin_mlp(tape, pv, "phase_embd_layer", phase.reshape(B), aux["div_full"])
    pe_n = t_expand(tape, pe, copies)
    Tl = z.shape[1]
    cz = mlp2(tape, pv, "contextfc", z.reshape(Nb * Tl, z.shape[2]))
    ctx = t_concat_tokens(tape, cz.view(Nb, Tl, 32), pe_n.view(Nb, 1, 32)).view(Nb * (Tl + 1), 32)
    h = x0
    for i in range(_num_blocks(pv)):
        h = block_forward(tape, pv, f"transformerblocks.{i}", h, ctx, mask, None, Nb, L, Tl + 1, shared=(q0, copies) if i == 0 else None)
    return head(tape, pv, "get_flux", x0, h).view(Nb, L)
