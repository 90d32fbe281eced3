"""Tensor-level wrappers over the C ABI: argument checking, output allocation, pointer plumbing.

No arithmetic happens here; every function enqueues one (or two) kernels of the native library on
torch's current stream.  2-D operands may be row-strided views (``stride(1) == 1``), which is how
column slices of concatenated feature buffers and packed q|k|v projections are addressed.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _native as N

LN_EPS = 1e-5
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
FAMILY = {"laplace": 0, "normal": 1}


class Drop:
    """One dropout mask of the step: probability, device seed cell, and a stream id."""
    __slots__ = ("p", "seed", "sid")

    def __init__(self, p: float, seed: Optional[torch.Tensor], sid: int):
        self.p = float(p) if seed is not None else 0.0
        self.seed = seed
        self.sid = int(sid) & 0xFFFFFFFF


NO_DROP = Drop(0.0, None, 0)


class Profiler:
    """Optional per-call CUDA-event timing of the heavy ops (used by bench.py for the roofline line).
    Events are recorded on torch's current stream, the one the kernels are enqueued on."""

    def __init__(self):
        self.records = {}          # key -> list of (start_event, end_event)

    def span(self, key):
        return _Span(self, key)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for key, evs in self.records.items():
            ms = [a.elapsed_time(b) for a, b in evs]
            out[key] = dict(calls=len(ms), total_ms=sum(ms), avg_ms=sum(ms) / max(len(ms), 1))
        return out


class _Span:
    def __init__(self, prof, key):
        self.prof, self.key = prof, key

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True); self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        self.b.record()
        self.prof.records.setdefault(self.key, []).append((self.a, self.b))


class _NullSpan:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullSpan()
PROFILER = None


def _span(name, *dims):
    return PROFILER.span((name,) + tuple(int(d) for d in dims)) if PROFILER is not None else _NULL


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t


def _rows(t: Optional[torch.Tensor], name: str):
    """(ptr, ld) of a row-strided 2-D view."""
    if t is None:
        return 0, 0
    _f32(t, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name}: need a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]))


def _c(t: Optional[torch.Tensor], name: str):
    if t is None:
        return 0
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------
def lin_fwd(X, W, b, *, Xadd=None, act=ACT_NONE, H=None, R=None, gamma=None, beta=None, S=None,
            drop: Drop = NO_DROP, Y=None):
    T, K = X.shape
    Nn = W.shape[0]
    if W.shape[1] != K:
        raise ValueError(f"lin_fwd: X is [*, {K}] but W is {tuple(W.shape)}")
    if Y is None:
        Y = torch.empty(T, Nn, device=X.device, dtype=torch.float32)
    xp, ldx = _rows(X, "X"); ap, lda = _rows(Xadd, "Xadd"); hp, ldh = _rows(H, "H"); rp, ldr = _rows(R, "R"); yp, ldy = _rows(Y, "Y")
    with _span("lin_fwd_ln" if R is not None else "lin_fwd", T, K, Nn):
        N.check(N.lib().vaesne_lin_fwd(xp, ldx, ap, lda, T, K, Nn, _c(W, "W"), _c(b, "b"), act, hp, ldh, rp, ldr,
                                       _c(gamma, "gamma"), _c(beta, "beta"), LN_EPS, _c(S, "S"),
                                       drop.p, N.ptr(drop.seed), drop.sid, yp, ldy, N.stream_of(X)))
    return Y


def lin_bwd(dY, X, W, *, Xadd=None, S=None, gamma=None, dgamma=None, dbeta=None, dR=None, dR_acc=False,
            drop: Drop = NO_DROP, act=ACT_NONE, A=None, dW=None, db=None, dX=None, dX_acc=False):
    T, Nn = dY.shape
    K = W.shape[1]
    dyp, lddy = _rows(dY, "dY"); xp, ldx = _rows(X, "X"); xap, ldxa = _rows(Xadd, "Xadd")
    ap, lda = _rows(A, "A"); drp, lddr = _rows(dR, "dR"); dxp, lddx = _rows(dX, "dX")
    with _span("lin_bwd_ln" if S is not None else "lin_bwd", T, K, Nn):
        N.check(N.lib().vaesne_lin_bwd(dyp, lddy, T, K, Nn, _c(S, "S"), _c(gamma, "gamma"), LN_EPS, _c(dgamma, "dgamma"), _c(dbeta, "dbeta"),
                                       drp, lddr, int(dR_acc), drop.p, N.ptr(drop.seed), drop.sid, act, ap, lda,
                                       xp, ldx, xap, ldxa, _c(W, "W"), _c(dW, "dW"), _c(db, "db"), dxp, lddx, int(dX_acc), N.stream_of(dY)))


# ------------------------------------------------------------------------------------------------
def _attn_operand(t: torch.Tensor, name: str):
    """[N, L, C] view with unit inner stride and token stride ld; batch stride must be L*ld."""
    _f32(t, name)
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError(f"{name}: need [N, L, C] with unit inner stride")
    ld = t.stride(1)
    if t.shape[0] > 1 and t.stride(0) != t.shape[1] * ld:
        raise ValueError(f"{name}: batch stride {t.stride(0)} != L*ld {t.shape[1] * ld}")
    return t.data_ptr(), ld


def _mask_args(mask: Optional[torch.Tensor]):
    if mask is None:
        return 0, 1, 0
    if mask.dtype != torch.bool or mask.dim() != 2 or not mask.is_contiguous():
        raise ValueError("key_padding_mask must be a contiguous 2-D torch.bool tensor")
    return mask.data_ptr(), mask.shape[0], mask.shape[1]


_MAX_ROWS = 65535          # batch rows of one attention launch (a CUDA grid dimension)


def _row_chunks(Nb: int, mask):
    """Row ranges of at most _MAX_ROWS whose starts are multiples of the mask's row count (the kernels index it n % rows)."""
    if Nb <= _MAX_ROWS:
        return None
    period = mask.shape[0] if mask is not None else 1
    step = max(period, (_MAX_ROWS // period) * period)
    return [(a, min(Nb, a + step)) for a in range(0, Nb, step)]


# ---- sequences beyond the 1024 tokens the tcgen05 kernels stage at once: blocks of <= 1024 queries x <= 1024 keys -------------
_TC_MAX, _TC_MIN = 1024, 96


def _blocks(L: int):
    nb = -(-L // _TC_MAX)
    sz = -(-L // nb)
    return [(a, min(L, a + sz)) for a in range(0, L, sz)]


def _blocked(Lq: int, Lk: int, q) -> bool:
    """Long attention runs as tensor-core blocks when it is longer than one block on either side (on a CUDA device, unless the
    tcgen05 path is switched off); the key blocks of a query block are combined through their log-sum-exps."""
    import os
    if not q.is_cuda or N.is_emulated() or os.environ.get("VAESNE_NO_TC", "0") not in ("", "0"):
        return False
    return (Lq > _TC_MAX or Lk > _TC_MAX) and min(b - a for a, b in _blocks(Lq) + _blocks(Lk)) >= _TC_MIN


def _slab(t, a, b):
    """Contiguous copy of tokens a..b of a [N, L, C] view with unit inner stride (one strided block copy)."""
    Nb, L, Cc = t.shape
    out = torch.empty(Nb, b - a, Cc, device=t.device, dtype=torch.float32)
    ld = t.stride(1)
    N.check(N.lib().vaesne_copy3d(t.data_ptr() + 4 * a * ld, L * ld, ld, out.data_ptr(), (b - a) * Cc, Cc, Nb, b - a, Cc, 0, N.stream_of(t)))
    return out


def _unslab(src, dst, a, accumulate):
    """dst[:, a:a+Lb, :] (+)= src for a [N, L, C] view dst with unit inner stride."""
    Nb, Lb, Cc = src.shape
    ld = dst.stride(1)
    N.check(N.lib().vaesne_copy3d(src.data_ptr(), Lb * Cc, Cc, dst.data_ptr() + 4 * a * ld, dst.shape[1] * ld, ld, Nb, Lb, Cc, int(accumulate),
                                  N.stream_of(src)))


def _mask_block(mask, a, b):
    if mask is None or a >= mask.shape[1]:
        return None                                        # keys beyond mask_len are never masked
    return mask[:, a:min(b, mask.shape[1])].contiguous()


def _block_drop(drop: Drop, qi: int, ki: int) -> Drop:
    return Drop(drop.p, drop.seed, drop.sid + 104729 * (qi * 8 + ki + 1)) if drop.seed is not None else drop


def _attn_fwd_blocked(q, k, v, mask, drop):
    Nb, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    O = torch.empty(Nb, Lq, 32, device=q.device, dtype=torch.float32)
    LSE = torch.empty(Nb, 4, Lq, device=q.device, dtype=torch.float32)
    kb = _blocks(Lk)
    if len(kb) > 8:
        raise ValueError(f"attention over {Lk} keys: at most 8 blocks of {_TC_MAX}")
    ks = [(_slab(k, a, b), _slab(v, a, b), _mask_block(mask, a, b)) for a, b in kb]
    for qi, (q0, q1) in enumerate(_blocks(Lq)):
        qs = _slab(q, q0, q1)
        parts = []
        for ki, (kc, vc, mc) in enumerate(ks):
            Ob = torch.empty(Nb, q1 - q0, 32, device=q.device, dtype=torch.float32)
            Lb = torch.empty(Nb, 4, q1 - q0, device=q.device, dtype=torch.float32)
            mp, mrows, mlen = _mask_args(mc)
            d = _block_drop(drop, qi, ki)
            with _span("attn_fwd", Nb, q1 - q0, kc.shape[1]):
                N.check(N.lib().vaesne_attn_fwd_ex(qs.data_ptr(), 32, kc.data_ptr(), 32, vc.data_ptr(), 32, Nb, q1 - q0, kc.shape[1], mp, mrows, mlen,
                                                   d.p, N.ptr(d.seed), d.sid, Ob.data_ptr(), 32, Lb.data_ptr(), 1, N.stream_of(q)))
            parts.append((Ob, Lb))
        N.check(N.lib().vaesne_attn_combine(N.ptr_table([p[0] for p in parts]), N.ptr_table([p[1] for p in parts]), len(parts), Nb, q1 - q0, Lq, q0,
                                            O.data_ptr(), 32, LSE.data_ptr(), N.stream_of(q)))
    return O, LSE


def _attn_bwd_blocked(q, k, v, mask, O, LSE, dO, dq, dk, dv, drop):
    Nb, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    kb = _blocks(Lk)
    ks = [(_slab(k, a, b), _slab(v, a, b), _mask_block(mask, a, b)) for a, b in kb]
    for qi, (q0, q1) in enumerate(_blocks(Lq)):
        Lb = q1 - q0
        qs, Os, dOs = _slab(q, q0, q1), _slab(O, q0, q1), _slab(dO, q0, q1)
        Ls = torch.empty(Nb, 4, Lb, device=q.device, dtype=torch.float32)          # LSE[:, :, q0:q1]
        N.check(N.lib().vaesne_copy3d(LSE.data_ptr() + 4 * q0, Lq, Lq, Ls.data_ptr(), Lb, Lb, Nb * 4, 1, Lb, 0, N.stream_of(q)))
        ws = torch.empty(Nb, 4, Lb, device=q.device, dtype=torch.float32)
        for ki, ((k0, k1), (kc, vc, mc)) in enumerate(zip(kb, ks)):
            dqb = torch.empty(Nb, Lb, 32, device=q.device, dtype=torch.float32)
            dkb = torch.empty(Nb, k1 - k0, 32, device=q.device, dtype=torch.float32)
            dvb = torch.empty(Nb, k1 - k0, 32, device=q.device, dtype=torch.float32)
            mp, mrows, mlen = _mask_args(mc)
            d = _block_drop(drop, qi, ki)
            with _span("attn_bwd", Nb, Lb, k1 - k0):
                N.check(N.lib().vaesne_attn_bwd_ex(qs.data_ptr(), 32, kc.data_ptr(), 32, vc.data_ptr(), 32, Nb, Lb, k1 - k0, mp, mrows, mlen,
                                                   d.p, N.ptr(d.seed), d.sid, Os.data_ptr(), 32, Ls.data_ptr(), dOs.data_ptr(), 32, ws.data_ptr(),
                                                   dqb.data_ptr(), 32, dkb.data_ptr(), 32, dvb.data_ptr(), 32, 1, N.stream_of(q)))
            _unslab(dqb, dq, q0, accumulate=ki > 0)
            _unslab(dkb, dk, k0, accumulate=qi > 0)
            _unslab(dvb, dv, k0, accumulate=qi > 0)


def attn_fwd(q, k, v, mask, drop: Drop = NO_DROP):
    """q [N,Lq,32] view, k/v [N,Lk,32] views -> O [N,Lq,32], LSE [N,4,Lq]."""
    Nb, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    chunks = _row_chunks(Nb, mask)
    if chunks is not None:      # more batch rows than one launch takes (K=100 reconstructions of large batches): rows are independent
        outs = [attn_fwd(q[a:b], k[a:b], v[a:b], mask, Drop(drop.p, drop.seed, drop.sid + 7919 * (i + 1)) if drop.seed is not None else drop)
                for i, (a, b) in enumerate(chunks)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
    qp, ldq = _attn_operand(q, "q"); kp, ldk = _attn_operand(k, "k"); vp, ldv = _attn_operand(v, "v")
    mp, mrows, mlen = _mask_args(mask)
    if _blocked(Lq, Lk, q):
        return _attn_fwd_blocked(q, k, v, mask, drop)
    O = torch.empty(Nb, Lq, 32, device=q.device, dtype=torch.float32)
    LSE = torch.empty(Nb, 4, Lq, device=q.device, dtype=torch.float32)
    with _span("attn_fwd", Nb, Lq, Lk):
        N.check(N.lib().vaesne_attn_fwd(qp, ldq, kp, ldk, vp, ldv, Nb, Lq, Lk, mp, mrows, mlen, drop.p, N.ptr(drop.seed), drop.sid,
                                        O.data_ptr(), 32, LSE.data_ptr(), N.stream_of(q)))
    return O, LSE


def attn_bwd(q, k, v, mask, O, LSE, dO, dq, dk, dv, drop: Drop = NO_DROP):
    """Writes dq/dk/dv (views shaped like q/k/v)."""
    Nb, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    chunks = _row_chunks(Nb, mask)
    if chunks is not None:
        for i, (a, b) in enumerate(chunks):
            attn_bwd(q[a:b], k[a:b], v[a:b], mask, O[a:b], LSE[a:b], dO[a:b], dq[a:b], dk[a:b], dv[a:b],
                     Drop(drop.p, drop.seed, drop.sid + 7919 * (i + 1)) if drop.seed is not None else drop)
        return
    qp, ldq = _attn_operand(q, "q"); kp, ldk = _attn_operand(k, "k"); vp, ldv = _attn_operand(v, "v")
    dqp, lddq = _attn_operand(dq, "dq"); dkp, lddk = _attn_operand(dk, "dk"); dvp, lddv = _attn_operand(dv, "dv")
    dop, lddo = _attn_operand(dO, "dO")
    mp, mrows, mlen = _mask_args(mask)
    if _blocked(Lq, Lk, q):
        _attn_operand(O, "O")
        return _attn_bwd_blocked(q, k, v, mask, O, LSE, dO, dq, dk, dv, drop)
    ws = torch.empty(Nb, 4, Lq, device=q.device, dtype=torch.float32)
    with _span("attn_bwd", Nb, Lq, Lk):
        N.check(N.lib().vaesne_attn_bwd(qp, ldq, kp, ldk, vp, ldv, Nb, Lq, Lk, mp, mrows, mlen, drop.p, N.ptr(drop.seed), drop.sid,
                                        _c(O, "O"), 32, _c(LSE, "LSE"), dop, lddo, ws.data_ptr(), dqp, lddq, dkp, lddk, dvp, lddv,
                                        N.stream_of(q)))


# ------------------------------------------------------------------------------------------------
def sincos_feat(x: torch.Tensor, div: torch.Tensor, out: torch.Tensor):
    """x [T] -> out[:, 0:2*nf] (row-strided view)."""
    op, ld = _rows(out, "out")
    N.check(N.lib().vaesne_sincos_feat(_c(_f32(x, "x"), "x"), x.numel(), _c(div, "div"), div.numel(), op, ld, N.stream_of(x)))


def gather_rows(idx: torch.Tensor, table: torch.Tensor, out: torch.Tensor, accumulate=False):
    if idx.dtype != torch.int64:
        raise TypeError("band indices must be int64")
    op, ld = _rows(out, "out")
    N.check(N.lib().vaesne_gather_rows(_c(idx, "idx"), idx.numel(), _c(table, "table"), table.shape[0], op, ld, int(accumulate), N.stream_of(out)))


def scatter_rows(idx: torch.Tensor, dout: torch.Tensor, dtable: torch.Tensor):
    dp, ld = _rows(dout, "dout")
    N.check(N.lib().vaesne_scatter_rows(_c(idx, "idx"), idx.numel(), dp, ld, _c(dtable, "dtable"), dtable.shape[0], N.stream_of(dout)))


def expand_rows(src: torch.Tensor, copies: int) -> torch.Tensor:
    """[Bs, ...] -> [copies*Bs, ...] with row r reading source row r % Bs."""
    Bs = src.shape[0]
    dst = torch.empty((copies * Bs,) + tuple(src.shape[1:]), device=src.device, dtype=torch.float32)
    row = src.numel() // max(Bs, 1)
    N.check(N.lib().vaesne_expand_rows(_c(_f32(src, "src"), "src"), row, Bs, copies, dst.data_ptr(), N.stream_of(src)))
    return dst


def expand_rows_bwd(ddst: torch.Tensor, Bs: int, copies: int, dsrc: Optional[torch.Tensor] = None, accumulate=False) -> torch.Tensor:
    row = ddst.numel() // max(Bs * copies, 1)
    if dsrc is None:
        dsrc = torch.empty((Bs,) + tuple(ddst.shape[1:]), device=ddst.device, dtype=torch.float32)
    N.check(N.lib().vaesne_expand_rows_bwd(_c(ddst, "ddst"), row, Bs, copies, _c(dsrc, "dsrc"), int(accumulate), N.stream_of(ddst)))
    return dsrc


def copy3d(src, sgs, srs, dst, dgs, drs, G, R, Cc, accumulate=False, src_off=0, dst_off=0):
    N.check(N.lib().vaesne_copy3d(src.data_ptr() + 4 * src_off, sgs, srs, dst.data_ptr() + 4 * dst_off, dgs, drs, G, R, Cc,
                                  int(accumulate), N.stream_of(src)))


# ------------------------------------------------------------------------------------------------
def latent_fwd(botts: Sequence[torch.Tensor], noises: Sequence[torch.Tensor], fams: Sequence[int], T: int,
               fam_prior: int = 0, pz_mu=None, pz_s=None, want_lat=False):
    M = len(botts)
    B, twoT, Z = botts[0].shape
    K = noises[0].shape[0]
    dev = botts[0].device
    z = torch.empty(M, K, B, T, Z, device=dev, dtype=torch.float32)
    mus = [torch.empty(B, T, Z, device=dev, dtype=torch.float32) for _ in range(M)]
    ss = [torch.empty(B, T, Z, device=dev, dtype=torch.float32) for _ in range(M)]
    lat = torch.empty(M, K, B, device=dev, dtype=torch.float32) if want_lat else None
    pi = torch.empty(M, K, B, M, device=dev, dtype=torch.float32) if want_lat else None
    for t in list(botts) + list(noises):
        _c(_f32(t, "latent input"), "latent input")
    bt, nt, mt, st, ft = N.ptr_table(botts), N.ptr_table(noises), N.ptr_table(mus), N.ptr_table(ss), N.int_table(fams)
    N.check(N.lib().vaesne_latent_fwd(M, K, B, T, Z, bt, nt, ft, fam_prior, N.ptr(pz_mu), N.ptr(pz_s), z.data_ptr(), mt, st,
                                      N.ptr(lat), N.ptr(pi), N.stream_of(z)))
    return z, mus, ss, lat, pi


def latent_bwd(botts, noises, fams, T, fam_prior, pz_mu, pz_s, dz, dlat, pi, dmu_ext=None, ds_ext=None, kl_coef=0.0):
    M = len(botts)
    B, twoT, Z = botts[0].shape
    K = noises[0].shape[0]
    dbotts = [torch.empty_like(b) for b in botts]
    bt, nt, ft, dt = N.ptr_table(botts), N.ptr_table(noises), N.int_table(fams), N.ptr_table(dbotts)
    me = N.ptr_table(dmu_ext) if dmu_ext is not None else None
    se = N.ptr_table(ds_ext) if ds_ext is not None else None
    N.check(N.lib().vaesne_latent_bwd(M, K, B, T, Z, bt, nt, ft, fam_prior, N.ptr(pz_mu), N.ptr(pz_s), N.ptr(dz), N.ptr(dlat), N.ptr(pi),
                                      me, se, float(kl_coef), dt, N.stream_of(botts[0])))
    return dbotts


def kl_fwd(mu, s, fam, pz_mu, pz_s):
    B = mu.shape[0]
    kld = torch.empty(B, device=mu.device, dtype=torch.float32)
    N.check(N.lib().vaesne_kl_fwd(_c(mu, "mu"), _c(s, "s"), fam, _c(pz_mu, "pz_mu"), _c(pz_s, "pz_s"), B, mu[0].numel(), kld.data_ptr(), N.stream_of(mu)))
    return kld


def loglik_fwd(loc, x, mask, fam, scale_masked, scaling, lpx, accumulate):
    """loc [R,B,L], x [B,L], mask bool [B,L] -> lpx [R,B]."""
    R, B, L = loc.shape
    N.check(N.lib().vaesne_loglik_fwd(_c(loc, "loc"), _c(_f32(x, "x"), "x"), _c(mask, "mask"), R, B, L, fam, float(scale_masked), float(scaling),
                                      _c(lpx, "lpx"), int(accumulate), N.stream_of(loc)))


def loglik_bwd(loc, x, mask, fam, scale_masked, scaling, coef, gscale, gptr=None):
    R, B, L = loc.shape
    dloc = torch.empty_like(loc)
    N.check(N.lib().vaesne_loglik_bwd(_c(loc, "loc"), _c(x, "x"), _c(mask, "mask"), R, B, L, fam, float(scale_masked), float(scaling),
                                      _c(coef, "coef"), float(gscale), N.ptr(gptr), dloc.data_ptr(), N.stream_of(loc)))
    return dloc


def kl_bwd(mu, s, fam, pz_mu, pz_s, coef, gptr=None):
    B = mu.shape[0]
    dmu, ds = torch.empty_like(mu), torch.empty_like(s)
    N.check(N.lib().vaesne_kl_bwd(_c(mu, "mu"), _c(s, "s"), fam, _c(pz_mu, "pz_mu"), _c(pz_s, "pz_s"), B, mu[0].numel(), float(coef),
                                  N.ptr(gptr), dmu.data_ptr(), ds.data_ptr(), N.stream_of(mu)))
    return dmu, ds


def scale(src, mult=1.0, gptr=None):
    dst = torch.empty_like(src)
    N.check(N.lib().vaesne_scale(_c(src, "src"), src.numel(), float(mult), N.ptr(gptr), dst.data_ptr(), N.stream_of(src)))
    return dst


def iwae_combine(lat, lpx, want_lw=False):
    R, B = lpx.shape
    w = torch.empty_like(lpx)
    lw = torch.empty_like(lpx) if want_lw else None
    obj = torch.empty((), device=lpx.device, dtype=torch.float32)
    N.check(N.lib().vaesne_iwae_combine(N.ptr(lat), _c(lpx, "lpx"), R, B, w.data_ptr(), N.ptr(lw), obj.data_ptr(), N.stream_of(lpx)))
    return obj, w, lw


def elbo_combine(lpx, kld):
    K, B = lpx.shape
    obj = torch.empty((), device=lpx.device, dtype=torch.float32)
    N.check(N.lib().vaesne_elbo_combine(_c(lpx, "lpx"), _c(kld, "kld"), K, B, obj.data_ptr(), N.stream_of(lpx)))
    return obj


def adamw_flat(p, g, m, v, lr, b1, b2, eps, wd, step, grad_scale=1.0):
    N.check(N.lib().vaesne_adamw_flat(_c(p, "p"), _c(g, "g"), _c(m, "m"), _c(v, "v"), p.numel(), lr, b1, b2, eps, wd,
                                      _c(step, "step"), float(grad_scale), N.stream_of(p)))


def step_advance(step, seed):
    t = step if step is not None else seed
    N.check(N.lib().vaesne_step_advance(N.ptr(step), N.ptr(seed), N.stream_of(t)))


# ------------------------------------------------------------------------------------------------
def l2norm_fwd(x, eps=1e-12):
    B, Pd = x.shape
    y = torch.empty_like(x)
    inv = torch.empty(B, device=x.device, dtype=torch.float32)
    N.check(N.lib().vaesne_l2norm_fwd(_c(_f32(x, "x"), "x"), B, Pd, float(eps), y.data_ptr(), inv.data_ptr(), N.stream_of(x)))
    return y, inv


def l2norm_bwd(y, inv, dy, dx=None, accumulate=False):
    B, Pd = y.shape
    if dx is None:
        dx = torch.empty_like(y)
    N.check(N.lib().vaesne_l2norm_bwd(_c(y, "y"), _c(inv, "inv"), _c(_f32(dy, "dy"), "dy"), B, Pd, _c(dx, "dx"), int(accumulate), N.stream_of(y)))
    return dx


def ce_rows_fwd(A, Bm, inv_tau, label_off=0):
    """loss[i] = logsumexp_j(inv_tau * A_i . Bm_j) - inv_tau * A_i . Bm_{i + label_off}; also returns lse (saved for the backward)."""
    n, Pd = A.shape
    m = Bm.shape[0]
    lse = torch.empty(n, device=A.device, dtype=torch.float32)
    loss = torch.empty(n, device=A.device, dtype=torch.float32)
    N.check(N.lib().vaesne_ce_rows_fwd(_c(_f32(A, "A"), "A"), n, _c(_f32(Bm, "Bm"), "Bm"), m, Pd, float(inv_tau), int(label_off),
                                       lse.data_ptr(), loss.data_ptr(), N.stream_of(A)))
    return loss, lse


def ce_rows_bwd(A, Bm, inv_tau, label_off, lse, w, gptr, dA=None, dA_acc=False, dB=None, dB_acc=False):
    n, Pd = A.shape
    m = Bm.shape[0]
    N.check(N.lib().vaesne_ce_rows_bwd(_c(A, "A"), n, _c(Bm, "Bm"), m, Pd, float(inv_tau), int(label_off), _c(lse, "lse"), float(w), N.ptr(gptr),
                                       N.ptr(dA), int(dA_acc), N.ptr(dB), int(dB_acc), N.stream_of(A)))


def sum_scale(a, b, scale):
    out = torch.empty((), device=a.device, dtype=torch.float32)
    N.check(N.lib().vaesne_sum_scale(_c(a, "a"), N.ptr(b), a.numel(), float(scale), out.data_ptr(), N.stream_of(a)))
    return out


def augment(x, mask, copies, sigma_elem, sigma_row, mask_p, seed, stream_id):
    """x [B, L] float32 (or None), mask [B, L] bool (or None) -> (x_out [copies*B, L] | None, mask_out [copies*B, L] bool | None)."""
    ref = x if x is not None else mask
    B, L = (ref.shape[0], ref.shape[1]) if ref.dim() == 2 else (ref.shape[0], 1)
    R = B * int(copies)
    shape = (R,) + tuple(ref.shape[1:])
    xo = torch.empty(shape, device=ref.device, dtype=torch.float32) if x is not None else None
    mo = torch.empty(shape, device=ref.device, dtype=torch.bool) if (mask is not None or (mask_p > 0 and x is None)) else None
    if mask is None and mask_p > 0 and x is not None:
        mo = torch.empty(shape, device=ref.device, dtype=torch.bool)
    if mask is not None and (mask.dtype != torch.bool or not mask.is_contiguous()):
        raise ValueError("augment: mask must be a contiguous torch.bool tensor")
    N.check(N.lib().vaesne_augment(N.ptr(None if x is None else _f32(x, "x")), N.ptr(mask), R, B, L, float(sigma_elem), float(sigma_row), float(mask_p),
                                   seed.data_ptr(), int(stream_id) & 0xFFFFFFFF, N.ptr(xo), N.ptr(mo), N.stream_of(ref)))
    return xo, mo


_SEED_CELLS = {}


def next_seed(device) -> torch.Tensor:
    """A fresh device-resident 64-bit dropout seed (int64[1]); the per-device cell is initialised from
    torch's global generator, so torch.manual_seed controls it."""
    key = str(device)
    cell = _SEED_CELLS.get(key)
    if cell is None:
        init = int(torch.empty((), dtype=torch.int64).random_().item())
        cell = torch.tensor([init], dtype=torch.int64, device=device)
        _SEED_CELLS[key] = cell
    out = torch.empty(1, dtype=torch.int64, device=device)
    N.check(N.lib().vaesne_seed_next(cell.data_ptr(), out.data_ptr(), N.stream_of(cell)))
    return out


def masked_scale(kind_big: float) -> float:
    """fp32 value of ``1 + big`` — what ``torch.ones_like(x) + big*mask`` holds on masked entries."""
    return float(torch.tensor(1.0, dtype=torch.float32) + torch.tensor(kind_big, dtype=torch.float32))
