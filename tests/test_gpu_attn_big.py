"""Parity of the tcgen05 attention kernels at a shape that FILLS the machine (thousands of CTAs, 1-2 per SM, the dQ
red.global.add accumulation and TMEM allocation under contention) — the regime bench.py runs, which the small-N tests never
reach.  Reference: fp64 torch attention evaluated ON THE GPU in row chunks, with the kernels' dropout mask restated in torch
integer arithmetic (same hash as tests/attn_tc_ref.py / csrc/common.cuh).  Tolerance: rel 1e-3 of each tensor's scale."""
import math

import numpy as np
import pytest
import torch

import attn_tc_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-3
M32 = 0xFFFFFFFF


def _mix32(h):
    h = h ^ (h >> 16); h = (h * 0x7feb352d) & M32
    h = h ^ (h >> 15); h = (h * 0x846ca68b) & M32
    return h ^ (h >> 16)


def _hash_ctr(s0, s1, stream, lo, hi):
    h = _mix32((((lo * 0x9E3779B1) & M32) + s0) & M32)
    h = _mix32(h ^ ((((hi * 0x85EBCA77) & M32) + s1) & M32))
    return _mix32((h + ((stream * 0xC2B2AE3D) & M32)) & M32)


def _keep_chunk(seed, stream, p, n0, n1, H, Lq, Lk, key_mask):
    """bool [n1-n0, H, Lq, Lk] on the GPU; key_mask bool [rows, Lk] (True = masked) or None, row n uses n % rows."""
    dev = "cuda"
    s0, s1 = seed & M32, (seed >> 32) & M32
    thr, _ = R.drop_threshold(p)
    n = torch.arange(n0, n1, device=dev, dtype=torch.int64)
    nh = n[:, None] * H + torch.arange(H, device=dev, dtype=torch.int64)[None, :]                      # [n, H]
    ctr = nh[:, :, None] * Lq + torch.arange(Lq, device=dev, dtype=torch.int64)[None, None, :]         # [n, H, Lq]
    A = _hash_ctr(s0, s1, stream, ctr & M32, ctr >> 32) & 0xFFFF
    if key_mask is None:
        slot = torch.arange(Lk, device=dev, dtype=torch.int64)[None, :].expand(n1 - n0, Lk)
    else:
        km = key_mask[n % key_mask.shape[0]]
        slot = torch.cumsum((~km).to(torch.int64), 1) - 1                                                # compacted key slot
    B = _hash_ctr(s1, s0, (stream ^ 0x5bd1e995) & M32, slot[:, None, :].expand(n1 - n0, H, Lk), nh[:, :, None].expand(n1 - n0, H, Lk)) & 0xFFFF
    r = A[:, :, :, None] ^ B[:, :, None, :]                  # 16-bit words, read as fp16 bit patterns
    return ~_dropped(r, thr)


def _dropped(r, thr):
    """fp16 ordered compare r >= thr on integer bit patterns (NaN patterns: False; -0 == +0)."""
    def key(x):                                              # monotone integer key of a non-NaN fp16 pattern
        mag = x & 0x7FFF
        return torch.where((x & 0x8000) != 0, -mag, mag)
    nan = (r & 0x7FFF) > 0x7C00
    t = torch.tensor(thr, device=r.device, dtype=torch.int64)
    return (key(r) >= key(t)) & ~nan


def _check(N, p, chunk=32):
    from VAESNe import _ops as P
    dev = "cuda"
    Lq = Lk = 982
    g = torch.Generator(device=dev).manual_seed(1234)
    qkv = torch.randn(N, Lq, 96, device=dev, generator=g)
    dO = torch.randn(N, Lq, 32, device=dev, generator=g)
    mask = torch.rand(64, Lk, device=dev, generator=g) < 0.15
    mask[:, 0] = False
    mask[::2, 900:] = True                                   # a padded tail on half of the rows
    seed_val, sid = 0x1234ABCD5678EF01, 91
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev)
    drop = P.Drop(p, seed, sid) if p > 0 else P.NO_DROP
    _, dscale = R.drop_threshold(p) if p > 0 else (0, 1.0)
    q, k, v = qkv[..., :32], qkv[..., 32:64], qkv[..., 64:]
    O, LSE = P.attn_fwd(q, k, v, mask, drop)
    dqkv = torch.full((N, Lq, 96), float("nan"), device=dev)
    P.attn_bwd(q, k, v, mask, O, LSE, dO, dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:], drop)
    torch.cuda.synchronize()
    assert torch.isfinite(O).all() and torch.isfinite(dqkv).all()

    err = {k_: 0.0 for k_ in ("O", "LSE", "dq", "dk", "dv")}
    ref = dict(err)
    for n0 in range(0, N, chunk):
        n1 = min(N, n0 + chunk)
        qd = q[n0:n1].double().requires_grad_(); kd = k[n0:n1].double().requires_grad_(); vd = v[n0:n1].double().requires_grad_()
        B = n1 - n0
        qh = qd.view(B, Lq, 4, 8).transpose(1, 2) * math.sqrt(1 / 8)
        kh = kd.view(B, Lk, 4, 8).transpose(1, 2)
        vh = vd.view(B, Lk, 4, 8).transpose(1, 2)
        km = mask[torch.arange(n0, n1, device=dev) % 64]
        s = (qh @ kh.transpose(-1, -2)).masked_fill(km[:, None, None, :], float("-inf"))
        lse = torch.logsumexp(s, -1)
        pr = torch.softmax(s, -1)
        if p > 0:
            pr = pr * _keep_chunk(seed_val, sid, p, n0, n1, 4, Lq, Lk, mask).double() * dscale
        o = (pr @ vh).transpose(1, 2).reshape(B, Lq, 32)
        o.backward(dO[n0:n1].double())
        for name, got, want in (("O", O[n0:n1], o.detach()), ("LSE", LSE[n0:n1], lse.detach()), ("dq", dqkv[n0:n1, :, :32], qd.grad),
                                ("dk", dqkv[n0:n1, :, 32:64], kd.grad), ("dv", dqkv[n0:n1, :, 64:], vd.grad)):
            err[name] = max(err[name], (got.double() - want).abs().max().item())
            ref[name] = max(ref[name], want.abs().max().item())
        del s, pr, o, lse
    rel = {k_: err[k_] / ref[k_] for k_ in err}
    assert max(rel.values()) < TOL, rel
    return rel


def test_keep_mask_restatement_matches_numpy():
    """The torch-on-GPU restatement used below equals the numpy one the small tests are pinned to."""
    g = torch.Generator().manual_seed(3)
    km = torch.rand(2, 300, generator=g) < 0.3
    km[:, 0] = False
    full = km[torch.arange(5) % 2].numpy()
    want = R.keep_mask(0x1234ABCD5678EF01, 77, 0.1, 5, 4, 260, full, 300)
    got = _keep_chunk(0x1234ABCD5678EF01, 77, 0.1, 0, 5, 4, 260, 300, km.cuda()).cpu().numpy()
    live = np.broadcast_to(~full[:, None, None, :], want.shape)
    assert np.array_equal(got[live], want[live])


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_tc_attention_full_machine(p):
    rel = _check(2048, p)
    print("attn_tc N=2048 p=%g rel errors: %s" % (p, {k: "%.2e" % v for k, v in rel.items()}))
