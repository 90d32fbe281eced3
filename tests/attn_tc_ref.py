"""Checker for the tcgen05 attention kernels (test infrastructure): numpy restatement of their dropout
mask (vaesne-dev_b200/csrc/attn_tc.cu: drop_row_word / drop_col_word / keep rule; common.cuh: mix32,
hash_ctr) and an fp64 torch reference of masked attention with an explicit dropout mask."""
import math

import numpy as np
import torch

M32 = np.uint64(0xFFFFFFFF)


def _u32(x):
    return np.asarray(x, dtype=np.uint64) & M32


def mix32(h):
    h = _u32(h)
    h = h ^ (h >> np.uint64(16)); h = _u32(h * np.uint64(0x7feb352d))
    h = h ^ (h >> np.uint64(15)); h = _u32(h * np.uint64(0x846ca68b))
    h = h ^ (h >> np.uint64(16))
    return h


def hash_ctr(s0, s1, stream, ctr):
    ctr = np.asarray(ctr, dtype=np.uint64)
    lo, hi = ctr & M32, ctr >> np.uint64(32)
    h = mix32(_u32(lo * np.uint64(0x9E3779B1)) + np.uint64(s0))
    h = mix32(h ^ _u32(_u32(hi * np.uint64(0x85EBCA77)) + np.uint64(s1)))
    h = mix32(h + _u32(np.uint64(stream) * np.uint64(0xC2B2AE3D)))
    return h


def drop_threshold(p):
    """(threshold fp16 bit pattern, scale): the number of dropped 16-bit patterns is round(p * 65536), see make_tcdrop."""
    n = int(float(np.float32(p)) * 65536.0 + 0.5)         # the C ABI carries p as a float
    n = min(max(n, 1), 63490)
    if n <= 31744:
        thr = 0x7C01 - n
    else:
        n = max(n, 31746)
        thr = 0x8000 | (n - 31746)
    scale = np.float32(1.0 / (1.0 - n / 65536.0))
    return thr, float(scale)


def dropped(r16, thr):
    """r16 (uint16 array) read as fp16 compares >= the threshold pattern (NaN patterns: False)."""
    r = np.asarray(r16, dtype=np.uint16).view(np.float16)
    t = np.array([thr], dtype=np.uint16).view(np.float16)[0]
    with np.errstate(invalid="ignore"):
        return r >= t


def keep_mask(seed, stream, p, N, H, Lq, key_mask_full, Lk):
    """bool [N, H, Lq, Lk]: True = kept.  Dropout is indexed by COMPACTED key slot (masked keys removed),
    so the per-row key-padding mask enters; masked keys are reported as kept (their probability is 0)."""
    s0, s1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    thr, _ = drop_threshold(p)
    out = np.ones((N, H, Lq, Lk), dtype=bool)
    for n in range(N):
        kept = np.arange(Lk) if key_mask_full is None else np.nonzero(~key_mask_full[n])[0]
        slots = np.arange(len(kept), dtype=np.uint64)
        for h in range(H):
            nh = n * H + h
            A = hash_ctr(s0, s1, stream, np.uint64(nh) * np.uint64(Lq) + np.arange(Lq, dtype=np.uint64)) & np.uint64(0xFFFF)
            B = hash_ctr(s1, s0, (stream ^ 0x5bd1e995) & 0xFFFFFFFF, (np.uint64(nh) << np.uint64(32)) | slots) & np.uint64(0xFFFF)
            r = (A[:, None] ^ B[None, :]).astype(np.uint16)
            out[n, h][:, kept] = ~dropped(r, thr)
    return out


def attn_reference_drop(q, k, v, mask_full, dO, keep, drop_scale):
    """fp64 reference with an explicit dropout mask `keep` [N,4,Lq,Lk] (bool tensor) applied to softmax(P)."""
    q = q.double().requires_grad_(); k = k.double().requires_grad_(); v = v.double().requires_grad_()
    N, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    qh = q.view(N, Lq, 4, 8).transpose(1, 2) * math.sqrt(1 / 8)
    kh = k.view(N, Lk, 4, 8).transpose(1, 2)
    vh = v.view(N, Lk, 4, 8).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if mask_full is not None:
        s = s.masked_fill(mask_full[:, None, None, :], float("-inf"))
    lse = torch.logsumexp(s, -1)
    p = torch.softmax(s, -1) * keep.double() * drop_scale
    o = (p @ vh).transpose(1, 2).reshape(N, Lq, 32)
    o.backward(dO.double())
    return o.detach(), lse.detach(), q.grad, k.grad, v.grad


def keep_mask_general(seed, stream, p, N, H, Lq, Lk):
    """Dropout mask of the general / few-key kernels (csrc/common.cuh drop_mult; csrc/attn.cu, attn_small.cu):
    element index ((n*H + h)*Lq + i)*Lk + j, one 32-bit hash per two consecutive elements, 16-bit thresholds."""
    s0, s1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    t = float(np.float32(p)) * 65536.0 + 0.5
    thresh = 65535 if t > 65535.0 else int(t)
    scale = float(np.float32(1.0 / (1.0 - np.float32(thresh) * np.float32(1.0 / 65536.0))))
    idx = np.arange(N * H * Lq * Lk, dtype=np.uint64)
    r = hash_ctr(s0, s1, stream, idx >> np.uint64(1))
    r16 = np.where((idx & np.uint64(1)) == 1, r >> np.uint64(16), r & np.uint64(0xFFFF))
    return (r16 >= np.uint64(thresh)).reshape(N, H, Lq, Lk), scale
