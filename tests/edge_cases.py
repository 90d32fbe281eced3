"""Edge cases of the kernels behind the C ABI (test infrastructure, shared by the emulator and the GPU tests): empty inputs,
a fully masked row (the reference's softmax over an empty set is NaN), single-token sequences, and lengths just past the
limits of the specialised attention paths (8 | 9 keys, 64 | 65 tokens, 1024 | 1025 tokens on the GPU)."""
import torch

import ops_cases as OC
from helpers import rel_err


def run_empty(device):
    from VAESNe import _ops as P
    q = torch.zeros(0, 60, 96, device=device)
    O, LSE = P.attn_fwd(q[..., :32], q[..., 32:64], q[..., 64:], None)
    assert O.shape == (0, 60, 32) and LSE.numel() == 0
    dq = torch.zeros(0, 60, 96, device=device)
    P.attn_bwd(q[..., :32], q[..., 32:64], q[..., 64:], None, O, LSE, torch.zeros(0, 60, 32, device=device), dq[..., :32], dq[..., 32:64], dq[..., 64:])
    X = torch.zeros(0, 32, device=device)
    W, b = torch.randn(32, 32).to(device), torch.randn(32).to(device)
    Y = P.lin_fwd(X, W, b)
    assert Y.shape == (0, 32)


def run_fully_masked_row(device, L):
    """Row 1 has every key masked: torch's MultiheadAttention returns NaN for it (softmax of all -inf) and so do the kernels;
    the other rows are unaffected."""
    from VAESNe import _ops as P
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(3, L, 96, generator=g)
    mask = torch.zeros(3, L, dtype=torch.bool)
    mask[1] = True
    mask[2, L // 2:] = True
    o_ref, lse_ref, *_ = OC.attn_reference(qkv[..., :32], qkv[..., 32:64], qkv[..., 64:], mask, torch.zeros(3, L, 32))
    qd = qkv.to(device)
    O, LSE = P.attn_fwd(qd[..., :32], qd[..., 32:64], qd[..., 64:], mask.to(device))
    O = O.cpu()
    assert torch.isnan(O[1]).all() and torch.isnan(o_ref[1]).all()
    tol = 1e-3 if L >= 256 else OC.TOL
    assert rel_err(O[[0, 2]], o_ref[[0, 2]]) < tol and rel_err(LSE.cpu()[[0, 2]], lse_ref[[0, 2]]) < tol


def run_boundary_lengths(device, cases):
    from VAESNe import _ops as P
    for c in cases:
        (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(c, device)
        o_ref, lse_ref, dq_ref, dk_ref, dv_ref = OC.attn_reference(q, k, v, mask_full, dO)
        md = mask.to(device) if mask is not None else None
        O, LSE = P.attn_fwd(qd, kd, vd, md)
        N, Lq, Lk = c["N"], c["Lq"], c["Lk"]
        if c["packed"] == "qkv":
            dqkv = torch.zeros(N, Lq, 96, device=device); dq, dk, dv = dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:]
        else:
            dq = torch.zeros(N, Lq, 32, device=device); dkv = torch.zeros(N, Lk, 64, device=device); dk, dv = dkv[..., :32], dkv[..., 32:]
        P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to(device), dq, dk, dv)
        tol = 1e-3 if (256 <= Lq <= 1024 and 256 <= Lk <= 1024 and str(device).startswith("cuda")) else OC.TOL
        for name, got, ref in (("O", O, o_ref), ("LSE", LSE, lse_ref), ("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
            assert rel_err(got.cpu(), ref) < tol, (c["id"], name, rel_err(got.cpu(), ref))


BOUNDARY_SMALL = [
    dict(id="one_token", N=2, Lq=1, Lk=1, mask=False, packed="qkv"),
    dict(id="cross_40x8", N=2, Lq=40, Lk=8, mask=True, packed="q+kv"),        # last few-key shape
    dict(id="cross_40x9", N=2, Lq=40, Lk=9, mask=True, packed="q+kv"),        # first short-sequence shape
    dict(id="self_64", N=2, Lq=64, Lk=64, mask=True, packed="qkv"),            # last short-sequence shape
    dict(id="self_65", N=2, Lq=65, Lk=65, mask=True, packed="qkv"),            # first general shape
    dict(id="cross_31x5", N=2, Lq=31, Lk=5, mask=False, packed="q+kv"),       # too few queries for the few-key path
]
BOUNDARY_GPU = [
    dict(id="self_255", N=2, Lq=255, Lk=255, mask=True, packed="qkv"),         # last general shape before tcgen05
    dict(id="self_256", N=2, Lq=256, Lk=256, mask=True, packed="qkv"),
    dict(id="self_1024", N=1, Lq=1024, Lk=1024, mask=True, packed="qkv"),      # largest tcgen05 shape
    dict(id="self_1025", N=1, Lq=1025, Lk=1025, mask=True, packed="qkv"),      # past it: general kernels
    dict(id="cross_300x1030", N=1, Lq=300, Lk=1030, mask=True, packed="q+kv"),
]


def run_many_rows(device):
    """More batch rows than one launch takes: the binding splits the rows (here with the limit lowered to 5)."""
    from VAESNe import _ops as P
    old = P._MAX_ROWS
    P._MAX_ROWS = 5
    try:
        run_boundary_lengths(device, [dict(id="rows_13_mask2", N=13, Lq=20, Lk=20, mask=True, mask_rows=2, packed="qkv"),
                                      dict(id="rows_12_cross", N=12, Lq=40, Lk=5, mask=False, packed="q+kv")])
    finally:
        P._MAX_ROWS = old
