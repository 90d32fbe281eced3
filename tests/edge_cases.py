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
    tol = 1e-3 if OC.is_tc_shape(L, L, device) else OC.TOL
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
        tol = 1e-3 if OC.is_tc_shape(Lq, Lk, device) else OC.TOL
        for name, got, ref in (("O", O, o_ref), ("LSE", LSE, lse_ref), ("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
            assert rel_err(got.cpu(), ref) < tol, (c["id"], name, rel_err(got.cpu(), ref))


BOUNDARY_SMALL = [
    dict(id="one_token", N=2, Lq=1, Lk=1, mask=False, packed="qkv"),
    dict(id="cross_40x8", N=2, Lq=40, Lk=8, mask=True, packed="q+kv"),        # last few-key shape
    dict(id="cross_40x9", N=2, Lq=40, Lk=9, mask=True, packed="q+kv"),        # first short-sequence shape
    dict(id="self_64", N=2, Lq=64, Lk=64, mask=True, packed="qkv"),            # last shape of the 64-token instantiation
    dict(id="self_65", N=2, Lq=65, Lk=65, mask=True, packed="qkv"),            # first of the 128-token one
    dict(id="self_128", N=2, Lq=128, Lk=128, mask=True, packed="qkv"),
    dict(id="cross_129x70", N=2, Lq=129, Lk=70, mask=True, packed="q+kv"),     # first of the 256-token one
    dict(id="cross_31x5", N=2, Lq=31, Lk=5, mask=False, packed="q+kv"),       # too few queries for the few-key path
]
BOUNDARY_GPU = [
    dict(id="self_95", N=2, Lq=95, Lk=95, mask=True, packed="qkv"),            # last one-CTA-per-row shape before tcgen05
    dict(id="self_96", N=2, Lq=96, Lk=96, mask=True, packed="qkv"),            # first tcgen05 shape (one partial tile)
    dict(id="cross_100x130", N=3, Lq=100, Lk=130, mask=True, packed="q+kv"),
    dict(id="self_200_rowmod", N=4, Lq=200, Lk=200, mask=True, mask_rows=2, packed="qkv"),
    dict(id="self_255", N=2, Lq=255, Lk=255, mask=True, packed="qkv"),
    dict(id="cross_300x100", N=2, Lq=300, Lk=100, mask=True, packed="q+kv"),
    dict(id="cross_300x90", N=2, Lq=300, Lk=90, mask=True, packed="q+kv"),     # too few keys for tcgen05, too many queries for one CTA per row: general
    dict(id="self_256", N=2, Lq=256, Lk=256, mask=True, packed="qkv"),
    dict(id="self_1024", N=1, Lq=1024, Lk=1024, mask=True, packed="qkv"),      # largest tcgen05 shape
    dict(id="self_1025", N=1, Lq=1025, Lk=1025, mask=True, packed="qkv"),      # past it: 2 x 2 blocks of 513 / 512 tokens
    dict(id="cross_300x1030", N=2, Lq=300, Lk=1030, mask=True, packed="q+kv"), # two key blocks, combined through their LSEs
    dict(id="cross_1500x400_rowmod", N=4, Lq=1500, Lk=400, mask=True, mask_rows=2, packed="q+kv"),   # two query blocks
    dict(id="self_2100_masklen", N=1, Lq=2100, Lk=2100, mask=True, mask_len=1500, packed="qkv"),      # 3 x 3 blocks; the last key block has no mask
    dict(id="cross_50x1100", N=2, Lq=50, Lk=1100, mask=True, packed="q+kv"),   # too few queries for a tensor-core block: general kernels
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


def run_window_timing(device="cuda"):
    """No performance cliff between the fast-path windows: the time per score element of forward + backward at 100 and 200
    tokens (one-CTA-per-row kernels, the lengths of real light curves beyond Goldstein's 60 points) stays within 2x of the
    60-token shape, and 95 | 96 (one-CTA-per-row | tcgen05) within 2x of each other; beyond, it only falls."""
    from VAESNe import _ops as P

    def per_element(L, N):
        g = torch.Generator().manual_seed(L)
        qkv = torch.randn(N, L, 96, generator=g).to(device)
        dO = torch.randn(N, L, 32, generator=g).to(device)
        dqkv = torch.empty_like(qkv)
        q, k, v = qkv[..., :32], qkv[..., 32:64], qkv[..., 64:]

        def step():
            O, LSE = P.attn_fwd(q, k, v, None)
            P.attn_bwd(q, k, v, None, O, LSE, dO, dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:])
        for _ in range(2):
            step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            step()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / 5 / (N * 4.0 * L * L) * 1e6          # ns per score element, fwd + bwd

    t = {L: per_element(L, max(256, int(4096 * (60.0 / L) ** 2))) for L in (60, 95, 96, 128, 200, 256, 1024, 1100, 2100)}
    assert t[95] < 2 * t[60] and t[96] < 2 * t[95] and t[95] < 2 * t[96], t
    assert t[128] < 1.2 * t[96] and t[200] < 1.2 * t[128] and t[256] < 1.2 * t[200], t
    assert t[1100] < 2.5 * t[1024] and t[2100] < 2.5 * t[1024], t                 # beyond one block: tensor-core blocks (the general kernels are 8x)
    return t


def run_blocked_dropout(device="cuda"):
    """Sequences beyond one tensor-core block with dropout: every (query block, key block) draws its own mask stream; the
    combined output stays an unbiased estimate (V = 1 -> O ~ 1), the same call is deterministic, and the backward regenerates
    the forward's masks (dV = P_dropped^T dO: with dO = 1 the column sums of the dropped probabilities add up to Lq)."""
    from VAESNe import _ops as P
    g = torch.Generator().manual_seed(1)
    Nb, L = 3, 1500
    qk = (torch.randn(Nb, L, 64, generator=g) * 0.5).to(device)
    v1 = torch.ones(Nb, L, 32, device=device)
    seed = torch.tensor([424242], dtype=torch.int64, device=device)
    drop = P.Drop(0.1, seed, 5)
    O, LSE = P.attn_fwd(qk[..., :32], qk[..., 32:], v1, None, drop)
    O2, _ = P.attn_fwd(qk[..., :32], qk[..., 32:], v1, None, drop)
    assert torch.equal(O, O2)
    assert abs(O.mean().item() - 1.0) < 5e-3 and O.std().item() > 1e-3, (O.mean().item(), O.std().item())
    dq = torch.empty(Nb, L, 32, device=device); dk = torch.empty_like(dq); dv = torch.empty_like(dq)
    P.attn_bwd(qk[..., :32], qk[..., 32:], v1, None, O, LSE, torch.ones_like(O), dq, dk, dv, drop)
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()
    # sum over keys of dV[:, key, c] = sum over (query, key) of dropped P = sum over queries of O (V = 1)
    assert abs(dv[..., 0].sum().item() - O[..., 0].sum().item()) < 2e-3 * O[..., 0].sum().item()
