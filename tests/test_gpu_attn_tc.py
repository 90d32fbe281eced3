"""Parity of the tcgen05 attention kernels (attn_tc.cu: four-warpgroup forward, fused backward) on the B200, through
the C ABI.  Tolerance: rel 1e-3 (north_star's fp32/TF32 bound) on max|err|/max|ref| per tensor; the
measured errors are ~2e-4 (P and dS enter the second product as tf32).  Dropout: the kernels' mask is
restated bit-exactly in numpy (tests/attn_tc_ref.py) and fed to an fp64 reference."""
import numpy as np
import pytest
import torch

import ops_cases as OC
import attn_tc_ref as R
from helpers import rel_err

pytestmark = pytest.mark.gpu
TC_TOL = 1e-3

CASES = [c for c in OC.ATTN_CASES_FULL if c["Lq"] >= 256 and c["Lk"] >= 256] + [
    dict(id="self_982_nomask", N=2, Lq=982, Lk=982, mask=False, packed="qkv"),
    dict(id="self_300x260_mask", N=3, Lq=300, Lk=260, mask=True, packed="q+kv"),
    dict(id="self_1024_mask", N=2, Lq=1024, Lk=1024, mask=True, packed="qkv"),
    dict(id="self_257_mask", N=2, Lq=257, Lk=257, mask=True, packed="qkv"),
    dict(id="cross_640x385", N=2, Lq=640, Lk=385, mask=True, mask_len=300, packed="q+kv"),
    dict(id="self_96_mask", N=3, Lq=96, Lk=96, mask=True, packed="qkv"),                     # the window's lower edge: one partial tile
    dict(id="cross_130x100_mask", N=3, Lq=130, Lk=100, mask=True, packed="q+kv"),
    dict(id="self_200_mask_rowmod", N=4, Lq=200, Lk=200, mask=True, mask_rows=2, packed="qkv"),
]


def _grads(c, dev):
    if c["packed"] == "qkv":
        dqkv = torch.full((c["N"], c["Lq"], 96), float("nan"), device=dev)
        return dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:]
    dq = torch.full((c["N"], c["Lq"], 32), float("nan"), device=dev)
    dkv = torch.full((c["N"], c["Lk"], 64), float("nan"), device=dev)
    return dq, dkv[..., :32], dkv[..., 32:]


@pytest.mark.parametrize("scale", [1.0, 3.0])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["id"])
def test_tc_attention_matches_fp64(case, scale):
    OC.run_attn_case(case, "cuda", tol=TC_TOL, scale=scale)


# fp16 operands of the tcgen05 kernels (V in the forward; Q, K, V, dO in the fused backward) are range-managed by exact
# powers of two, so magnitudes far outside fp16's 6e-5 .. 65504 must not change the accuracy.  (q, k, v, dO) factors; the
# q*k product is kept moderate so the softmax stays non-degenerate.
RANGE_SCALES = [(300.0, 1.0 / 300.0, 1e5, 1.0), (1e-4, 1e4, 1e-6, 1e6), (3e5, 1e-5 / 3.0, 30.0, 1e-7), (1.0, 1.0, 7e4, 3e4)]


@pytest.mark.parametrize("scales", RANGE_SCALES, ids=lambda s: "q%g_k%g_v%g_g%g" % s)
@pytest.mark.parametrize("case", [CASES[0], CASES[3]], ids=lambda c: c["id"])
def test_tc_attention_operand_range(case, scales):
    OC.run_attn_case(case, "cuda", tol=TC_TOL, scale=scales)


def test_tc_backward_rows_with_underflowed_upstream_gradient_stay_finite():
    """A sample whose importance weight underflowed hands the decoder a gradient of ~1e-38 (seen in training at K=8): its
    row must come out finite (the normalisers are clamped; an unclamped 2^126 times another normaliser overflowed to inf and
    inf * 0 poisoned the whole parameter gradient), and the other rows keep their accuracy."""
    from VAESNe import _ops as P
    case = CASES[0]
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, "cuda")
    dO = dO.clone()
    dO[1] *= 1e-38
    dO[3] = 0.0
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = OC.attn_reference(q, k, v, mask_full, dO)
    md = mask.to("cuda")
    O, LSE = P.attn_fwd(qd, kd, vd, md)
    dq, dk, dv = _grads(case, "cuda")
    P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to("cuda"), dq, dk, dv)
    for name, got, ref in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert torch.isfinite(got).all(), name
        assert rel_err(got.cpu()[[0, 2]], ref[[0, 2]]) < TC_TOL, (name, rel_err(got.cpu()[[0, 2]], ref[[0, 2]]))
        assert got[[1, 3]].abs().max().item() < 1e-30, name


@pytest.mark.parametrize("case", CASES[:2] + CASES[-2:], ids=lambda c: c["id"])
def test_tc_attention_dropout_mask_is_the_restated_one(case):
    from VAESNe import _ops as P
    dev = "cuda"
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, dev)
    seed_val, sid, p = 0x1234ABCD5678EF01, 77, 0.1
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev)
    drop = P.Drop(p, seed, sid)
    N, Lq, Lk = case["N"], case["Lq"], case["Lk"]
    keep = R.keep_mask(seed_val, sid, p, N, 4, Lq, None if mask_full is None else mask_full.numpy(), Lk)
    _, dscale = R.drop_threshold(p)
    unmasked = ~mask_full.numpy() if mask_full is not None else np.ones((N, Lk), bool)
    rate = 1.0 - keep[np.broadcast_to(unmasked[:, None, None, :], keep.shape)].mean()
    assert abs(rate - p) < 2e-3, rate                       # the mask really drops ~p of the live entries
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = R.attn_reference_drop(q, k, v, mask_full, dO, torch.from_numpy(keep), dscale)
    md = mask.to(dev) if mask is not None else None
    O, LSE = P.attn_fwd(qd, kd, vd, md, drop)
    assert rel_err(O.cpu(), o_ref) < TC_TOL, ("O", rel_err(O.cpu(), o_ref))
    assert rel_err(LSE.cpu(), lse_ref) < TC_TOL
    dq, dk, dv = _grads(case, dev)
    P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to(dev), dq, dk, dv, drop)
    for name, got, ref in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert rel_err(got.cpu(), ref) < TC_TOL, (name, rel_err(got.cpu(), ref))


def test_tc_masked_keys_get_exact_zero_gradients():
    from VAESNe import _ops as P
    case = CASES[0]
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, "cuda")
    md = mask.to("cuda")
    O, LSE = P.attn_fwd(qd, kd, vd, md)
    dq, dk, dv = _grads(case, "cuda")
    P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to("cuda"), dq, dk, dv)
    mf = mask_full.to("cuda")
    assert torch.equal(dk[mf], torch.zeros_like(dk[mf])) and torch.equal(dv[mf], torch.zeros_like(dv[mf]))
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()


@pytest.mark.parametrize("case", [c for c in OC.ATTN_CASES_FULL if c["id"] in ("cross_982x5", "cross_60x4", "self_60x60_mask_rowmod")], ids=lambda c: c["id"])
def test_general_and_few_key_kernels_dropout_mask_is_the_restated_one(case):
    """attn.cu / attn_small.cu: the counter-hash dropout mask restated in numpy, forward and backward."""
    from VAESNe import _ops as P
    dev = "cuda"
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, dev)
    seed_val, sid, p = 0x0BADC0DE12345678, 41, 0.1
    seed = torch.tensor([seed_val], dtype=torch.int64, device=dev)
    drop = P.Drop(p, seed, sid)
    N, Lq, Lk = case["N"], case["Lq"], case["Lk"]
    keep, dscale = R.keep_mask_general(seed_val, sid, p, N, 4, Lq, Lk)
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = R.attn_reference_drop(q, k, v, mask_full, dO, torch.from_numpy(keep), dscale)
    md = mask.to(dev) if mask is not None else None
    O, LSE = P.attn_fwd(qd, kd, vd, md, drop)
    assert rel_err(O.cpu(), o_ref) < OC.TOL, ("O", rel_err(O.cpu(), o_ref))
    assert rel_err(LSE.cpu(), lse_ref) < OC.TOL
    dq, dk, dv = _grads(case, dev)
    P.attn_bwd(qd, kd, vd, md, O, LSE, dO.to(dev), dq, dk, dv, drop)
    for name, got, ref in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert rel_err(got.cpu(), ref) < OC.TOL, (name, rel_err(got.cpu(), ref))
