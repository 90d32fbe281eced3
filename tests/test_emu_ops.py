"""Per-op checks of the CUDA kernel sources, run through the CPU emulator (tests/emu) against
plain fp32/fp64 torch references.  The same checks run on the real B200 in test_gpu_ops.py."""
import math

import pytest
import torch

from helpers import rel_err
import ops_cases as OC


@pytest.mark.parametrize("case", OC.LIN_CASES_SMALL, ids=lambda c: c["id"])
def test_lin(emu, case):
    OC.run_lin_case(case, "cpu")


@pytest.mark.parametrize("case", OC.ATTN_CASES_SMALL, ids=lambda c: c["id"])
def test_attn(emu, case):
    OC.run_attn_case(case, "cpu")


def test_misc(emu):
    OC.run_misc_cases("cpu")


@pytest.mark.parametrize("fam", ["laplace", "normal"])
def test_latent_and_objectives(emu, fam):
    OC.run_latent_case(fam, "cpu")
    OC.run_loglik_case(fam, "cpu")


def test_dropout_statistics(emu):
    OC.run_dropout_case("cpu")


def test_adamw(emu):
    OC.run_adamw_case("cpu")


@pytest.mark.parametrize("case", [c for c in OC.ATTN_CASES_SMALL if c["id"] in ("cross_60x4", "self_60x60_mask_rowmod", "cross_8x200_mask")],
                         ids=lambda c: c["id"])
def test_attention_dropout_mask_is_the_restated_one(emu, case):
    """attn_small.cu / attn_mid.cu / attn.cu on the emulator: the counter-hash dropout mask restated in numpy
    (tests/attn_tc_ref.py), forward and backward — the masks of the three paths are the same function of (seed, stream, index)."""
    import torch
    import attn_tc_ref as R
    from helpers import rel_err
    from VAESNe import _ops as P
    (q, k, v, mask, mask_full, dO), (qd, kd, vd) = OC.make_attn_inputs(case, "cpu")
    seed_val, sid, p = 0x0BADC0DE12345678, 41, 0.1
    seed = torch.tensor([seed_val], dtype=torch.int64)
    drop = P.Drop(p, seed, sid)
    N, Lq, Lk = case["N"], case["Lq"], case["Lk"]
    keep, dscale = R.keep_mask_general(seed_val, sid, p, N, 4, Lq, Lk)
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = R.attn_reference_drop(q, k, v, mask_full, dO, torch.from_numpy(keep), dscale)
    O, LSE = P.attn_fwd(qd, kd, vd, mask, drop)
    assert rel_err(O, o_ref) < OC.TOL and rel_err(LSE, lse_ref) < OC.TOL
    if case["packed"] == "qkv":
        dqkv = torch.zeros(N, Lq, 96); dq, dk, dv = dqkv[..., :32], dqkv[..., 32:64], dqkv[..., 64:]
    else:
        dq = torch.zeros(N, Lq, 32); dkv = torch.zeros(N, Lk, 64); dk, dv = dkv[..., :32], dkv[..., 32:]
    P.attn_bwd(qd, kd, vd, mask, O, LSE, dO, dq, dk, dv, drop)
    for name, got, ref in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        assert rel_err(got, ref) < OC.TOL, (name, rel_err(got, ref))


def test_empty_inputs(emu):
    import edge_cases
    edge_cases.run_empty("cpu")


@pytest.mark.parametrize("L", [5, 40, 100])
def test_fully_masked_row_is_nan_like_the_reference(emu, L):
    import edge_cases
    edge_cases.run_fully_masked_row("cpu", L)


def test_boundary_lengths(emu):
    import edge_cases
    edge_cases.run_boundary_lengths("cpu", edge_cases.BOUNDARY_SMALL)


def test_more_rows_than_one_launch(emu):
    import edge_cases
    edge_cases.run_many_rows("cpu")
