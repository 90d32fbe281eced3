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
