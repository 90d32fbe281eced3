"""Whole-model parity on the B200 against the live-reference goldens (loss, reconstructions, every
parameter gradient; dropout 0, injected noise)."""
import pytest

import model_cases as MC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["photo_elbo", "spec_elbo"])
def test_elbo(name):
    MC.run_elbo_case(name, "cuda")


@pytest.mark.parametrize("name", ["mm_goldstein", "mm_ztf", "mm_normal"])
def test_m_iwae(name):
    MC.run_mm_case(name, "cuda")


def test_contrastive():
    MC.run_contrast_case("cuda")


@pytest.mark.parametrize("name", ["photo_end2end", "spec_end2end"])
def test_end2end(name):
    MC.run_end2end_case(name, "cuda")


def test_regression_head_encode_path():
    MC.run_reghead_case("cuda")


@pytest.mark.parametrize("name", ["bright_photo_elbo", "bright_spec_elbo"])
def test_bright_variants(name):
    MC.run_bright_case(name, "cuda")


def test_contras_regression_heads():
    MC.run_contras_heads_case("cuda")


def test_generate():
    MC.run_generate_case("cuda")


def test_script_flow():
    import script_flow
    script_flow.run("cuda", n=13, Lp=60, Ls=982, K=2)


@pytest.mark.parametrize("name", ["noconcat_photo_elbo", "noconcat_spec_elbo"])
def test_concat_false_embeddings(name):
    MC.run_noconcat_case(name, "cuda")


def test_reference_checkpoint_loads():
    """torch.load of a whole-module pickle written by the reference (cannon/test_photospectra.py:153)."""
    import pickle_case
    pickle_case.run("cuda")


def test_first_decoder_block_with_shared_projection_only():
    from VAESNe import _stacks
    _stacks._INPROJ_ONLY = True
    try:
        MC.run_mm_case("mm_goldstein", "cuda")
    finally:
        _stacks._INPROJ_ONLY = False
