"""Whole-model parity of the product path against the live-reference goldens, executed through the
CPU emulator of the CUDA kernels (same checks run on the B200 in test_gpu_model.py)."""
import pytest

import model_cases as MC


def test_photo_elbo(emu):
    MC.run_elbo_case("photo_elbo", "cpu")


def test_spec_elbo(emu):
    MC.run_elbo_case("spec_elbo", "cpu")


def test_mm_normal(emu):
    MC.run_mm_case("mm_normal", "cpu")


def test_photo_end2end(emu):
    MC.run_end2end_case("photo_end2end", "cpu")


def test_spec_end2end(emu):
    MC.run_end2end_case("spec_end2end", "cpu")


def test_contrastive(emu):
    MC.run_contrast_case("cpu")


def test_regression_head_encode_path(emu):
    MC.run_reghead_case("cpu")


def test_mm_goldstein(emu):
    MC.run_mm_case("mm_goldstein", "cpu")


def test_mm_ztf(emu):
    MC.run_mm_case("mm_ztf", "cpu")


@pytest.mark.parametrize("name", ["bright_photo_elbo", "bright_spec_elbo"])
def test_bright_variants(emu, name):
    MC.run_bright_case(name, "cpu")


def test_script_flow(emu):
    """cannon/ZTF_photospect.py's construction / DataLoader / AdamW / training_step / torch.save flow."""
    import script_flow
    script_flow.run("cpu")


@pytest.mark.parametrize("name", ["noconcat_photo_elbo", "noconcat_spec_elbo"])
def test_concat_false_embeddings(emu, name):
    MC.run_noconcat_case(name, "cpu")


def test_reference_checkpoint_loads(emu):
    """torch.load of a whole-module pickle written by the reference (cannon/test_photospectra.py:153)."""
    import pickle_case
    pickle_case.run("cpu")
