"""Every name the in-scope cannon/*.py scripts import from the package exists here (photometry / spectra / mmVAE /
contrastive / regression scripts; the image-VAE and plotting scripts are out of scope).  A missing name is an
ImportError on line 12 of cannon/ZTF_photospect.py, long before any kernel runs."""
import importlib

import pytest

SURFACE = {
    "VAESNe.training_util": ["training_step"],
    "VAESNe.mmVAE": ["photospecMMVAE"],
    "VAESNe.data_util": ["get_goldstein_params", "multimodalDataset"],
    "VAESNe.losses": ["elbo", "m_iwae", "_m_iwae", "negInfoNCE"],
    "VAESNe.SpectraVAE": ["SpectraVAE", "BrightSpectraVAE"],                 # cannon/ZTF_photospect.py:12, test_photospectra.py:12
    "VAESNe.PhotometricVAE": ["PhotometricVAE", "BrightPhotometricVAE"],     # cannon/ZTF_photospect.py:13
    "VAESNe.regression": ["VAEregressionHead", "specend2endregression", "photoend2endregression",
                          "contrasspecregressionHead", "contrasphotoregressionHead"],
    "VAESNe.contrastiveNets": ["ContraPhotSpec"],
}


@pytest.mark.parametrize("module", sorted(SURFACE))
def test_names_exist(module):
    mod = importlib.import_module(module)
    for name in SURFACE[module]:
        assert hasattr(mod, name), f"{module}.{name}"


def test_bright_constructors_take_the_reference_arguments():
    from VAESNe.PhotometricVAE import BrightPhotometricVAE
    from VAESNe.SpectraVAE import BrightSpectraVAE
    p = BrightPhotometricVAE(num_bands=2, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1,
                             dropout=0.1, selfattn=False, beta=0.5)
    s = BrightSpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.1,
                         selfattn=True, beta=0.5)
    assert p.brightnessfc.mlp[0].in_features == 4 and s.brightnessfc.mlp[0].in_features == 5
    assert list(p.state_dict())[-4:] == ["brightnessfc.mlp.0.weight", "brightnessfc.mlp.0.bias", "brightnessfc.mlp.2.weight", "brightnessfc.mlp.2.bias"]
    with pytest.raises(AssertionError):
        BrightSpectraVAE(latent_len=1)
