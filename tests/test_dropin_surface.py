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


@pytest.mark.parametrize("heads", [2, 8])
def test_unsupported_head_count_raises_instead_of_computing_with_four(emu, heads):
    """num_heads is invisible in the parameter shapes, so the stacks check the nn.MultiheadAttention containers
    (reference ctor: util_layers.py:265-271).  A wrong answer here used to be silent (max abs diff 0.03 on outputs of scale 1)."""
    import torch
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.util_layers import TransformerBlock
    g = torch.Generator().manual_seed(0)
    px = (torch.randn(2, 12, generator=g), torch.randn(2, 12, generator=g), torch.randint(0, 2, (2, 12), generator=g),
          torch.zeros(2, 12, dtype=torch.bool))
    m = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=2, model_dim=32, num_heads=heads, ff_dim=32, num_layers=1, dropout=0.0)
    with pytest.raises(NotImplementedError, match="num_heads"):
        m.encode(px)
    with pytest.raises(NotImplementedError, match="num_heads"):
        m(px, K=1)
    sx = (torch.randn(2, 20, generator=g), torch.randn(2, 20, generator=g), torch.randn(2, generator=g), torch.zeros(2, 20, dtype=torch.bool))
    s = SpectraVAE(latent_len=4, latent_dim=2, model_dim=32, num_heads=heads, ff_dim=32, num_layers=1, dropout=0.0)
    with pytest.raises(NotImplementedError, match="num_heads"):
        s.encode(sx)
    blk = TransformerBlock(32, heads, 32, 0.0)
    with pytest.raises(NotImplementedError):
        blk(torch.randn(1, 4, 32), torch.randn(1, 3, 32))
    # the default-constructor geometry (model_dim 64) is refused with the same clear message, at the first forward
    d = PhotometricVAE(num_bands=2)
    with pytest.raises(NotImplementedError, match="model_dim=32"):
        d.encode(px)
    # the supported geometry still runs
    ok = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.0)
    assert ok.encode(px).shape == (2, 4, 2)
