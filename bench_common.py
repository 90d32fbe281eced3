"""Synthetic Goldstein / ZTF-shaped batches for bench.py — shared by the product arm and the reference arm (SURVEY §8d:
photometry flux/time ~ N(0,1), band ~ U{0..nb-1}, 30 % masked with the first point observed; spectra flux ~ N(0,1), wavelength =
linspace(-1.7, 1.7, L), phase ~ N(0,1), 10 % masked plus a padded tail of up to 200 bins on every second row).  Same recipe
as the test-suite's generator; restated here so that the product arm of the benchmark never imports oracle/."""
import torch


def synth_photometry(B, L=60, num_bands=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    flux = torch.randn(B, L, generator=g)
    time = torch.randn(B, L, generator=g)
    band = torch.randint(0, num_bands, (B, L), generator=g)
    mask = torch.rand(B, L, generator=g) < 0.3
    mask[:, 0] = False
    return flux, time, band, mask


def synth_spectra(B, L=982, seed=0):
    g = torch.Generator().manual_seed(seed + 1000)
    flux = torch.randn(B, L, generator=g)
    wavelength = torch.linspace(-1.7, 1.7, L)[None].repeat(B, 1)
    phase = torch.randn(B, generator=g)
    mask = torch.rand(B, L, generator=g) < 0.1
    tail = torch.randint(0, min(201, L), (B,), generator=g)
    rows = torch.arange(0, B, 2)
    cols = torch.arange(L)[None, :]
    pad = cols >= (L - tail[rows])[:, None]
    mask[rows] |= pad & (tail[rows] > 0)[:, None]
    mask[:, 0] = False
    return flux, wavelength, phase, mask


def synth_batch(B, seed, num_bands=2, Lp=60, Ls=982):
    return [synth_photometry(B, Lp, num_bands, seed=seed), synth_spectra(B, Ls, seed=seed)]


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs, as the cannon/*.py scripts build them.  `ns` is the namespace the classes are taken from:
# the product package (vaesne-dev_b200/VAESNe) or the unmodified reference (baseline/_ref/VAESNe) — the constructors are the
# same, which is the point of a drop-in.
# ------------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (script, per-script batch, K, lr, multimodal)
    "photometry_elbo": dict(script="cannon/test_photometry.py:52-72", batch=32, K=1, lr=2.5e-4, bands=6),
    "spectra_elbo": dict(script="cannon/test_spectra.py:53-79", batch=32, K=1, lr=2.5e-4, bands=6),
    "mmvae_goldstein": dict(script="cannon/test_photospectra.py:90-135", batch=16, K=2, lr=1e-4, bands=6),
    "mmvae_ztf": dict(script="cannon/ZTF_photospect.py:76-119", batch=16, K=8, lr=1e-3, bands=2),
    "contrastive": dict(script="cannon/test_photospectra_contrast.py:89-127", batch=16, K=1, lr=1e-3, bands=6),
    "photo_end2end": dict(script="cannon/photometry2goldstein_end2end.py:45-75", batch=32, K=1, lr=1e-3, bands=6, outdim=5),
}


def build_config(name, ns, dropout=0.1):
    """-> (model, loss_fn(model, batch) -> objective to MAXIMISE (training_step negates it), make_batch(B, seed))."""
    import torch.nn.functional as F
    c = CONFIGS[name]
    nb = c["bands"]
    if name == "photometry_elbo":
        m = ns.PhotometricVAE(num_bands=nb, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                              dropout=dropout, selfattn=False, beta=0.5)
        return m, (lambda mod, x: ns.elbo(mod, x, K=1)), (lambda B, seed: synth_photometry(B, 60, nb, seed=seed)), False
    if name == "spectra_elbo":
        m = ns.SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=dropout,
                          selfattn=False, beta=1., concat=True)
        return m, (lambda mod, x: ns.elbo(mod, x, K=1)), (lambda B, seed: synth_spectra(B, 982, seed=seed)), False
    if name in ("mmvae_goldstein", "mmvae_ztf"):
        ztf = name == "mmvae_ztf"
        beta = 0.5 if ztf else 1.0
        pv = ns.PhotometricVAE(num_bands=nb, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                               dropout=dropout, selfattn=False, beta=beta)
        sv = ns.SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=dropout,
                           selfattn=ztf, beta=beta)
        m = ns.photospecMMVAE([pv, sv], beta=beta)
        K = c["K"]
        return m, (lambda mod, x: ns.m_iwae(mod, x, K=K)), (lambda B, seed: synth_batch(B, seed, nb)), True
    if name == "contrastive":
        m = ns.ContraPhotSpec(4, 4, 8, nb, 32, 4, 32, 4, dropout, 32, 4, 4, 32, dropout, False)
        return m, (lambda mod, x: ns.negInfoNCE(mod, x, temperature=0.1)), (lambda B, seed: synth_batch(B, seed, nb)), True
    if name == "photo_end2end":
        m = ns.photoend2endregression(c["outdim"], nb, 4, 4, 32, 4, 32, 4, dropout, False)

        def make(B, seed):
            x = synth_photometry(B, 60, nb, seed=seed)
            y = torch.randn(B, c["outdim"], generator=torch.Generator().manual_seed(seed + 5))
            return x + (y,)

        def loss(mod, x):       # MSE regression (script :60-75), written as an objective to maximise for training_step's "-loss_fn"
            return -F.mse_loss(mod(tuple(x[:4])), x[4])
        return m, loss, make, False
    raise KeyError(name)


class Namespace:
    """The classes / objectives of one VAESNe package (product or reference), imported lazily by the caller."""

    def __init__(self):
        from VAESNe.PhotometricVAE import PhotometricVAE
        from VAESNe.SpectraVAE import SpectraVAE
        from VAESNe.mmVAE import photospecMMVAE
        from VAESNe.contrastiveNets import ContraPhotSpec
        from VAESNe.regression import photoend2endregression
        from VAESNe.losses import elbo, m_iwae, negInfoNCE
        from VAESNe.training_util import training_step
        self.PhotometricVAE, self.SpectraVAE, self.photospecMMVAE = PhotometricVAE, SpectraVAE, photospecMMVAE
        self.ContraPhotSpec, self.photoend2endregression = ContraPhotSpec, photoend2endregression
        self.elbo, self.m_iwae, self.negInfoNCE, self.training_step = elbo, m_iwae, negInfoNCE, training_step
