"""Generate tests/golden/*.npz from the LIVE reference (TEST INFRASTRUCTURE).

Run in the build container only (the reference does not exist on the GPU box):

    python oracle/make_golden.py

For every case this script
  1. builds the *reference's own* module from /root/reference/package/VAESNe,
  2. overwrites its parameters with ``oracle.random_params(shapes, seed)`` (so a
     fixture only stores a seed, and attention biases / LN gains are non-trivial),
  3. seeds the global RNG, records the reparameterisation noise the reference
     will draw, re-seeds, and runs the reference loss + backward with dropout 0,
  4. stores inputs, noise, loss, reconstructions and every parameter gradient.

Nothing from the reference is copied; only its numerical outputs are recorded.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/package")

import torch.distributions as dist  # noqa: E402
from oracle import vaesne_oracle as O  # noqa: E402

from VAESNe.PhotometricVAE import PhotometricVAE, BrightPhotometricVAE  # noqa: E402  (reference)
from VAESNe.SpectraVAE import SpectraVAE, BrightSpectraVAE  # noqa: E402
from VAESNe.mmVAE import photospecMMVAE  # noqa: E402
from VAESNe.losses import elbo, m_iwae, negInfoNCE  # noqa: E402
from VAESNe.contrastiveNets import ContraPhotSpec  # noqa: E402
from VAESNe.regression import (photoend2endregression, specend2endregression, VAEregressionHead,  # noqa: E402
                               contrasphotoregressionHead, contrasspecregressionHead)

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


_LAST_SHAPES = {}


def load_random(model, seed):
    sd = model.state_dict()
    shapes = {k: list(v.shape) for k, v in sd.items()}
    _LAST_SHAPES.clear()
    _LAST_SHAPES.update(shapes)
    p = O.random_params(shapes, seed)
    model.load_state_dict(p)
    return p


def grads_of(model):
    return {"grad." + n: (q.grad.detach().numpy().copy() if q.grad is not None else np.zeros(tuple(q.shape), np.float32))
            for n, q in model.named_parameters() if q.requires_grad}


def pack_x(prefix, x):
    return {f"{prefix}.{i}": t.numpy() for i, t in enumerate(x)}


def record_noise(seed, shapes, family="laplace"):
    torch.manual_seed(seed)
    ns = [O.draw_noise(family, s) for s in shapes]
    torch.manual_seed(seed)
    return ns


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    arrays.setdefault("shapes", json.dumps(_LAST_SHAPES))
    np.savez(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def case_photo_elbo():
    """config 1: cannon/test_photometry.py:52-72 hyper-parameters, dropout 0, B=3, K=1."""
    m = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32,
                       num_layers=4, dropout=0.0, selfattn=False, beta=0.5)
    load_random(m, 11)
    x = O.synth_photometry(3, 60, 6, seed=1)
    (u,) = record_noise(101, [(1, 3, 4, 2)])
    m.train()
    loss = elbo(m, x, K=1)
    loss.backward()
    torch.manual_seed(101)
    with torch.no_grad():
        qz, px, zs = m(x, 1)
    enc_mean = m.encode(x)
    save("photo_elbo", seed=11, noise_seed=101, u=u.numpy(), loss=loss.item(), loc=px.loc.numpy(), zs=zs.numpy(),
         mu=qz.loc.numpy(), scale=qz.scale.numpy(), enc_mean=enc_mean.numpy(), **pack_x("x", x), **grads_of(m))


def case_spec_elbo():
    """config 2: cannon/test_spectra.py:53-79 hyper-parameters, dropout 0, B=2, K=1."""
    m = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                   dropout=0.0, selfattn=False, beta=1.0, concat=True)
    load_random(m, 12)
    x = O.synth_spectra(2, 982, seed=2)
    (u,) = record_noise(102, [(1, 2, 4, 4)])
    m.train()
    loss = elbo(m, x, K=1)
    loss.backward()
    torch.manual_seed(102)
    with torch.no_grad():
        qz, px, zs = m(x, 1)
    enc_mean = m.encode(x)
    save("spec_elbo", seed=12, noise_seed=102, u=u.numpy(), loss=loss.item(), loc=px.loc.numpy(), zs=zs.numpy(),
         mu=qz.loc.numpy(), scale=qz.scale.numpy(), enc_mean=enc_mean.numpy(), **pack_x("x", x), **grads_of(m))


def _bright(m, x, seed, noise_seed, K, name, ushape):
    load_random(m, seed)
    (u,) = record_noise(noise_seed, [ushape])
    m.train()
    loss = elbo(m, x, K=K)
    loss.backward()
    torch.manual_seed(noise_seed)
    with torch.no_grad():
        qz, px, zs = m(x, K)
    save(name, seed=seed, noise_seed=noise_seed, K=K, u=u.numpy(), loss=loss.item(), loc=px.loc.numpy(), zs=zs.numpy(),
         mu=qz.loc.numpy(), scale=qz.scale.numpy(), **pack_x("x", x), **grads_of(m))


def case_bright():
    """Bright variants (imported by cannon/ZTF_photospect.py:12-13, test_photospectra.py:12-13): K=2 ELBO, dropout 0."""
    m = BrightPhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32,
                             num_layers=2, dropout=0.0, selfattn=False, beta=0.5)
    _bright(m, O.synth_photometry(3, 60, 6, seed=31), 31, 131, 2, "bright_photo_elbo", (2, 3, 4, 2))
    m = BrightSpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2,
                         dropout=0.0, selfattn=False, beta=1.0)
    _bright(m, O.synth_spectra(2, 300, seed=32), 32, 132, 2, "bright_spec_elbo", (2, 2, 4, 4))


def case_noconcat():
    """concat=False embeddings (sum instead of concat + MLP; PhotometricLayers.py:132-135, SpectraLayers.py:124-126)."""
    m = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=2,
                       dropout=0.0, selfattn=False, concat=False, beta=0.5)
    _bright(m, O.synth_photometry(3, 60, 6, seed=41), 41, 141, 2, "noconcat_photo_elbo", (2, 3, 4, 2))
    m = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0,
                   selfattn=True, concat=False, beta=1.0)
    _bright(m, O.synth_spectra(2, 200, seed=42), 42, 142, 2, "noconcat_spec_elbo", (2, 2, 4, 4))


def _mm(num_bands, K, B, beta, selfattn_spec, seed, noise_seed, name, Lp=60, Ls=982, family=dist.Laplace):
    pv = PhotometricVAE(num_bands=num_bands, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32,
                        num_layers=4, dropout=0.0, selfattn=False, concat=True,
                        prior=family, likelihood=family, posterior=family)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                    dropout=0.0, selfattn=selfattn_spec, concat=True,
                    prior=family, likelihood=family, posterior=family)
    m = photospecMMVAE([pv, sv], prior_dist=family, beta=beta)
    load_random(m, seed)
    x = [O.synth_photometry(B, Lp, num_bands, seed=seed), O.synth_spectra(B, Ls, seed=seed)]
    fam = "laplace" if family is dist.Laplace else "normal"
    us = record_noise(noise_seed, [(K, B, 4, 4), (K, B, 4, 4)], fam)
    m.train()
    loss = m_iwae(m, x, K=K)
    loss.backward()
    torch.manual_seed(noise_seed)
    with torch.no_grad():
        qz, px, zss = m(x, K)
    locs = {f"loc.{e}.{d}": px[e][d].loc.numpy() for e in range(2) for d in range(2)}
    save(name, seed=seed, noise_seed=noise_seed, K=K, beta=beta, num_bands=num_bands, selfattn=int(selfattn_spec),
         family=fam, u0=us[0].numpy(), u1=us[1].numpy(), loss=loss.item(),
         mu0=qz[0].loc.numpy(), s0=qz[0].scale.numpy(), mu1=qz[1].loc.numpy(), s1=qz[1].scale.numpy(),
         zs0=zss[0].numpy(), zs1=zss[1].numpy(), **locs,
         **pack_x("x0", x[0]), **pack_x("x1", x[1]), **grads_of(m))
    return m, x


def case_mm_goldstein():
    """config 3: cannon/test_photospectra.py:90-135 (6 bands, K=2, beta=1), B=2."""
    m, x = _mm(6, 2, 2, 1.0, False, 13, 103, "mm_goldstein")
    # encode path through VAEregressionHead (photometry2goldstein_mmvae.py:55-57)
    head = VAEregressionHead(m.vaes[0], 5, MLPlatent=[128] * 4)
    hshapes = {k: list(v.shape) for k, v in head.outfc.state_dict().items()}
    head.outfc.load_state_dict(O.random_params(hshapes, 23))
    with torch.no_grad():
        y = head(x[0])
    np.savez(os.path.join(OUT, "mm_goldstein_reghead.npz"), seed=23, y=y.numpy(), shapes=json.dumps(hshapes),
             enc0=m.vaes[0].encode(x[0]).numpy(), enc1=m.vaes[1].encode(x[1]).numpy())


def case_mm_ztf():
    """config 4: cannon/ZTF_photospect.py:76-119 (2 bands, beta=0.5, spectra selfattn=True), K=3, B=2."""
    _mm(2, 3, 2, 0.5, True, 14, 104, "mm_ztf")


def case_mm_normal():
    """Normal prior/likelihood/posterior (north_star 'Gaussian posteriors'), short sequences."""
    _mm(6, 2, 3, 1.0, False, 15, 105, "mm_normal", Lp=12, Ls=40, family=dist.Normal)


def case_contrast():
    """config 5a: cannon/test_photospectra_contrast.py:89-127, dropout 0, B=4, tau=0.1."""
    m = ContraPhotSpec(4, 4, 8, 6, 32, 4, 32, 4, 0.0, 32, 4, 4, 32, 0.0, False)
    load_random(m, 16)
    x = [O.synth_photometry(4, 60, 6, seed=16), O.synth_spectra(4, 982, seed=16)]
    m.train()
    loss = negInfoNCE(m, x, temperature=0.1)
    loss.backward()
    with torch.no_grad():
        z1, z2 = m(x)
    save("contrast", seed=16, loss=loss.item(), z1=z1.numpy(), z2=z2.numpy(),
         **pack_x("x0", x[0]), **pack_x("x1", x[1]), **grads_of(m))


def case_end2end():
    """config 5b: cannon/photometry2goldstein_end2end.py:45-75 (outdim unknown without data -> 5), MSE, B=3."""
    m = photoend2endregression(5, 6, 4, 4, 32, 4, 32, 4, 0.0, False)
    load_random(m, 17)
    x = O.synth_photometry(3, 60, 6, seed=17)
    tgt = torch.randn(3, 5, generator=torch.Generator().manual_seed(170))
    m.train()
    y = m(x)
    loss = torch.nn.functional.mse_loss(y, tgt)
    loss.backward()
    save("photo_end2end", seed=17, loss=loss.item(), y=y.detach().numpy(), target=tgt.numpy(),
         **pack_x("x", x), **grads_of(m))

    m2 = specend2endregression(5, 4, 4, 32, 4, 4, 32, 0.0, False)
    load_random(m2, 18)
    x2 = O.synth_spectra(2, 982, seed=18)
    tgt2 = torch.randn(2, 5, generator=torch.Generator().manual_seed(180))
    m2.train()
    y2 = m2(x2)
    loss2 = torch.nn.functional.mse_loss(y2, tgt2)
    loss2.backward()
    save("spec_end2end", seed=18, loss=loss2.item(), y=y2.detach().numpy(), target=tgt2.numpy(),
         **pack_x("x", x2), **grads_of(m2))


def case_contras_heads():
    """regression.py:28-65: frozen ContraPhotSpec encoders (eval, no grad) -> flatten -> MLP; outputs + the heads' gradients."""
    net = ContraPhotSpec(4, 4, 8, 6, 32, 4, 32, 2, 0.0, 32, 4, 2, 32, 0.0, False)
    load_random(net, 51)
    net_shapes = dict(_LAST_SHAPES)
    x = [O.synth_photometry(3, 60, 6, seed=51), O.synth_spectra(3, 300, seed=51)]
    out = {}
    for tag, cls, xi, seed in (("photo", contrasphotoregressionHead, x[0], 52), ("spec", contrasspecregressionHead, x[1], 53)):
        head = cls(net, 5, MLPlatent=[64, 64])
        hshapes = {k: list(v.shape) for k, v in head.outfc.state_dict().items()}
        head.outfc.load_state_dict(O.random_params(hshapes, seed))
        tgt = torch.randn(3, 5, generator=torch.Generator().manual_seed(seed))
        y = head(xi)
        loss = torch.nn.functional.mse_loss(y, tgt)
        head.zero_grad()
        loss.backward()
        out.update({f"{tag}.y": y.detach().numpy(), f"{tag}.target": tgt.numpy(), f"{tag}.loss": loss.item(), f"{tag}.seed": seed,
                    f"{tag}.shapes": json.dumps(hshapes)})
        out.update({f"{tag}.grad.{n}": q.grad.numpy().copy() for n, q in head.outfc.named_parameters()})
        assert all(q.grad is None for q in net.parameters())          # frozen
    np.savez(os.path.join(OUT, "contras_heads.npz"), seed=51, shapes=json.dumps(net_shapes), **out,
             **pack_x("x0", x[0]), **pack_x("x1", x[1]))
    print("contras_heads")


def case_generate():
    """photospecMMVAE.generate(N, x) (mmVAE.py:108-118) and SpectraVAE.generate(N, x) (SpectraVAE.py:198-206): prior samples
    decoded on the data's positions.  The latents are drawn from torch's global RNG: they are recorded next to the outputs."""
    pv = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0)
    m = photospecMMVAE([pv, sv], beta=1.0)
    load_random(m, 61)
    x = [O.synth_photometry(2, 60, 6, seed=61), O.synth_spectra(2, 300, seed=61)]
    N = 3
    torch.manual_seed(611)
    latents = m.pz(*m.pz_params).rsample(torch.Size([N, 2]))
    torch.manual_seed(611)
    gen = m.generate(N, x)
    torch.manual_seed(612)
    lat_s = sv.pz(*sv.pz_params).rsample(torch.Size([N, 1]))
    torch.manual_seed(612)
    xs1 = tuple(t[:1] for t in x[1])
    gen_s = sv.generate(N, xs1)
    save("generate", seed=61, N=N, latents=latents.detach().numpy(), gen0=gen[0].numpy(), gen1=gen[1].numpy(),
         lat_s=lat_s.detach().numpy(), gen_s=gen_s.numpy(), **pack_x("x0", x[0]), **pack_x("x1", x[1]))


if __name__ == "__main__":
    if len(sys.argv) > 1:          # regenerate selected cases only: python oracle/make_golden.py case_generate ...
        torch.set_num_threads(8)
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    torch.set_num_threads(8)
    case_photo_elbo()
    case_spec_elbo()
    case_mm_goldstein()
    case_mm_ztf()
    case_mm_normal()
    case_contrast()
    case_end2end()
    case_bright()
    case_noconcat()
    case_contras_heads()
    case_generate()
