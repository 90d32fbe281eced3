"""Pins the CPU oracle (oracle/vaesne_oracle.py) to the live-reference goldens.

The goldens were produced by oracle/make_golden.py from /root/reference itself;
if the restatement drifts from the reference these tests fail on CPU."""
import math

import pytest
import torch

from oracle import vaesne_oracle as O
from helpers import load_golden, golden_params, golden_x, golden_grads, rel_err, mm_config

TOL = 2e-5   # fp32 restatement vs fp32 reference: only op-ordering noise


def _grads(loss, p):
    names = [k for k, v in p.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, gs)}


def _req(p):
    for k, v in p.items():
        if v.is_floating_point() and "_pz_params" not in k:
            v.requires_grad_(True)
    return p


def _check_grads(got, want, tol):
    assert set(want) <= set(got)
    worst = max((rel_err(got[k], want[k]), k) for k in want)
    assert worst[0] < tol, worst


@pytest.mark.parametrize("name,kind,Z,beta", [("photo_elbo", "photometry", 2, 0.5), ("spec_elbo", "spectra", 4, 1.0)])
def test_elbo_matches_reference(name, kind, Z, beta):
    g = load_golden(name)
    p = _req(golden_params(g))
    x = golden_x(g, "x")
    cfg = O.VAEConfig(kind, 4, Z, beta=beta)
    u = torch.from_numpy(g["u"])
    (mu, s), (loc, _), zs = O.vae_forward(p, "", cfg, x, u) if False else O.vae_forward(_strip(p), "", cfg, x, u)
    assert rel_err(mu, g["mu"]) < TOL and rel_err(s, g["scale"]) < TOL
    assert rel_err(zs, g["zs"]) < TOL and rel_err(loc, g["loc"]) < TOL
    loss = O.elbo(_strip(p), "", cfg, x, u)
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    _check_grads(_grads(loss, p), golden_grads(g), 2e-4)
    assert rel_err(mu, g["enc_mean"]) < TOL


@pytest.mark.parametrize("name,kind,Z,beta", [("bright_photo_elbo", "photometry", 2, 0.5), ("bright_spec_elbo", "spectra", 4, 1.0)])
def test_bright_elbo_matches_reference(name, kind, Z, beta):
    g = load_golden(name)
    p = _req(golden_params(g))
    x = golden_x(g, "x")
    cfg = O.VAEConfig(kind, 4, Z, num_layers=2, beta=beta, bright=True)
    u = torch.from_numpy(g["u"])
    (mu, s), (loc, _), zs = O.vae_forward(_strip(p), "", cfg, x, u)
    assert rel_err(zs, g["zs"]) < TOL and rel_err(loc, g["loc"]) < TOL
    loss = O.elbo(_strip(p), "", cfg, x, u)
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    _check_grads(_grads(loss, p), golden_grads(g), 2e-4)


@pytest.mark.parametrize("name,kind,Z,beta", [("noconcat_photo_elbo", "photometry", 2, 0.5), ("noconcat_spec_elbo", "spectra", 4, 1.0)])
def test_concat_false_elbo_matches_reference(name, kind, Z, beta):
    g = load_golden(name)
    p = _req(golden_params(g))
    x = golden_x(g, "x")
    cfg = O.VAEConfig(kind, 4, Z, num_layers=2, beta=beta)
    u = torch.from_numpy(g["u"])
    (mu, s), (loc, _), zs = O.vae_forward(_strip(p), "", cfg, x, u)
    assert rel_err(mu, g["mu"]) < TOL and rel_err(loc, g["loc"]) < TOL
    loss = O.elbo(_strip(p), "", cfg, x, u)
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    _check_grads(_grads(loss, p), golden_grads(g), 2e-4)


def _strip(p):
    """single-VAE parameter names have no 'vaes.N' prefix; the oracle joins name + '.enc...'"""
    return {"." + k: v for k, v in p.items()}


@pytest.mark.parametrize("name", ["mm_goldstein", "mm_ztf", "mm_normal"])
def test_m_iwae_matches_reference(name):
    g = load_golden(name)
    p = _req(golden_params(g))
    x = [golden_x(g, "x0"), golden_x(g, "x1")]
    cfg = mm_config(g)
    us = [torch.from_numpy(g["u0"]), torch.from_numpy(g["u1"])]
    qs, px, zss = O.mmvae_forward(p, cfg, x, us)
    for e in range(2):
        assert rel_err(zss[e], g[f"zs{e}"]) < TOL
        for d in range(2):
            assert rel_err(px[e][d][0], g[f"loc.{e}.{d}"]) < 5e-5
    loss = O.m_iwae(p, cfg, x, us)
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    _check_grads(_grads(loss, p), golden_grads(g), 5e-4)


def test_contrastive_matches_reference():
    g = load_golden("contrast")
    p = _req(golden_params(g))
    x = [golden_x(g, "x0"), golden_x(g, "x1")]
    z1, z2 = O.contrastive_forward(p, x)
    assert rel_err(z1, g["z1"]) < TOL and rel_err(z2, g["z2"]) < TOL
    loss = O.neg_info_nce(z1, z2, 0.1)
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    _check_grads(_grads(loss, p), golden_grads(g), 5e-4)


@pytest.mark.parametrize("name,fn", [("photo_end2end", O.photo_end2end), ("spec_end2end", O.spec_end2end)])
def test_end2end_matches_reference(name, fn):
    g = load_golden(name)
    p = _req(golden_params(g))
    y = fn(p, golden_x(g, "x"))
    assert rel_err(y, g["y"]) < TOL
    loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g["target"]))
    _check_grads(_grads(loss, p), golden_grads(g), 5e-4)


# ---- torch-independent known answers --------------------------------------
def test_closed_forms():
    one = torch.tensor(1.0)
    # masked-point log-likelihood constants, fp32 evaluation of 1 + 1e8 / 1 + 1e10
    s8 = torch.ones(1) + 1e8 * torch.tensor([True])
    s10 = torch.ones(1) + 1e10 * torch.tensor([True])
    assert s8.item() == 1e8 and s10.item() == float(torch.tensor(1e10))
    assert abs(O.log_prob("laplace", one, one, s8).item() - (-math.log(2e8))) < 1e-5
    assert abs(-math.log(2e8) - (-19.1138)) < 1e-4 and abs(-math.log(2e10) - (-23.7190)) < 1e-4
    # Laplace log-prob / KL / LSE
    assert abs(O.log_prob("laplace", torch.tensor(0.3), torch.tensor(-0.2), torch.tensor(0.7)).item()
               - (-math.log(1.4) - 0.5 / 0.7)) < 1e-6
    k = O.kl("laplace", torch.tensor(0.5), torch.tensor(2.0), "laplace", torch.tensor(0.0), torch.tensor(1.0)).item()
    assert abs(k - (-math.log(2.0) + 0.5 + 2.0 * math.exp(-0.25) - 1)) < 1e-6
    assert abs(O.log_mean_exp(torch.tensor([[0.0], [math.log(3.0)]])).item() - math.log(2.0)) < 1e-6
    assert abs(982 / 60 - 16.3667) < 1e-4


def test_microbatch_split_values():
    x = [(torch.zeros(16, 60),), (torch.zeros(16, 982),)]
    big = [(torch.zeros(10 ** 6, 60),), (torch.zeros(10 ** 6, 982),)]
    assert O.microbatch_split(x, 2) == 16
    assert O.microbatch_split(big, 2) == 884249   # live reference value (SURVEY rounds to 884 250)
    assert O.microbatch_split(big, 8) == 221062
