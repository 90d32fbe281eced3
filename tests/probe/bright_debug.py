"""Dev probe: spread of the per-parameter gradient error of BrightSpectraVAE on the GPU against the oracle (CPU fp32) over
random parameter draws — the mean-centred reconstruction makes the decoder gradients cancel, which amplifies the TF32-class
round-off of the tcgen05 attention relative to a parameter's own scale."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
from oracle import vaesne_oracle as O
from helpers import rel_err
from VAESNe import _noise
from VAESNe.SpectraVAE import BrightSpectraVAE, SpectraVAE
from VAESNe.losses import elbo
for bright in ((True, False) if not os.environ.get('PLAIN') else (False,)):
  for Ls in ((300, 982) if not os.environ.get('PLAIN') else (982,)):
    for seed in (32, 33, 34):
        cls = BrightSpectraVAE if bright else SpectraVAE
        m = cls(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2, dropout=0.0, selfattn=False, beta=1.0)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        p = O.random_params(shapes, seed)
        m.load_state_dict(p); m.to("cuda").train()
        x = O.synth_spectra(2, Ls, seed=seed)
        u = O.draw_noise("laplace", (2, 2, 4, 4), generator=torch.Generator().manual_seed(seed)) if "generator" in O.draw_noise.__code__.co_varnames else torch.rand(2, 2, 4, 4) * 1.9 - 0.95
        po = {"." + k: v.clone().requires_grad_(v.is_floating_point() and "_pz_params" not in k) for k, v in p.items()}
        cfg = O.VAEConfig("spectra", 4, 4, num_layers=2, beta=1.0, bright=bright)
        lo = O.elbo(po, "", cfg, x, u); lo.backward()
        _noise.clear(); _noise.inject([u])
        l = elbo(m, tuple(t.to("cuda") for t in x), K=2); l.backward()
        worst = max((rel_err(q.grad.cpu(), po["." + n].grad), n) for n, q in m.named_parameters() if q.requires_grad and float(po["." + n].grad.abs().max()) > 1e-4)
        print(f"bright={bright} Ls={Ls} seed={seed} loss rel {abs(l.item() - lo.item()) / abs(lo.item()):.1e} worst grad rel {worst[0]:.2e} {worst[1]}", flush=True)
