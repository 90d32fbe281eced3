"""reconstruct(data, K) — SURVEY 8(f)-1: forward-only cross-modal decodes at full sequence lengths against the oracle
(the decoders' first self-attention sub-layer is computed once per object and replicated over the K samples)."""
import pytest
import torch

import bench
from helpers import rel_err

pytestmark = pytest.mark.gpu


def test_reconstruct_matches_oracle_at_full_lengths():
    from oracle import vaesne_oracle as O
    from VAESNe import _noise
    dev = torch.device("cuda:0")
    B, K = 3, 5
    model = bench.build_model(dev, dropout=0.1)          # eval mode inside reconstruct: dropout is off
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cfg = O.MMVAEConfig([O.VAEConfig("photometry", 4, 4), O.VAEConfig("spectra", 4, 4)], beta=0.5)
    cfg.apply_scaling()
    x = bench.synth_batch(B, 5)
    g = torch.Generator().manual_seed(9)
    us = [torch.rand(K, B, 4, 4, generator=g) * 1.9 - 0.95 for _ in range(2)]
    with torch.no_grad():
        _, px, _ = O.mmvae_forward(params, cfg, x, us, dropout=0.0)
    _noise.clear(); _noise.inject(us)
    xd = [tuple(t.to(dev) for t in mod) for mod in x]
    rec = model.reconstruct(xd, K=K)
    _noise.clear()
    for e in range(2):
        for d in range(2):
            got = rec[e][d].cpu().reshape(px[e][d][0].shape)
            assert rel_err(got, px[e][d][0]) < 2e-4, (e, d, rel_err(got, px[e][d][0]))
