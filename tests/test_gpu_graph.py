"""CUDA-graph replay of the training step equals the eager step (training_util.GraphedStep)."""
import pytest
import torch

import bench

pytestmark = pytest.mark.gpu


def _run(graph: bool, steps: int = 6):
    from VAESNe import _noise
    from VAESNe.losses import m_iwae
    from VAESNe.optim import FusedAdamW
    from VAESNe.training_util import training_step

    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    net = bench.build_model(dev, dropout=0.0)
    opt = FusedAdamW(net.parameters(), lr=1e-3)
    B, K = 8, 2
    batches = [bench.synth_batch(B, seed=10 + i) for i in range(steps)]
    g = torch.Generator().manual_seed(5)
    T, Z = net.vaes[0].latent_len, net.vaes[0].latent_dim
    noise = [(torch.rand(K, B, T, Z, generator=g) * 1.8 - 0.9).to(dev) for _ in net.vaes]

    def loss_fn(model, x):
        _noise.inject(noise)            # the same draw every step, so eager and replayed steps see identical inputs
        return m_iwae(model, x, K=K)

    losses = [training_step(net, opt, [b], loss_fn=loss_fn, multimodal=True, cuda_graph=graph) for b in batches]
    _noise.clear()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    return losses, flat, opt


def test_graphed_step_matches_eager():
    le, pe, _ = _run(False)
    lg, pg, opt = _run(True)
    table = opt.__dict__["_vaesne_graphed"]
    entry = next(iter(next(iter(table.values())).entries.values()))
    assert "graph" in entry, "the step was not captured"
    for a, b in zip(le, lg):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(a)), (le, lg)
    # the loss of step t depends on every earlier update, so the trajectory above is the real check; parameters whose
    # gradient is analytically zero (key biases: softmax is shift-invariant) carry round-off that Adam normalises to
    # +-lr steps in either run, so only the bulk is compared tightly
    d = (pe - pg).abs()
    assert float(d.median()) < 1e-6 and float((d > 1e-4).float().mean()) < 0.02, (float(d.median()), float((d > 1e-4).float().mean()))
    assert float(d.max()) <= 2 * len(le) * 1e-3
