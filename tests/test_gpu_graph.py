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


def test_graphed_encode_matches_eager_and_follows_parameter_updates():
    """vae.graph_encode = True: the third encode() of a signature is captured, later ones replay; results equal the eager
    path bit for bit, also after the parameters changed in place (the graph reads their storage) and after an optimiser
    re-homed them (new storage -> new capture)."""
    from VAESNe.optim import FusedAdamW
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = bench.build_model(dev, dropout=0.1)
    for vae, mod in ((net.vaes[0], 0), (net.vaes[1], 1)):
        xs = [tuple(t.to(dev) for t in bench.synth_batch(32, seed=40 + i)[mod]) for i in range(5)]
        want = [vae.encode(x).clone() for x in xs]
        vae.graph_encode = True
        got = [vae.encode(x).clone() for x in xs]
        assert all(torch.equal(a, b) for a, b in zip(want, got))
        entry = [e for e in vae.__dict__["_encode_graphs"].values() if "graph" in e]
        assert len(entry) == 1
        with torch.no_grad():
            for p in vae.enc.parameters():
                p.mul_(1.01)
        vae.graph_encode = False; w2 = vae.encode(xs[0]).clone()
        vae.graph_encode = True; g2 = vae.encode(xs[0]).clone()
        assert torch.equal(w2, g2) and not torch.equal(w2, want[0])
    FusedAdamW(net.parameters(), lr=1e-3)                  # re-homes every parameter into one flat buffer
    vae = net.vaes[0]
    x = tuple(t.to(dev) for t in bench.synth_batch(32, seed=50)[0])
    vae.graph_encode = False; w3 = vae.encode(x).clone()
    vae.graph_encode = True
    for _ in range(4):
        g3 = vae.encode(x).clone()
    assert torch.equal(w3, g3)
