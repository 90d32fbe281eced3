"""Worker of the world_size-2 data-parallel test (test infrastructure): one process per rank, gloo on CPU, the kernels through
the CPU emulator.  Each rank takes its contiguous shard of the global batch (and of the injected noise), runs
loss -> backward (asynchronous per-stack bucket all-reduce) -> FusedAdamW, and rank 0 saves gradients and updated parameters."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "vaesne-dev_b200"), ROOT]


def build(objective):
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.mmVAE import photospecMMVAE
    torch.manual_seed(3)
    pv = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.0)
    if objective == "elbo":
        return pv
    if objective == "contrast":
        from VAESNe.contrastiveNets import ContraPhotSpec
        torch.manual_seed(3)
        return ContraPhotSpec(4, 4, 8, 2, 32, 4, 32, 1, 0.0, 32, 4, 1, 32, 0.0, False)
    sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.0, selfattn=True)
    return photospecMMVAE([pv, sv], beta=0.5)


def data(B):
    from oracle import vaesne_oracle as O
    x = [O.synth_photometry(B, 16, 2, seed=9), O.synth_spectra(B, 24, seed=9)]
    g = torch.Generator().manual_seed(4)
    us = [torch.rand(2, B, 4, 4, generator=g) * 1.8 - 0.9 for _ in range(2)]
    return x, us


def step_torch_optimizer_accumulating(model, x, us):
    """What a reference script does — torch.optim.AdamW, nothing that knows about the buckets — plus gradient accumulation
    over two backward passes: the gradients must be fully reduced when backward() returns, and the second pass must add
    REDUCED values to the first pass's (reduced) gradients."""
    from VAESNe import _noise
    from VAESNe.losses import m_iwae
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    total = 0.0
    for _ in range(2):
        _noise.clear()
        _noise.inject(us)
        loss = -m_iwae(model, x, K=2)
        loss.backward()
        total += float(loss.detach())
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}     # read BEFORE any optimiser
    opt.step()
    params = {n: p.detach().clone() for n, p in model.named_parameters()}
    return total, grads, params


def step(model, x, us, objective, average):
    from VAESNe import _noise
    from VAESNe.losses import elbo, m_iwae
    from VAESNe.optim import FusedAdamW
    if objective == "accumulate":
        return step_torch_optimizer_accumulating(model, x, us)
    opt = FusedAdamW(model.parameters(), lr=1e-2, grad_average=average)
    _noise.clear()
    if objective == "elbo":
        _noise.inject([us[0]])
        loss = -elbo(model, x[0], K=2)
    elif objective == "contrast":
        from VAESNe.losses import negInfoNCE
        loss = -negInfoNCE(model, x, temperature=0.1)
    else:
        _noise.inject(us)
        loss = -m_iwae(model, x, K=2)
    loss.backward()
    # FusedAdamW.step waits for the bucket all-reduces; read the (reduced) gradients after it
    opt.step()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    params = {n: p.detach().clone() for n, p in model.named_parameters()}
    return float(loss.detach()), grads, params


def main():
    objective, out = sys.argv[1], sys.argv[2]
    import build_emu
    from VAESNe import _native, parallel
    _native.use_library(build_emu.build())
    rank, world, _ = parallel.init_from_env("gloo")
    model = build(objective)
    parallel.broadcast_parameters(model)
    B = 4
    x, us = data(B)
    xs = parallel.shard(x, rank, world, multimodal=True)
    n = B // world
    us_s = [u[:, rank * n:(rank + 1) * n].contiguous() for u in us]
    loss, grads, params = step(model, xs, us_s, objective, average=(objective == "elbo"))
    t = parallel.all_reduce_scalar(torch.tensor([loss]), average=(objective == "elbo"))
    if rank == 0:
        torch.save({"loss": float(t), "grads": grads, "params": params}, out)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
