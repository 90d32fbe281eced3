"""ResidentLoader: device-resident replacement of DataLoader(multimodalDataset(...), shuffle=True)."""
import torch
from torch.utils.data import TensorDataset

from VAESNe.data_util import ResidentLoader, multimodalDataset


def _data(n=37):
    g = torch.Generator().manual_seed(0)
    photo = TensorDataset(torch.randn(n, 6, generator=g), torch.randn(n, 6, generator=g), torch.randint(0, 2, (n, 6), generator=g),
                          torch.rand(n, 6, generator=g) < 0.3)
    spec = TensorDataset(torch.randn(n, 9, generator=g), torch.randn(n, 9, generator=g), torch.arange(n, dtype=torch.float32),
                         torch.rand(n, 9, generator=g) < 0.1)
    return photo, spec


def test_epoch_covers_every_sample_once_and_keeps_modalities_aligned():
    photo, spec = _data()
    torch.manual_seed(3)
    loader = ResidentLoader(multimodalDataset(photo, spec), batch_size=8)
    assert len(loader) == 5
    seen = []
    for batch in loader:
        assert isinstance(batch, list) and len(batch) == 2 and all(isinstance(m, tuple) for m in batch)
        idx = batch[1][2].long()                      # the spectra "phase" column carries the sample index
        assert torch.equal(batch[0][0], photo.tensors[0][idx]) and torch.equal(batch[0][3], photo.tensors[3][idx])
        assert torch.equal(batch[1][0], spec.tensors[0][idx]) and batch[0][2].dtype == torch.int64
        seen.append(idx)
    seen = torch.cat(seen)
    assert sorted(seen.tolist()) == list(range(37))
    again = torch.cat([b[1][2].long() for b in loader])
    assert not torch.equal(seen, again)               # a fresh permutation every epoch


def test_single_modality_no_shuffle_drop_last_and_rank_sharding():
    photo, _ = _data(20)
    loader = ResidentLoader(photo, batch_size=6, shuffle=False, drop_last=True)
    batches = list(loader)
    assert len(batches) == 3 and isinstance(batches[0], tuple)
    assert torch.equal(batches[1][0], photo.tensors[0][6:12])
    g0, g1 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    spec = TensorDataset(torch.arange(20, dtype=torch.float32))
    r0 = torch.cat([b[0] for b in ResidentLoader(spec, 4, generator=g0, rank=0, world=2)])
    r1 = torch.cat([b[0] for b in ResidentLoader(spec, 4, generator=g1, rank=1, world=2)])
    assert len(r0) == 10 and len(r1) == 10 and sorted(torch.cat([r0, r1]).tolist()) == list(range(20))


def test_training_step_consumes_the_loader(emu):
    """One epoch of the product's training_step over a ResidentLoader (CPU emulator build of the kernels)."""
    import torch.distributions as dist
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.losses import elbo
    from VAESNe.training_util import training_step
    torch.manual_seed(0)
    n = 12
    g = torch.Generator().manual_seed(1)
    photo = TensorDataset(torch.randn(n, 10, generator=g), torch.randn(n, 10, generator=g), torch.randint(0, 2, (n, 10), generator=g),
                          torch.zeros(n, 10, dtype=torch.bool))
    model = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.0,
                           prior=dist.Laplace, likelihood=dist.Laplace, posterior=dist.Laplace)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    loss = training_step(model, opt, ResidentLoader(photo, batch_size=4), loss_fn=elbo)
    assert loss == loss and abs(loss) < 1e6


def test_ranks_get_equal_step_counts_and_batch_sizes_when_n_is_not_a_multiple_of_world():
    """n = 1025, world = 2, batch 512 used to give rank 0 two batches and rank 1 one: the extra step blocks forever in the
    per-stack gradient all-reduce.  The n % world leftovers of an epoch are dropped instead."""
    spec = TensorDataset(torch.arange(1025, dtype=torch.float32))
    loaders = [ResidentLoader(spec, 512, generator=torch.Generator().manual_seed(7), rank=r, world=2) for r in range(2)]
    assert len(loaders[0]) == len(loaders[1]) == 1
    b0, b1 = [list(l) for l in loaders]
    assert [len(b[0]) for b in b0] == [len(b[0]) for b in b1] == [512]
    assert not set(b0[0][0].tolist()) & set(b1[0][0].tolist())
    for n, world, bs in ((37, 4, 5), (10, 3, 2)):
        ls = [ResidentLoader(TensorDataset(torch.arange(n, dtype=torch.float32)), bs, shuffle=False, rank=r, world=world) for r in range(world)]
        sizes = [[len(b[0]) for b in l] for l in ls]
        assert all(s == sizes[0] for s in sizes) and all(len(l) == len(ls[0]) for l in ls)


def test_fused_adamw_state_dict_round_trip_and_encode_leaves_the_rng_alone(emu):
    import torch.distributions as dist
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.losses import elbo
    from VAESNe.optim import FusedAdamW

    def make():
        torch.manual_seed(0)
        return PhotometricVAE(num_bands=2, latent_len=4, latent_dim=2, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.0)
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(4, 10, generator=g), torch.randn(4, 10, generator=g), torch.randint(0, 2, (4, 10), generator=g),
         torch.zeros(4, 10, dtype=torch.bool))

    def one_step(model, opt, seed):
        torch.manual_seed(seed)
        opt.zero_grad()
        (-elbo(model, x)).backward()
        opt.step()

    a = make(); oa = FusedAdamW(a.parameters(), lr=1e-2)
    one_step(a, oa, 11); one_step(a, oa, 12)
    sd = oa.state_dict()
    some = next(iter(sd["state"].values()))
    assert set(some) == {"step", "exp_avg", "exp_avg_sq"} and float(some["step"]) == 2.0      # torch.optim.AdamW's layout
    # resume in a fresh optimiser: the third step must equal the uninterrupted run's third step
    b = make(); b.load_state_dict(a.state_dict()); ob = FusedAdamW(b.parameters(), lr=1e-2)
    ob.load_state_dict(sd)
    one_step(a, oa, 13); one_step(b, ob, 13)
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), n
    # a resumed run without the state would restart Adam: the update differs
    c = make(); c.load_state_dict(b.state_dict()); oc = FusedAdamW(c.parameters(), lr=1e-2)
    one_step(b, ob, 14); one_step(c, oc, 14)
    assert any(not torch.equal(p, q) for p, q in zip(b.parameters(), c.parameters()))
    # encode() draws no sample: the global RNG stream is untouched (reference encode, PhotometricVAE.py:179-186)
    torch.manual_seed(5); before = torch.rand(3)
    torch.manual_seed(5); a.encode(x); after = torch.rand(3)
    assert torch.equal(before, after)
