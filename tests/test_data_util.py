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
