"""The flow of cannon/ZTF_photospect.py:76-150 on synthetic data (test infrastructure): models built with the script's
keyword arguments, a stock DataLoader over multimodalDataset, torch.optim.AdamW, training_step with a lambda loss,
whole-module torch.save / torch.load."""
import io

import torch
from torch.optim import AdamW
from torch.utils.data import DataLoader, TensorDataset


def run(device, n=9, Lp=20, Ls=40, K=2, epochs=2):      # 9 samples at batch 4: the last batch holds a single sample
    from VAESNe.SpectraVAE import SpectraVAE
    from VAESNe.PhotometricVAE import PhotometricVAE
    from VAESNe.mmVAE import photospecMMVAE
    from VAESNe.losses import m_iwae
    from VAESNe.data_util import multimodalDataset
    from VAESNe.training_util import training_step

    g = torch.Generator().manual_seed(0)
    flux, wavelength = torch.randn(n, Ls, generator=g), torch.linspace(-1.7, 1.7, Ls)[None].repeat(n, 1)
    phase, mask = torch.randn(n, generator=g), torch.rand(n, Ls, generator=g) < 0.1
    photoflux, phototime = torch.randn(n, Lp, generator=g), torch.randn(n, Lp, generator=g)
    photoband, photomask = torch.randint(0, 2, (n, Lp), generator=g), torch.rand(n, Lp, generator=g) < 0.2
    mask[:, 0] = False; photomask[:, 0] = False
    train_loader = DataLoader(multimodalDataset(TensorDataset(photoflux, phototime, photoband, photomask),
                                                TensorDataset(flux, wavelength, phase, mask)), batch_size=4, shuffle=True)
    torch.manual_seed(1)
    my_spectravae = SpectraVAE(spectra_length=Ls, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=2,
                               dropout=0.1, selfattn=True).to(device)
    my_photovae = PhotometricVAE(photometric_length=Lp, num_bands=2, latent_len=4, latent_dim=4, model_dim=32, num_heads=4,
                                 ff_dim=32, num_layers=2, dropout=0.1, selfattn=False).to(device)
    my_mmvae = photospecMMVAE(vaes=[my_photovae, my_spectravae], beta=0.5).to(device)
    optimizer = AdamW(my_mmvae.parameters(), lr=1e-3)
    before = torch.cat([p.detach().reshape(-1).cpu() for p in my_mmvae.parameters()])
    losses = [training_step(my_mmvae, optimizer, train_loader, loss_fn=lambda model, x: m_iwae(model, x, K=K), multimodal=True)
              for _ in range(epochs)]
    after = torch.cat([p.detach().reshape(-1).cpu() for p in my_mmvae.parameters()])
    assert all(l == l and abs(l) < 1e12 for l in losses), losses
    assert float((after - before).abs().max()) > 1e-5                      # the optimiser stepped
    assert all(p.grad is not None for p in my_mmvae.parameters() if p.requires_grad)
    # whole-module checkpoint, as the scripts write it
    buf = io.BytesIO()
    torch.save(my_mmvae, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    x = tuple(t[:3].to(device) for t in (photoflux, phototime, photoband, photomask))
    assert torch.equal(my_mmvae.vaes[0].encode(x), again.vaes[0].encode(x))
    return losses
