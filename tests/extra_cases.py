"""Cases for the contrastive-objective kernels and the device-side augmentation, shared by the CPU-emulator and GPU tests
(test infrastructure).  References: plain torch fp64 restatements of losses.py:98-110 and of the scripts' augmentation."""
import numpy as np
import torch
import torch.nn.functional as F

from helpers import rel_err


def run_infonce_case(device, B=13, Pd=8, tau=0.1):
    from VAESNe._functions import ce_rows_sum, infonce_objective, l2normalize
    g = torch.Generator().manual_seed(11)
    z1 = torch.randn(B, Pd, generator=g) * 3
    z2 = torch.randn(B, Pd, generator=g) * 0.5
    a = z1.detach().double().requires_grad_(); b = z2.detach().double().requires_grad_()
    an, bn = F.normalize(a, dim=-1), F.normalize(b, dim=-1)
    logits = an @ bn.T / tau
    labels = torch.arange(B)
    ref = -(F.cross_entropy(logits, labels) + F.cross_entropy(logits.T, labels)) / 2
    (ref * 1.7).backward()
    x1 = z1.clone().to(device).requires_grad_(); x2 = z2.clone().to(device).requires_grad_()
    out = infonce_objective(x1, x2, tau)
    (out * 1.7).backward()
    assert abs(out.item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item())), (out.item(), ref.item())
    assert rel_err(x1.grad.cpu(), a.grad) < 2e-5 and rel_err(x2.grad.cpu(), b.grad) < 2e-5, (rel_err(x1.grad.cpu(), a.grad), rel_err(x2.grad.cpu(), b.grad))
    # the row form with a label offset and more columns than rows (the data-parallel shape: columns = gathered batch)
    m, off = 2 * B + 3, 5
    cg = torch.randn(m, Pd, generator=g)
    a2 = z1.detach().double().requires_grad_(); c2 = cg.detach().double().requires_grad_()
    lg = F.normalize(a2, dim=-1) @ F.normalize(c2, dim=-1).T / tau
    ref2 = F.cross_entropy(lg, torch.arange(B) + off, reduction="sum")
    ref2.backward()
    y1 = z1.detach().clone().to(device).requires_grad_(); yc = cg.detach().clone().to(device).requires_grad_()
    out2 = ce_rows_sum(l2normalize(y1), l2normalize(yc), 1.0 / tau, off)
    out2.backward()
    assert abs(out2.item() - ref2.item()) < 2e-6 * abs(ref2.item())
    assert rel_err(y1.grad.cpu(), a2.grad) < 2e-5 and rel_err(yc.grad.cpu(), c2.grad) < 2e-5


def run_augment_case(device, B=64, L=60, copies=10):
    from VAESNe import _ops as P
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, L, generator=g)
    mask = torch.rand(B, L, generator=g) < 0.3
    seed = torch.tensor([987654321], dtype=torch.int64, device=device)
    xd, md = x.to(device), mask.to(device)
    # index logic, bit-exact: no noise, no masking -> the 10x repeat of the scripts (x.repeat((10, 1)))
    xo, mo = P.augment(xd, md, copies, 0.0, 0.0, 0.0, seed, 1)
    assert torch.equal(xo.cpu(), x.repeat((copies, 1))) and torch.equal(mo.cpu(), mask.repeat((copies, 1)))
    # per-element noise: N(0, sigma), independent across copies; a masked point stays masked, the extra masking rate is mask_p
    xo, mo = P.augment(xd, md, copies, 0.02, 0.0, 0.1, seed, 1)
    d = (xo.cpu() - x.repeat((copies, 1))).double()
    n = d.numel()
    assert abs(d.mean().item()) < 5 * 0.02 / np.sqrt(n) and abs(d.std().item() / 0.02 - 1) < 0.02, (d.mean().item(), d.std().item())
    assert abs((d[:B] * d[B:2 * B]).mean().item()) < 5 * 0.02 ** 2 / np.sqrt(B * L)          # copies draw independent noise
    kurt = ((d / d.std()) ** 4).mean().item()
    assert abs(kurt - 3.0) < 0.15, kurt                                                         # Gaussian, not uniform (1.8)
    rep = mask.repeat((copies, 1))
    assert bool((mo.cpu() | ~rep).all())                                                        # mask_out is a superset of mask
    extra = (mo.cpu() & ~rep).double().sum() / (~rep).double().sum()
    assert abs(extra.item() - 0.1) < 0.01, extra.item()
    # per-row shift: one draw per OUTPUT row, constant along the row (phototime + 0.1 * randn(B)[:, None])
    to, _ = P.augment(xd, None, copies, 0.0, 0.1, 0.0, seed, 2)
    sh = (to.cpu() - x.repeat((copies, 1))).double()
    assert float((sh - sh[:, :1]).abs().max()) < 1e-6
    assert abs(sh[:, 0].std().item() / 0.1 - 1) < 0.12 and abs(sh[:, 0].mean().item()) < 0.02
    # deterministic given (seed, stream); another stream or seed gives another draw
    xo2, _ = P.augment(xd, md, copies, 0.02, 0.0, 0.1, seed, 1)
    assert torch.equal(xo, xo2)
    xo3, _ = P.augment(xd, md, copies, 0.02, 0.0, 0.1, seed, 7)
    assert not torch.equal(xo, xo3)


def synthetic_npz(n=12, Lp=20, Ls=50, bands=6, seed=3):
    """A mapping with the keys and conventions of the reference's preprocessed .npz (mask: 1 = observed, 0 = padded)."""
    r = np.random.default_rng(seed)
    return {"flux": r.normal(size=(n, Ls)), "wavelength": np.tile(np.linspace(-1.7, 1.7, Ls), (n, 1)), "mask": (r.random((n, Ls)) > 0.1).astype(np.int64),
            "phase": r.normal(size=(n,)), "photoflux": r.normal(size=(n, Lp)), "phototime": r.normal(size=(n, Lp)),
            "photomask": (r.random((n, Lp)) > 0.3).astype(np.int64), "photowavelength": r.integers(0, bands, size=(n, Lp)),
            "training_idx": np.arange(0, n - 4), "testing_idx": np.arange(n - 4, n)}


def run_npz_contract_case(device):
    from VAESNe.augment import GpuAugmenter, load_photospectra_npz
    from VAESNe.data_util import ResidentLoader
    npz = synthetic_npz()
    ds = load_photospectra_npz(npz, "train", device=device)
    photo, spec = ds.datasets
    assert len(ds) == 8 and photo.tensors[2].dtype == torch.int64 and photo.tensors[3].dtype == torch.bool and spec.tensors[0].dtype == torch.float32
    assert torch.equal(spec.tensors[3].cpu(), torch.tensor(npz["mask"][:8] == 0))               # True = unobserved
    assert torch.equal(photo.tensors[2].cpu(), torch.tensor(npz["photowavelength"][:8]))
    aug = GpuAugmenter.ztf(seed=5)(ds, device=device)
    p2, s2 = aug.datasets
    assert len(aug) == 80 and p2.tensors[0].shape == (80, 20) and s2.tensors[2].shape == (80,)
    assert torch.equal(p2.tensors[2].cpu(), photo.tensors[2].cpu().repeat((10, 1)))             # bands: repeated, untouched
    assert torch.equal(s2.tensors[1].cpu(), spec.tensors[1].cpu().repeat((10, 1)))              # wavelengths: repeated, untouched
    assert float((s2.tensors[0].cpu() - spec.tensors[0].cpu().repeat((10, 1))).abs().max()) < 0.01 * 6
    batch = next(iter(ResidentLoader(aug, 16, device=device)))
    assert isinstance(batch, list) and batch[0][0].shape == (16, 20) and batch[1][0].shape == (16, 50)
