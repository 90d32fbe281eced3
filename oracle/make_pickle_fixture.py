"""Whole-module pickle of the REFERENCE's photospecMMVAE, as cannon/test_photospectra.py:153 / ZTF_photospect.py:147 write
checkpoints (TEST INFRASTRUCTURE; run in the build container only):

    python oracle/make_pickle_fixture.py

The pickle stores class paths (``VAESNe.mmVAE.photospecMMVAE``, ``torch.nn.MultiheadAttention`` ...) and tensors — no
source — so unpickling it with THIS repository's package on the path instantiates the drop-in classes around the reference's
attribute dictionaries.  tests/test_reference_pickle.py checks that such a checkpoint loads, encodes and trains."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/package")

from oracle import vaesne_oracle as O  # noqa: E402
from VAESNe.PhotometricVAE import PhotometricVAE  # noqa: E402  (reference)
from VAESNe.SpectraVAE import SpectraVAE  # noqa: E402
from VAESNe.mmVAE import photospecMMVAE  # noqa: E402
from VAESNe.losses import m_iwae  # noqa: E402

torch.manual_seed(7)
pv = PhotometricVAE(num_bands=2, latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.1, selfattn=False)
sv = SpectraVAE(latent_len=4, latent_dim=4, model_dim=32, num_heads=4, ff_dim=32, num_layers=1, dropout=0.1, selfattn=True)
m = photospecMMVAE([pv, sv], beta=0.5)
x = [O.synth_photometry(3, 24, 2, seed=5), O.synth_spectra(3, 48, seed=5)]
out = os.path.join(ROOT, "tests", "golden")
torch.save(m, os.path.join(out, "ref_module_mm.pth"))
m.eval()
torch.manual_seed(11)
us = [O.draw_noise("laplace", (2, 3, 4, 4)) for _ in range(2)]
torch.manual_seed(11)
with torch.no_grad():
    loss = m_iwae(m, x, K=2)
torch.save({"x": x, "enc0": m.vaes[0].encode(x[0]), "enc1": m.vaes[1].encode(x[1]), "us": us, "loss_eval": float(loss)},
           os.path.join(out, "ref_module_mm_out.pth"))
print("ok", float(loss), os.path.getsize(os.path.join(out, "ref_module_mm.pth")))
