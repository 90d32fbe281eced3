"""A whole-module checkpoint written by the REFERENCE (tests/golden/ref_module_mm.pth, made by oracle/make_pickle_fixture.py)
loads with this package on the path — unpickling instantiates the drop-in classes around the reference's attribute
dictionaries — and then encodes, evaluates the objective and trains (test infrastructure)."""
import os

import torch

from conftest import GOLDEN


def run(device):
    from VAESNe import _noise
    from VAESNe.losses import m_iwae
    from VAESNe.mmVAE import photospecMMVAE
    from VAESNe._functions import drop_p_of
    m = torch.load(os.path.join(GOLDEN, "ref_module_mm.pth"), weights_only=False)
    o = torch.load(os.path.join(GOLDEN, "ref_module_mm_out.pth"), weights_only=False)
    assert type(m) is photospecMMVAE
    m.to(device)
    x = [tuple(t.to(device) for t in mod) for mod in o["x"]]
    e0, e1 = m.vaes[0].encode(x[0]), m.vaes[1].encode(x[1])
    assert float((e0.cpu() - o["enc0"]).abs().max()) < 2e-5 and float((e1.cpu() - o["enc1"]).abs().max()) < 2e-5
    _noise.clear(); _noise.inject(o["us"])
    with torch.no_grad():                                  # eval mode (encode() left it there): dropout off, same noise
        loss = m_iwae(m, x, K=2)
    assert abs(float(loss) - o["loss_eval"]) < 2e-4 * abs(o["loss_eval"]), (float(loss), o["loss_eval"])
    # the checkpoint's dropout probability (0.1) is recovered from the pickled layers, and a training step runs
    m.train()
    assert drop_p_of(m.vaes[1].dec.generativetransformer) == 0.1 and drop_p_of(m.vaes[0].enc.inference_transformer) == 0.1
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    before = torch.cat([p.detach().reshape(-1).cpu() for p in m.parameters()])
    (-m_iwae(m, x, K=2)).backward()
    opt.step()
    after = torch.cat([p.detach().reshape(-1).cpu() for p in m.parameters()])
    assert torch.isfinite(after).all() and float((after - before).abs().max()) > 1e-5
