import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "vaesne-dev_b200"), ROOT]
import torch
import bench_common as BC
import bench
from VAESNe import _ops as P, _stacks as S
from VAESNe.losses import m_iwae
dev = torch.device("cuda")
model = bench.build_model(dev)
orig = P.attn_bwd
def hooked(q, k, v, mask, O, LSE, dO, dq, dk, dv, drop=P.NO_DROP):
    orig(q, k, v, mask, O, LSE, dO, dq, dk, dv, drop)
    if q.shape[1] < 256: return
    bad = [(n, (~torch.isfinite(t)).sum().item()) for n, t in (("dq", dq), ("dk", dk), ("dv", dv))]
    fin_in = all(torch.isfinite(t).all().item() for t in (q, k, v, O, LSE, dO))
    print("attn_bwd", tuple(q.shape), "p", drop.p, "bad", bad, "inputs finite", fin_in,
          "max q %.3g k %.3g v %.3g dO %.3g O %.3g" % (q.abs().max(), k.abs().max(), v.abs().max(), dO.abs().max(), O.abs().max()),
          "min|dO|row", dO.abs().amax(-1).min().item(), flush=True)
    if any(b for _, b in bad):
        nb = (~torch.isfinite(dq)).any(-1).any(-1).nonzero().flatten()
        print("  bad rows (n):", nb[:20].tolist(), "count", nb.numel())
        n0 = nb[0].item()
        print("  row", n0, "dq nonfinite per head", [(~torch.isfinite(dq[n0, :, h*8:(h+1)*8])).sum().item() for h in range(4)],
              "dk", [(~torch.isfinite(dk[n0, :, h*8:(h+1)*8])).sum().item() for h in range(4)],
              "dv", [(~torch.isfinite(dv[n0, :, h*8:(h+1)*8])).sum().item() for h in range(4)])
        for h in range(4):
            sl = slice(h*8, (h+1)*8)
            print("   head", h, "max|q| %.3g |k| %.3g |v| %.3g |dO| %.3g" % (q[n0, :, sl].abs().max(), k[n0, :, sl].abs().max(), v[n0, :, sl].abs().max(), dO[n0, :, sl].abs().max()))
        torch.save(dict(q=q[n0:n0+1].contiguous().cpu(), k=k[n0:n0+1].contiguous().cpu(), v=v[n0:n0+1].contiguous().cpu(), dO=dO[n0:n0+1].contiguous().cpu(),
                        O=O[n0:n0+1].cpu(), LSE=LSE[n0:n0+1].cpu(), mask=None if mask is None else mask[n0 % mask.shape[0]].cpu()), os.path.join(ROOT, "gpurun_out", "nan_case.pt"))
        sys.exit(0)
P.attn_bwd = hooked
x = [tuple(t.to(dev) for t in mod) for mod in BC.synth_batch(512, 1000)]
loss = -m_iwae(model, x, K=8)
loss.backward()
print("loss", loss.item())
